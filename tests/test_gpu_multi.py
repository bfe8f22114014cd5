"""GPU tests that need >= 2 devices: the z-slab decomposition with NCCL halo exchange must give the
SAME bits as the single-domain oracle for every slab count (no reduction enters the field update)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,args", [(2, ["--N", "32", "--steps", "40"]), (2, ["--N", "21", "--Re", "400", "--steps", "30"]),
                                        (2, ["--N", "16", "--method", "0", "--Re", "20", "--diag", "1", "--steps", "25"]),
                                        (4, ["--N", "30", "--steps", "30"]), (8, ["--N", "40", "--steps", "24"]),
                                        (2, ["--N", "256", "--Re", "1000", "--steps", "6"]),     # BASELINE config 2 grid, 2 slabs
                                        (2, ["--halo", "p2p", "--N", "32", "--steps", "40"]),
                                        (2, ["--halo", "p2p", "--N", "21", "--Re", "400", "--steps", "30", "--track", "0"]),
                                        (4, ["--halo", "p2p", "--N", "30", "--steps", "30", "--diag", "1"]),
                                        (8, ["--halo", "p2p", "--N", "40", "--steps", "24"]),
                                        # two iterations per pass on slabs: two ghost planes per side, 19 planes per face
                                        (2, ["--pair", "1", "--N", "32", "--steps", "41"]),
                                        (2, ["--pair", "1", "--pair-variant", "1", "--pair-lz", "3", "--N", "21", "--Re", "400", "--steps", "30", "--track", "0"]),
                                        (2, ["--pair", "1", "--pair-variant", "2", "--N", "16", "--method", "0", "--Re", "20", "--diag", "1", "--steps", "26"]),
                                        (4, ["--pair", "1", "--N", "33", "--steps", "30"]), (8, ["--pair", "1", "--N", "40", "--steps", "25"]),
                                        (2, ["--pair", "1", "--N", "256", "--Re", "1000", "--steps", "9"]),
                                        # BASELINE configs[2] (512^3) through 8 slabs vs the 64-bit oracle; configs[3] (1024^3): invariance
                                        (8, ["--mode", "big512", "--N", "512", "--Re", "1000", "--steps", "6"]),
                                        (8, ["--mode", "invariance", "--N", "1024", "--Re", "1000", "--steps", "9"]),
                                        (2, ["--mode", "invariance", "--N", "96", "--Re", "200", "--steps", "21"]),
                                        (2, ["--dim", "2", "--N", "24", "--M", "41", "--Re", "40", "--steps", "50"]),
                                        (4, ["--dim", "2", "--N", "100", "--M", "30", "--Re", "100", "--steps", "40"])])
def test_slab_run_matches_oracle(world, args):
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_check.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    j = json.loads(lines[-1])
    assert j["bit_equal"] and j["rho_equal"] and j["residual_ok"], j
