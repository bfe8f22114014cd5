#!/usr/bin/env python3
"""tools/slab_probe.py N world rank [iters] -- ONE rank of an N^3 / `world` z-slab decomposition on a
single GPU with the exchanges skipped (LBM_B200_SLAB_EMULATION=1): same slab, same boundary / interior
launches as the multi-GPU run, so ncu can capture them (it must not wrap a multi-rank command).
The numbers are launch durations and DRAM traffic only -- the field is NOT the cavity's."""
import json
import os
import sys

os.environ["LBM_B200_SLAB_EMULATION"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbmpkg import lbm  # noqa: E402

N, world, rank = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 40
opts = dict(kv.split("=") for kv in sys.argv[5:])
with lbm.Cavity3D(N=N, Re=1000, rank=rank, nranks=world, device=0, track_residual=False, nccl_id=bytes(lbm.NCCL_ID_BYTES)) as sim:
    for k, v in opts.items():
        sim.set_option(k, int(v))
    sim.step(iters)
    sim.sync()
    ms = sim.step_timed(iters) / iters
    info = sim.info()
print(json.dumps({"N": N, "world": world, "rank": rank, "planes": info["k1"] - info["k0"], "opts": opts, "pair": info["pair_enabled"],
                  "ms_per_iteration": round(ms, 4), "mlups_per_gpu": round(info["cells_local"] / ms / 1e3, 1)}))
