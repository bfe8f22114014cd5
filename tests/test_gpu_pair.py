"""GPU tests (-m gpu) of the two-iterations-per-pass kernel (csrc/lbm3d_pair.cuh, option `pair`):
temporal blocking must not change a single bit -- the distributions are compared with the oracle
(bit-exact) and with the hashes recorded from the real reference."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HASHES = json.load(open(os.path.join(GOLD, "ref_hashes.json")))
RESID_RTOL = 1e-12


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def parse_tag(tag):
    parts = tag.split("_")
    return dict(N=int(parts[0][1:]), Re=int(parts[1][2:]), METHOD=int(parts[2][1:]), GAMMA=float(parts[3][1:]),
                DIAG=int(parts[4][1:]), UMAX=float(parts[5][1:]))


CASES = [(tag, sorted(int(n) for n in v)) for tag, v in HASHES.items() if not tag.startswith("unstable_")]


@pytest.mark.parametrize("tag,steps", CASES, ids=[c[0] for c in CASES])
def test_pair_kernel_matches_reference_hashes(lbm, tag, steps):
    with lbm.Cavity3D(**parse_tag(tag)) as sim:
        sim.set_option("pair", 1)
        done = 0
        for n in steps:
            sim.step(n - done)
            done = n
            rec = HASHES[tag][str(n)]
            assert hashlib.sha256(sim.f().tobytes()).hexdigest() == rec["sha256"], "%s after %d iterations" % (tag, n)
            assert sim.residual() == pytest.approx(rec["last_diff"], rel=RESID_RTOL)


@pytest.mark.parametrize("cfg,variant,lz", [
    (dict(N=12, Re=100), 0, 0), (dict(N=12, Re=100), 1, 5), (dict(N=21, Re=400), 2, 4),
    (dict(N=31, Re=100, DIAG=1), 3, 7), (dict(N=33, Re=150, MAGIC=0.1875), 4, 0),
    (dict(N=13, Re=20, METHOD=0, DIAG=1), 5, 3), (dict(N=16, Re=100, GAMMA=0.5, DIAG=1), 1, 16),
    (dict(N=61, Re=200), 1, 9), (dict(N=64, Re=253), 2, 0), (dict(N=96, Re=300), 5, 33),
    (dict(N=40, Re=30, METHOD=0), 1, 11), (dict(N=45, Re=100, GAMMA=2.0), 0, 6),
    (dict(N=35, Re=100), 6, 8), (dict(N=50, Re=100, DIAG=1), 7, 13), (dict(N=70, Re=200), 4, 0),
    (dict(N=12, Re=100), 8, 0), (dict(N=34, Re=200, DIAG=1), 8, 5), (dict(N=64, Re=253), 9, 10), (dict(N=20, Re=20, METHOD=0, GAMMA=0.5), 10, 7),
    (dict(N=96, Re=300), 8, 31), (dict(N=33, Re=100), 8, 6),   # odd N: the TMA variant falls back (needs 16-byte strides)
], ids=lambda v: "-".join("%s%s" % kv for kv in v.items()) if isinstance(v, dict) else str(v))
def test_pair_kernel_tiles_and_segments_match_oracle(lbm, oracle, cfg, variant, lz):
    """Every kernel variant (tile rows, blocks per SM, register pipelining) and z-segment length, grids that are not multiples of the tile, tracked and
    untracked stepping, odd and even iteration counts: element-wise equal to the oracle."""
    o = oracle.Oracle3D(threads=8, **cfg)
    with lbm.Cavity3D(**cfg) as sim, lbm.Cavity3D(track_residual=False, **cfg) as un:
        for s in (sim, un):
            s.set_option("pair", 1)
            s.set_option("pair_variant", variant)
            s.set_option("pair_lz", lz)
        done = 0
        for n in (7, 2, 10, 1, 21):
            o.step(n)
            sim.step(n)
            un.step(n)
            done += n
            a, b, c = sim.f(), o.f(), un.f()
            bad = np.nonzero(bits(a) != bits(b))[0]
            assert bad.size == 0, "after %d iterations: %d slots differ, first %s" % (done, bad.size, bad[:5])
            assert np.array_equal(bits(c), bits(b)), "untracked, after %d iterations" % done
            assert sim.residual() == pytest.approx(o.last_diff, rel=RESID_RTOL)
        mg, mo = sim.macrovars(), o.macrovars()
        for k in ("u1", "u2", "u3", "rho", "vmag"):
            assert np.array_equal(bits(mg[k]), bits(mo[k])), k


def test_pair_kernel_equals_single_step_kernel_at_256(lbm):
    """Size-independent property at a bench-sized grid: 2 x 20 fused iterations == 40 single iterations."""
    cfg = dict(N=256, Re=1000, track_residual=False)
    with lbm.Cavity3D(**cfg) as a:
        a.set_option("pair", 0)
        a.step(41)
        fa = hashlib.sha256(a.f().tobytes()).hexdigest()
    with lbm.Cavity3D(**cfg) as b:
        b.set_option("pair", 1)
        b.step(41)
        assert b.info()["kernel_launches"] < 41
        fb = hashlib.sha256(b.f().tobytes()).hexdigest()
    assert fa == fb


def test_pair_kernel_reports_negative_feq(lbm):
    with lbm.Cavity3D(N=12, Re=100, METHOD=0) as sim:
        sim.set_option("pair", 1)
        with pytest.raises(lbm.UnstableError):
            sim.step(400)
            sim.residual()
