"""CPU tests of the D2Q9 oracles (oracle/lbm2d_oracle.c).

Scheme B (literal in-place emulation of the reference's sweep with P ranks) is pinned BIT-FOR-BIT
against the real 2D reference, which oracle/build_ref.py compiles unmodified over the fork/shared-
memory MPI shim (oracle/mpi_shim/mpi.h).  Scheme A (synchronous two-lattice, what the GPU computes)
shares the cell arithmetic and must reach the same steady state."""
import numpy as np
import pytest

SMALL = [(dict(N=24, M=40, Re="40", MAXITER="200", prntInt="1", THRESH="0"), 3),
         (dict(N=24, M=40, Re="40", MAXITER="200", prntInt="1", THRESH="0"), 5),
         (dict(N=17, M=23, Re="25", MAXITER="120", prntInt="1", THRESH="0"), 4)]


@pytest.mark.parametrize("cfg,ranks", SMALL)
def test_inplace_emulation_equals_real_reference(oracle, cfg, ranks):
    import os
    if not os.path.exists(oracle.ref2d_binary(**cfg)):
        pytest.skip("oracle/_ref 2D binaries not built")
    ref, out = oracle.ref2d_run(ranks, **cfg)
    o = oracle.Oracle2D(N=cfg["N"], M=cfg["M"], Re=float(cfg["Re"]), scheme="inplace", ranks=ranks)
    res = []
    for _ in range(int(float(cfg["MAXITER"]))):
        o.step(1)
        res.append(o.residual())
    m = o.macro()
    for k in ("u", "v", "rho"):
        assert np.array_equal(m[k].view(np.uint64), ref[k].view(np.uint64)), k
    # x = 0.5 + i, y = 0.5 + j (main.cpp:146-152)
    assert np.array_equal(ref["x"].reshape(cfg["M"], cfg["N"])[0], 0.5 + np.arange(cfg["N"]))
    assert np.array_equal(ref["y"].reshape(cfg["M"], cfg["N"])[:, 0], 0.5 + np.arange(cfg["M"]))
    # the program's own last printed line: "<t>\t<df/df0>\t<MLUPS>"
    last = [l for l in out.splitlines() if l[:1].isdigit()][-1].split("\t")
    assert last[1] == "%.3e" % (res[-1] / res[0])


def test_derived_constants_2d(oracle):
    d = oracle.Oracle2D(N=100, M=300, Re=100.0).derived()
    assert d["NU"] == 0.1 * 100 / 100.0 and d["TAU"] == d["NU"] * 3.0 + 0.5 and d["OMEGA"] == 1.0 / d["TAU"]
    assert d["OMEGAm"] == 1.0 / (0.5 + (0.25 / (d["TAU"] - 0.5)))


def test_sync_first_iteration_by_hand(oracle):
    """After one iteration from rest only the lid row moves: ft7/ft8 = -+6*wd*Ulid there."""
    N, M = 8, 6
    o = oracle.Oracle2D(N=N, M=M, Re=10.0)
    o.step(1)
    d = o.dist().reshape(M, 9, N)  # reference layout: [j][k][i]
    assert np.all(d[: M - 1] == 0.0)
    assert np.all(d[M - 1] != 0.0)
    o.residual()
    u = o.macro()["u"].reshape(M, N)
    assert np.all(u[M - 1] > 0.0) and np.all(u[: M - 1] == 0.0)   # the lid drags the top row in +x


@pytest.mark.slow
def test_both_schemes_reach_the_reference_steady_state(oracle):
    """Defaults of the 2D program (N=100, M=300, Re=100, THRESH=2e-9).  SURVEY App. D recorded from the
    real program with 3 ranks: 44 000 iterations; u in [-0.02294031719, 0.09694387877];
    u(i=50, j=297..299) = 0.08339286, 0.09005163, 0.09670737."""
    N, M = 100, 300
    out = {}
    for scheme in ("inplace", "sync"):
        o = oracle.Oracle2D(N=N, M=M, Re=100.0, scheme=scheme, ranks=3)
        o.step(1)
        d0 = o.residual()
        t = 0
        while True:
            o.step(1000)
            t += 1000
            df = o.residual() / d0
            if df < 2e-9:
                break
            assert t < 200000
        out[scheme] = (t, o.macro())
    t_in, m_in = out["inplace"]
    assert t_in == 44000
    u = m_in["u"].reshape(M, N)
    assert abs(u.min() - (-0.02294031719)) < 1e-10 and abs(u.max() - 0.09694387877) < 1e-10
    assert np.allclose(u[297:300, 50], [0.08339286, 0.09005163, 0.09670737], atol=5e-9)
    t_sy, m_sy = out["sync"]
    assert t_sy == 94000   # the synchronous (Jacobi-like) scheme needs more iterations than the in-place sweep
    assert np.abs(m_sy["u"] - m_in["u"]).max() < 1e-7 and np.abs(m_sy["v"] - m_in["v"]).max() < 1e-7
    assert np.abs(m_sy["rho"] - m_in["rho"]).max() < 1e-4
