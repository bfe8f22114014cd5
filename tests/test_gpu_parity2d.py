"""GPU tests (-m gpu) of the D2Q9 variant: bit-exact against the synchronous oracle (scheme A of
oracle/lbm2d_oracle.c); the residual (order of summation differs) to rel 1e-12; steady state against
the values recorded from the real 2D reference within the convergence tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("cfg", [dict(N=24, M=40, Re=40.0), dict(N=17, M=23, Re=25.0), dict(N=100, M=300, Re=100.0),
                                 dict(N=130, M=33, Re=400.0, Ulid=0.05, MAGIC=0.1875)],
                         ids=lambda c: "N%dM%d" % (c["N"], c["M"]))
def test_every_iteration_matches_sync_oracle(lbm, oracle, cfg):
    o = oracle.Oracle2D(scheme="sync", **cfg)
    with lbm.Cavity2D(**cfg) as sim:
        for it in range(10):
            o.step(1)
            sim.step(1)
            assert np.array_equal(bits(sim.dist()), bits(o.dist())), "iteration %d" % it
            assert sim.residual() == pytest.approx(o.residual(), rel=1e-12)
        o.step(490)
        sim.step(490)
        assert np.array_equal(bits(sim.dist()), bits(o.dist()))
        assert sim.residual() == pytest.approx(o.residual(), rel=1e-12)
        mg, mo = sim.macro(), o.macro()
        for k in ("u", "v", "rho"):
            assert np.array_equal(bits(mg[k]), bits(mo[k])), k


def test_restart_and_output_files(lbm, oracle, tmp_path):
    cfg = dict(N=20, M=30, Re=50.0)
    o = oracle.Oracle2D(scheme="sync", **cfg)
    o.step(40)
    d = o.dist()
    o.step(15)
    with lbm.Cavity2D(**cfg) as sim:
        sim.set_dist(d)
        sim.step(15)
        assert np.array_equal(bits(sim.dist()), bits(o.dist()))
        paths = sim.output_bin(str(tmp_path))
    names = sorted(p.split("/")[-1] for p in paths)
    assert names == ["Re=50_N=20_M=30_%s.bin" % s for s in ("rho", "u", "v", "x", "y")]
    x = np.fromfile(str(tmp_path / "Re=50_N=20_M=30_x.bin")).reshape(30, 20)   # plot.py:12 reshapes (M, N)
    y = np.fromfile(str(tmp_path / "Re=50_N=20_M=30_y.bin")).reshape(30, 20)
    assert np.array_equal(x[0], 0.5 + np.arange(20)) and np.array_equal(y[:, 0], 0.5 + np.arange(30))
    o.residual()
    assert np.array_equal(np.fromfile(str(tmp_path / "Re=50_N=20_M=30_u.bin")).view(np.uint64), bits(o.macro()["u"]))


def test_steady_state_matches_the_reference_program(lbm):
    """Defaults of the 2D program; the real reference (3 ranks) gives u in [-0.02294031719,
    0.09694387877] and u(i=50, j=297..299) = 0.08339286, 0.09005163, 0.09670737."""
    N, M = 100, 300
    with lbm.Cavity2D(N=N, M=M, Re=100.0) as sim:
        t, df = lbm.run_to_convergence_2d(sim, THRESH=2e-9)
        assert t == 94000 and df < 2e-9
        u = sim.macro()["u"].reshape(M, N)
    assert abs(u.min() - (-0.02294031719)) < 5e-9 and abs(u.max() - 0.09694387877) < 5e-9
    assert np.allclose(u[297:300, 50], [0.08339286, 0.09005163, 0.09670737], atol=1e-8)


def test_short_horizon_parity_at_config5_size(lbm, oracle):
    """8192^2 (BASELINE config 5), Re = 10000: 3 iterations against the synchronous oracle, bit-exact
    (compared through u, v, rho and the residual to keep host traffic small)."""
    cfg = dict(N=8192, M=8192, Re=10000.0)
    o = oracle.Oracle2D(scheme="sync", **cfg)
    o.step(3)
    r_o = o.residual()
    mo = o.macro()
    with lbm.Cavity2D(**cfg) as sim:
        sim.step(3)
        assert sim.residual() == pytest.approx(r_o, rel=1e-12)
        mg = sim.macro()
    for k in ("u", "v", "rho"):
        assert np.array_equal(bits(mg[k]), bits(mo[k])), k


def test_large_grid_properties(lbm):
    """8192^2 (BASELINE config 5), Re = 10000 (tau = 0.746): batching-invariance and symmetry-free
    sanity (finite, rho deviation small, lid row moving)."""
    N = 8192
    with lbm.Cavity2D(N=N, M=N, Re=10000.0) as a, lbm.Cavity2D(N=N, M=N, Re=10000.0) as b:
        a.step(12)
        for n in (5, 4, 3):
            b.step(n)
        ma, mb = a.macro(), b.macro()
        assert np.array_equal(bits(ma["u"]), bits(mb["u"])) and np.array_equal(bits(ma["rho"]), bits(mb["rho"]))
        assert np.isfinite(ma["rho"]).all() and abs(ma["rho"].mean() - 1.0) < 1e-6
        assert ma["u"].reshape(N, N)[N - 1].min() > 0.0


def test_vmag_is_computed_on_the_device(lbm):
    with lbm.Cavity2D(N=33, M=47, Re=40.0) as sim:
        sim.step(60)
        mv = sim.macro()
    assert np.array_equal(mv["vmag"], np.sqrt(mv["u"] * mv["u"] + mv["v"] * mv["v"])) and mv["vmag"].max() > 0.01
