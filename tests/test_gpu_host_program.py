"""GPU tests of the C host programs (cfd-codes_b200/host/main.c, main2d.c): the drop-in
replacements of the reference's two executables, driven exactly as a user would."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EXE3 = os.path.join(ROOT, "cfd-codes_b200", "lbm3d_cavity")
EXE2 = os.path.join(ROOT, "cfd-codes_b200", "lbm2d_cavity")


def test_lbm3d_cavity_reproduces_the_golden_run(tmp_path):
    """`lbm3d_cavity --N 29` = the reference built with N 29: same stdout blocks (SURVEY App. C:
    'Iteration 0: df 1.909e+00', converges at 'Iteration 43500: df/df0 8.340e-13') and the same
    Solution_*.dat as the reference's shipped example output (numeric compare; zero signs aside)."""
    r = subprocess.run([EXE3, "--N", "29"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    out = r.stdout
    assert out.startswith("Threads used:\t1\nRe =\t100\nRe_g =\t")
    assert "N =\t29\n" in out and "tau =\t5.900e-01\n" in out and "nu =\t3.000e-02\n" in out and "M =\t1.732e-01\n" in out
    assert "\nIteration 0:\ndf:\t1.909e+00\n" in out
    its = re.findall(r"\nIteration (\d+):\ndf/df0:\t(\S+)\nTime:\t\S+ s\nLUPS:\t\S+ s\n", out)
    assert its[0][0] == "500" and its[-1] == ("43500", "8.340e-13") and len(its) == 87
    path = tmp_path / "Solution_n=29Re=100_DIAG=0_2RT_M=0.250.dat"
    g = np.load(os.path.join(GOLD, "solution_n29_re100.npz"))
    with open(path) as fh:
        header = fh.readline().rstrip("\n")
        data = np.loadtxt(fh, delimiter=",")
    assert header == str(g["header"])
    assert data.shape == (29 ** 3, 7)
    for col, name in enumerate(("x", "y", "z", "u", "v", "w", "vmag")):
        assert np.array_equal(data[:, col] + 0.0, g[name] + 0.0), name   # identical at the 10 printed decimals


def test_lbm3d_cavity_reports_instability_like_the_reference(tmp_path):
    r = subprocess.run([EXE3, "--N", "12", "--METHOD", "0", "--SAVETXT", "0"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 1
    assert r.stdout.rstrip().endswith("Error: negative feq.  Therefore, unstable.")


@pytest.mark.parametrize("maxiter,skip,steps_printed", [(41, 500, 40), (1001, 500, 1000), (31, 10, 30)])
def test_lbm3d_cavity_maxiter_exit_prints_the_swapped_buffer(tmp_path, oracle, maxiter, skip, steps_printed):
    """When the loop ends at MAXITER without converging, the reference has already swapped f1/f2, so
    its datwrite prints the state of iteration MAXITER-2 (main.c:104-127) = `steps_printed` updates;
    also at a MAXITER that falls right after a residual check."""
    r = subprocess.run([EXE3, "--N", "12", "--MAXITER", str(maxiter), "--skip", str(skip), "--THRESH", "1e-30"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    o = oracle.Oracle3D(N=12, Re=100, threads=2)
    o.step(steps_printed)
    mv = o.macrovars()
    with open(tmp_path / "Solution_n=12Re=100_DIAG=0_2RT_M=0.250.dat") as fh:
        fh.readline()
        data = np.loadtxt(fh, delimiter=",")
    # file order: i outer, j, k inner (main.c:373-376); oracle arrays are LU3 = [k][j][i]
    for col, name in ((3, "u1"), (4, "u2"), (5, "u3"), (6, "vmag")):
        want = mv[name].reshape(12, 12, 12).transpose(2, 1, 0).ravel()
        assert np.array_equal(np.round(data[:, col] * 1e10), np.round(want * 1e10)), name


def test_lbm2d_cavity_outputs(tmp_path):
    r = subprocess.run([EXE2, "--N", "40", "--M", "60", "--Re", "50", "--THRESH", "1e-6"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    out = r.stdout
    assert out.startswith("MPI LBM Simulation of Lid Driven Cavity\nDomain Size: 40 x 60\nRe\t= 50\nnu\t= 0.08\ntau\t= 0.74\nUmax\t= 0.1\n")
    assert "Rank 0: 60 nodes from y =   0 to y =  59 (BOTTOM)" in out
    rows = re.findall(r"^(\d\.\d{3}e[+-]\d+)\t(\d\.\d{3}e[+-]\d+)\t", out, flags=re.M)
    assert rows[0][0] == "0.000e+00" and rows[0][1] == "1.000e+00" and float(rows[-1][1]) < 1e-6
    u = np.fromfile(tmp_path / "Re=50_N=40_M=60_u.bin").reshape(60, 40)
    x = np.fromfile(tmp_path / "Re=50_N=40_M=60_x.bin").reshape(60, 40)
    rho = np.fromfile(tmp_path / "Re=50_N=40_M=60_rho.bin")
    assert np.array_equal(x[0], 0.5 + np.arange(40)) and u[59].min() > 0.0 and u[59].mean() > 0.05 and abs(rho.mean() - 1.0) < 1e-3


def test_default_configuration_drop_in(tmp_path):
    """BASELINE config 1: `lbm3d_cavity` with NO flags = the reference's `./a.out` (N=50, Re=100, TRT,
    THRESH=1e-12).  The reference needs 131 000 iterations / 22 minutes on 8 cores; its
    result was recorded by tests/golden/make_golden.py --default-run.  Same iteration count, same
    printed residuals, every printed digit of the 125 000 x 4 output values identical."""
    g = np.load(os.path.join(GOLD, "solution_n50_re100_default.npz"))
    r = subprocess.run([EXE3], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    out = r.stdout
    for line in ("Re =\t100", "Re_g =\t1.961e+00", "U =\t1.000e-01", "N =\t50", "tau =\t6.530e-01", "nu =\t5.100e-02",
                 "M =\t1.732e-01"):                     # the banner of SURVEY App. C
        assert line + "\n" in out, line
    assert "\nIteration 0:\ndf:\t%s\n" % str(g["df0"]) in out
    its = re.findall(r"\nIteration (\d+):\ndf/df0:\t(\S+)\n", out)
    assert its[-1] == (str(int(g["iterations"])), str(g["df_final"]))
    with open(tmp_path / "Solution_n=50Re=100_DIAG=0_2RT_M=0.250.dat") as fh:
        assert fh.readline().rstrip("\n") == str(g["header"])
        data = np.loadtxt(fh, delimiter=",")
    for col, name in ((3, "u"), (4, "v"), (5, "w"), (6, "vmag")):
        assert np.array_equal(np.round(data[:, col] * 1e10).astype(np.int32), g[name + "_e10"]), name


def test_lbm3d_cavity_multi_gpu_matches_single_gpu(tmp_path):
    """`--gpus 2` (one host thread per GPU, NCCL halo) and `--gpus 2 --halo 1` (fused direct-NVLink
    halo, same-process peer access) must write the same Solution_*.dat as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = []
    for extra in ([], ["--gpus", "2"], ["--gpus", "2", "--halo", "1"]):
        d = tmp_path / ("run%d" % len(outs))
        d.mkdir()
        r = subprocess.run([EXE3, "--N", "30", "--Re", "60", "--THRESH", "1e-6"] + extra, cwd=d, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        its = re.findall(r"\nIteration (\d+):\ndf/df0:\t(\S+)\n", r.stdout)
        outs.append((its, open(d / "Solution_n=30Re=60_DIAG=0_2RT_M=0.250.dat").read()))
    assert outs[0][0] == outs[1][0] == outs[2][0] and len(outs[0][0]) >= 2
    assert outs[0][1] == outs[1][1] == outs[2][1]
