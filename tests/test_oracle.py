"""CPU tests (-m "not gpu"): pin the oracle (oracle/lbm3d_oracle.c) against the REAL reference.

Pins, in order of strength:
  1. sha256 of the raw lattice after n iterations == the hash recorded from oracle/_ref (the
     reference's own main.c compiled, tests/golden/make_golden.py) -- bit-exact, 11 configurations;
  2. live comparison with the oracle/_ref binaries when they are present;
  3. the reference's only shipped result, example output/Solution_n=29Re=100_DIAG=0_2RT_M=0.250.dat
     (tests/golden/solution_n29_re100.npz), reproduced to all 10 printed decimals.
"""
import hashlib
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HASHES = json.load(open(os.path.join(GOLD, "ref_hashes.json")))


def parse_tag(tag):
    parts = tag.split("_")
    return dict(N=int(parts[0][1:]), Re=int(parts[1][2:]), METHOD=int(parts[2][1:]), GAMMA=float(parts[3][1:]),
                DIAG=int(parts[4][1:]), UMAX=float(parts[5][1:]))


CASES = [(tag, sorted(int(n) for n in v)) for tag, v in HASHES.items() if not tag.startswith("unstable_")]


@pytest.mark.parametrize("tag,steps", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_hashes(oracle, tag, steps):
    o = oracle.Oracle3D(threads=8, **parse_tag(tag))
    done = 0
    for n in steps:
        o.step(n - done)
        done = n
        rec = HASHES[tag][str(n)]
        assert hashlib.sha256(o.f().tobytes()).hexdigest() == rec["sha256"], "%s after %d iterations" % (tag, n)
        # diff() has no fixed summation order in the reference (OpenMP reduction, main.c:680)
        assert o.last_diff == pytest.approx(rec["last_diff"], rel=1e-12)
    o.close()


def test_oracle_matches_live_reference_binary(oracle):
    cfg = dict(N=12, Re=100)
    if not oracle.ref_available(**cfg):
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    f, d, rc, _ = oracle.ref_steps(37, threads=4, **cfg)
    assert rc == 0
    o = oracle.Oracle3D(threads=3, **cfg)
    o.step(37)
    assert np.array_equal(o.f().view(np.uint64), f.view(np.uint64))
    g = np.load(os.path.join(GOLD, "ref_f_N12_Re100_300.npy"))
    o.step(300 - 37)
    assert np.array_equal(o.f().view(np.uint64), g.view(np.uint64))


def test_derived_constants(oracle):
    # SURVEY Appendix C (values printed by the C build of the reference; hex = exact)
    d = oracle.Oracle3D(N=50, Re=100).derived()
    assert d["NU"] == 0.051000000000000004 and d["TAU"] == 0.653
    assert d["OMEGA"] == float.fromhex("0x1.880968ac7f051p+0")
    assert d["OMEGAm"] == float.fromhex("0x1.dfda5d4e03ebcp-2")
    d = oracle.Oracle3D(N=29, Re=100).derived()
    assert d["OMEGA"] == float.fromhex("0x1.b1e5f75270d05p+0") and d["OMEGAm"] == float.fromhex("0x1.386822b63cbedp-2")
    d = oracle.Oracle3D(N=12, Re=100, DIAG=1).derived()
    assert d["ulid"] == 0.07071067811865475 and d["wlid"] == d["ulid"] and d["vlid"] == 0.0


def test_unstable_bgk_is_detected(oracle):
    rec = HASHES["unstable_N12_Re100_m0_g1.0_d0_u0.1"]
    assert rec["returncode"] == 1 and "negative feq" in rec["stdout_tail"]
    o = oracle.Oracle3D(N=12, Re=100, METHOD=0)
    assert o.step(400) == 1 and o.unstable


def test_buried_slots_stay_at_initial_weight(oracle):
    """2 slots per edge cell and 6 per corner cell are never written."""
    N = 12
    o = oracle.Oracle3D(N=N, Re=100)
    o.step(60)
    f = o.f().reshape(19, N, N, N)  # [q][k][j][i]
    c = oracle.C3
    opp = lambda q: 0 if q == 0 else (q + 1 if q % 2 else q - 1)
    n_buried = 0
    for k in range(N):
        for j in range(N):
            for i in range(N):
                if sum(v in (0, N - 1) for v in (i, j, k)) < 2:
                    continue
                def unknown(q):
                    s = (i - c[q][0], j - c[q][1], k - c[q][2])
                    return any(v < 0 or v >= N for v in s)
                for q in range(1, 19):
                    if unknown(q) and unknown(opp(q)):
                        assert f[q, k, j, i] == oracle.W3[q]
                        n_buried += 1
    assert n_buried == 12 * (N - 2) * 2 + 8 * 6


@pytest.mark.slow
def test_golden_file_n29(oracle):
    """Run the reference's driver loop (main.c:93-124) on the oracle to convergence and compare with
    the shipped example output.  SURVEY 4: converges at iteration 43500 with df/df0 = 8.340e-13."""
    g = np.load(os.path.join(GOLD, "solution_n29_re100.npz"))
    N = 29
    o = oracle.Oracle3D(N=N, Re=100, threads=8)
    o.step(1)
    df0 = o.last_diff
    assert "%.3e" % df0 == "1.909e+00"
    t = 0
    while True:
        o.step(500)
        t += 500
        df = o.last_diff / df0
        if df < 1e-12:
            break
        assert t < 100000
    assert t == 43500 and "%.3e" % df == "8.340e-13"
    mv = o.macrovars()
    # file order: i outer, j, k inner (main.c:373-376); arrays are LU3 = [k][j][i]
    for name, key in (("u", "u1"), ("v", "u2"), ("w", "u3"), ("vmag", "vmag")):
        a = mv[key].reshape(N, N, N).transpose(2, 1, 0).ravel()
        assert np.max(np.abs(a - g[name])) <= 5e-11, name
        # every value equals the golden one at the 10 printed decimals (zero signs aside)
        assert np.array_equal(np.round(a, 10) + 0.0, np.round(g[name], 10) + 0.0), name
    assert np.array_equal(g["x"].reshape(N, N, N)[:, 0, 0], np.arange(N))


def test_product_device_functions_compiled_for_host_match_oracle(oracle, tmp_path):
    """The product's collide<METHOD, G1> and apply_boundary (csrc/lbm3d_kernels.cuh + d3q19_rules.h) are
    `__host__ __device__`; tests/host_device_functions_check.cu compiles that same source for the CPU and
    compares it bit-for-bit with the oracle's single-cell entry points: 26 boundary classes, BGK and
    TRT, DIAG 0/1, GAMMA 1/0.5/2, including far-from-equilibrium cells (negative-feq flag)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "hdf_check")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-fmad=false", "-Xcompiler",
                           "-ffp-contract=off", "-I", os.path.join(root, "cfd-codes_b200", "csrc"),
                           os.path.join(root, "tests", "host_device_functions_check.cu"), "-L", os.path.join(root, "oracle"),
                           "-llbm_oracle", "-Xlinker", "-rpath," + os.path.join(root, "oracle"), "-o", exe])
    r = subprocess.run([exe, "8000"], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-800:]
