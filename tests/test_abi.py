"""CPU tests: the C-ABI library builds, loads and exports every symbol include/lbm_b200.h declares,
and fails LOUDLY without a GPU (no CPU fallback, no route through the oracle)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points(lbm):
    assert header_symbols() == sorted(lbm.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lbm):
    L = lbm.load_library()
    for s in header_symbols():
        assert hasattr(L, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", lbm.LIB_PATH], capture_output=True, text=True).stdout
    for s in header_symbols():
        assert re.search(r"\bT %s\b" % s, out), s
    assert b"sm_100a" in L.lbm_version()


def test_library_is_sm100a_only(lbm):
    out = subprocess.run(["cuobjdump", "--list-elf", lbm.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_fma_contraction_in_step_kernels(lbm):
    """-fmad=false: the update must keep separate DMUL/DADD.  DFMA may only appear
    inside the IEEE division / sqrt subroutines, never with our own operands; the cheap proxy is
    that the TRT/BGK step kernels contain hundreds of DADD/DMUL and only a handful of DFMA."""
    out = subprocess.run(["cuobjdump", "-sass", lbm.LIB_PATH], capture_output=True, text=True).stdout
    secs = re.split(r"\n\s*Function : ", out)
    sec = [x for x in secs if x.startswith("_ZN5lbm3d6k_stepILi1ELi0ELb1ELb0EEEvNS_2KPE")]
    assert len(sec) == 1
    out = sec[0]
    n_fma = len(re.findall(r"\bDFMA\b", out))
    n_add = len(re.findall(r"\bDADD\b", out))
    n_mul = len(re.findall(r"\bDMUL\b", out))
    assert n_add > 150 and n_mul > 80, (n_add, n_mul)
    assert n_fma <= 24, n_fma  # Newton iterations of 1.0/rho (inline fast path + slow-path subroutine) only


def test_default_params_are_the_reference_defaults(lbm):
    p = lbm.default_params()
    assert (p.dim, p.N, p.Re, p.UMAX, p.METHOD, p.GAMMA, p.MAGIC, p.DIAG) == (3, 50, 100.0, 0.1, 1, 1.0, 0.25, 0)
    assert (p.rank, p.nranks, p.track_residual) == (0, 1, 1)


def test_create_fails_loudly_without_gpu(lbm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lbm.LbmError) as e:
        lbm.Cavity3D(N=12)
    assert e.value.code == lbm.LBM_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_invalid_parameters_are_rejected(lbm):
    L = lbm.load_library()
    p = lbm.default_params()
    h = ctypes.c_void_p()
    for field, bad in (("N", 2), ("METHOD", 7), ("Re", 0.0), ("Re", 12.5), ("GAMMA", 0.0), ("dim", 5), ("nranks", 0)):
        q = lbm.LbmParams.from_buffer_copy(p)
        setattr(q, field, bad)
        assert L.lbm_create(ctypes.byref(q), ctypes.byref(h)) == lbm.LBM_ERR_INVALID, field
        assert L.lbm_last_error()


def test_slab_ranges_partition_the_grid(lbm):
    for n in (12, 29, 50, 512, 1024):
        for P in (1, 2, 3, 4, 8):
            r = [lbm.slab_range(n, k, P) for k in range(P)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(P - 1))
            assert all(b > a for a, b in r)


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "cfd-codes_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".c", ".cpp")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("oracle/_ref", "").replace("the oracle", "") or fn.endswith(".md"), \
                    os.path.join(dirpath, fn)
    out = subprocess.run(["ldd", os.path.join(pkg, "liblbm_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_fast_formatter_equals_printf(tmp_path):
    """host/fmt10.h (used by the Tecplot writer) must produce printf("%.10f") byte for byte."""
    exe = str(tmp_path / "fmt10_check")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "cfd-codes_b200", "host"), "-o", exe,
                           os.path.join(ROOT, "tests", "fmt10_check.c"), "-lm"])
    r = subprocess.run([exe, "1500000"], capture_output=True, text=True)
    assert r.returncode == 0 and "fmt10 ok" in r.stdout, r.stdout[-500:]


def test_step_kernel_resource_budget(lbm):
    """Performance invariants of the dominant kernel that can be checked without a GPU: the TRT step
    kernel must fit 5 blocks of 128 threads per SM (<= 102 registers, profiles/r01_ab_occupancy.log),
    must not spill (no local-memory stack), and uses no shared memory of its own (cuobjdump counts the
    1 KB per block that CUDA 12.9 reserves on sm_100 once any kernel of the module uses cp.async)."""
    out = subprocess.run(["cuobjdump", "-res-usage", lbm.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(r"Function _ZN5lbm3d6k_stepILi1ELi0ELb1ELb0EEEvNS_2KPE:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    assert m, out[:500]
    reg, stack, shared, local = (int(x) for x in m.groups())
    assert reg <= 102 and stack == 0 and shared in (0, 1024) and local == 0, (reg, stack, shared, local)
    # the D2Q9 kernel: 4 blocks of 256 threads per SM
    m = re.search(r"Function _ZN5lbm2d7k2_stepENS_3KP2E:\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m and int(m.group(1)) <= 64 and int(m.group(2)) == 0


def test_step_kernel_memory_instructions(lbm):
    """One load and one store per population and cell: exactly 19 LDG.E.64 on the fast path plus the
    stale-f14 side load, 19 STG.E.64 plus the side store, and no 32-bit or local-memory traffic."""
    out = subprocess.run(["cuobjdump", "-sass", lbm.LIB_PATH], capture_output=True, text=True).stdout
    sec = [x for x in re.split(r"\n\s*Function : ", out) if x.startswith("_ZN5lbm3d6k_stepILi1ELi0ELb1ELb0EEEvNS_2KPE")][0]
    assert len(re.findall(r"\bLDG\.E\.64\b", sec)) == 20 and len(re.findall(r"\bSTG\.E\.64\b", sec)) == 20
    assert not re.findall(r"\b(LDL|STL)\b", sec)


def _sass_section(lbm, prefix):
    out = subprocess.run(["cuobjdump", "-sass", lbm.LIB_PATH], capture_output=True, text=True).stdout
    sec = [x for x in re.split(r"\n\s*Function : ", out) if x.startswith(prefix)]
    assert len(sec) == 1, prefix
    return sec[0]


def test_pair_kernel_memory_instructions_and_no_fma(lbm):
    """k_step_pair (two iterations per pass), default variant (32 x 10 threads, two blocks per SM, direct
    pulls): one load and one store per population for TWO updates of a cell -- 19 LDG + the two stale-f14
    side loads, 19 STG + the two side stores -- the intermediate time level goes through shared memory
    (16 STS.64 / 16 LDS.64, two barriers per plane), and -fmad=false holds (DFMA only inside 1.0/rho)."""
    sec = _sass_section(lbm, "_ZN5lbm3d11k_step_pairILi1ELb1ELi10ELi2ELi3E")
    n = {k: len(re.findall(r"\b%s\b" % k, sec)) for k in (r"LDG\.E\.64", r"STG\.E\.64", r"LDS\.64", r"STS\.64", r"BAR\.SYNC", "DFMA", "DADD", "DMUL")}
    assert n[r"LDG\.E\.64"] == 21 and n[r"STG\.E\.64"] == 21, n
    assert n[r"LDS\.64"] == 16 and n[r"STS\.64"] == 16 and n[r"BAR\.SYNC"] == 2, n
    assert n["DFMA"] <= 24 and n["DADD"] > 300 and n["DMUL"] > 160, n
    out = subprocess.run(["cuobjdump", "-res-usage", lbm.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(r"Function _ZN5lbm3d11k_step_pairILi1ELb1ELi10ELi2ELi3EEEvNS_2KPE14CUtensorMap_st:\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m and int(m.group(1)) <= 102 and int(m.group(2)) <= 16, m and m.groups()   # 2 blocks x 320 threads per SM


def test_tma_variants_use_the_tma_unit(lbm):
    """The TMA-staged variants pull S(t) with tensor copies: 19 UTMALDG.4D at the prologue + 19 in the
    loop, completion on an mbarrier (SYNCS), and no per-thread global loads except the side buffer."""
    for prefix in ("_ZN5lbm3d11k_step_pairILi1ELb1ELi16ELi1ELi4E", "_ZN5lbm3d11k_step_pairILi1ELb1ELi8ELi2ELi4E"):
        sec = _sass_section(lbm, prefix)
        assert len(re.findall(r"\bUTMALDG\.4D\b", sec)) == 38, prefix
        assert len(re.findall(r"\bSYNCS\.PHASECHK\.TRANS64\.TRYWAIT\b", sec)) >= 1
        assert len(re.findall(r"\bLDG\.E\.64\b", sec)) == 2 and len(re.findall(r"\bSTG\.E\.64\b", sec)) == 21
