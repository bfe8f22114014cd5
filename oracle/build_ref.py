#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE: builds oracle/_ref/ from the real reference.

Compiles the reference's own 3D solver
(`/root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/main.c`) from where it lies,
once per parameter set, into small stepping binaries under `oracle/_ref/` (git-ignored; NOT
gpurun-ignored, so the binaries travel to the GPU box where /root/reference does not exist).

The reference's parameters are `#define`s (3D:7-15) that `-D` cannot override, so each build
pipes the file through a substitution of exactly those lines into a temporary directory, compiles
`ref_harness.c` against it (`#define main ref_main` / `#include`), and deletes the temporary copy.
No reference source is ever written into the repository.

Flags: the reference CI line (`gcc main.c -lm -fopenmp -O3`, .github/workflows/main.yml:118) plus
`-ffp-contract=off` (FMA contraction changes the bits).

Usage:  python oracle/build_ref.py            # build every config in CONFIGS (skips up-to-date ones)
        python oracle/build_ref.py --list
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_MAIN = "/root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/main.c"

# (N, Re, METHOD, GAMMA, DIAG, UMAX) -- every configuration the tests / bench need.
DEFAULT = dict(N=50, Re=100, METHOD=1, GAMMA="1.0", DIAG=0, UMAX="0.1")
CONFIGS = [
    dict(N=12, Re=100),                      # TRT small, SURVEY App. B
    dict(N=12, Re=20, METHOD=0),             # BGK (stable tau)
    dict(N=12, Re=100, DIAG=1),              # diagonal lid
    dict(N=12, Re=100, GAMMA="0.5"),         # preconditioned
    dict(N=13, Re=20, METHOD=0, DIAG=1),     # odd N, BGK + DIAG
    dict(N=12, Re=100, METHOD=0),            # BGK unstable (negative feq -> exit 1)
    dict(N=21, Re=400),                      # odd N, low tau
    dict(N=24, Re=100),                      # md5 check constant of SURVEY App. C
    dict(N=29, Re=100),                      # golden-file config
    dict(N=32, Re=65, METHOD=0),             # BGK at the default tau
    dict(N=50, Re=100),                      # reference defaults (config 1)
    dict(N=64, Re=253),                      # same tau as 256^3/Re=1000
    dict(N=256, Re=1000),                    # bench config 2 (CPU baseline, kind "reference")
    dict(N=3, Re=10),                        # smallest legal grid: one fluid cell, every rule class has one cell
    dict(N=4, Re=10, METHOD=0),              # no face-interior ... 2x2 faces, BGK
    dict(N=5, Re=30, DIAG=1, GAMMA="2.0", UMAX="0.05"),   # GAMMA > 1, other lid speed
    dict(N=16, Re=200, UMAX="0.2", GAMMA="0.8"),
]


def tag(cfg):
    c = dict(DEFAULT, **cfg)
    return "N{N}_Re{Re}_m{METHOD}_g{GAMMA}_d{DIAG}_u{UMAX}".format(**c)


def binary_path(cfg):
    return os.path.join(OUT, "lbm3d_ref_" + tag(cfg))


def patch(src, cfg):
    c = dict(DEFAULT, **cfg)
    for name in ("Re", "N", "UMAX", "METHOD", "GAMMA", "DIAG"):
        pat = re.compile(r"^(#define\s+%s\s+)(\S+)" % name, re.M)
        src, n = pat.subn(lambda m: m.group(1) + str(c[name]), src, count=1)
        if n != 1:
            raise RuntimeError("could not patch #define %s" % name)
    return src


def build(cfg, force=False):
    out = binary_path(cfg)
    harness = os.path.join(HERE, "ref_harness.c")
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) >= os.path.getmtime(harness)
            and os.path.getmtime(out) >= os.path.getmtime(__file__)):
        return out
    os.makedirs(OUT, exist_ok=True)
    with open(REF_MAIN) as f:
        src = patch(f.read(), cfg)
    with tempfile.TemporaryDirectory() as tmp:
        psrc = os.path.join(tmp, "ref_patched.c")
        with open(psrc, "w") as f:
            f.write(src)
        cmd = ["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-w",
               "-DREF_SRC=\"%s\"" % psrc, harness, "-lm", "-o", out]
        subprocess.check_call(cmd)
    return out


# ---- the 2D program (D2Q9), built over oracle/mpi_shim/mpi.h ---------------------------------------
REF_MAIN_2D = "/root/reference/2D Lid-driven Cavity with Lattice Boltzmann Method/main.cpp"
# N, M, Re, MAXITER (iterations to run), prntInt (macroVars cadence), THRESH
CONFIGS_2D = [
    dict(N=100, M=300, Re="100", MAXITER="1E6", prntInt="1e3", THRESH="2E-9"),   # the reference's defaults (to convergence)
    dict(N=24, M=40, Re="40", MAXITER="200", prntInt="1", THRESH="0"),           # 200 iterations, state after the last one
    dict(N=17, M=23, Re="25", MAXITER="120", prntInt="1", THRESH="0"),           # odd sizes
    # BASELINE config 5 grid for bench.py's cpu_baseline: 11 iterations, MLUPS printed at t = 5, 10; no output files
    dict(N=8192, M=8192, Re="10000", MAXITER="11", prntInt="5", THRESH="0", SAVE="false"),
]


def tag2d(c):
    return "N{N}_M{M}_Re{Re}_it{MAXITER}_p{prntInt}_th{THRESH}".format(**c)


def binary_path_2d(c):
    return os.path.join(OUT, "lbm2d_ref_" + tag2d(c))


def build2d(cfg, force=False):
    out = binary_path_2d(cfg)
    shim = os.path.join(HERE, "mpi_shim")
    if (not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(os.path.join(shim, "mpi.h"))
            and os.path.getmtime(out) >= os.path.getmtime(__file__)):
        return out
    os.makedirs(OUT, exist_ok=True)
    with open(REF_MAIN_2D) as f:
        src = f.read()
    subs = [(r"(const double Re = )[^;]+", cfg["Re"]), (r"(const int N = )[^;]+", str(cfg["N"])),
            (r"(const int M = )[^;]+", str(cfg["M"])), (r"(const double THRESH = )[^;]+", cfg["THRESH"]),
            (r"(const double MAXITER = )[^;]+", cfg["MAXITER"]), (r"(const int prntInt = )[^;]+", cfg["prntInt"])]
    if "SAVE" in cfg:   # SAVE is one of the reference's own user parameters (2D:16)
        subs.append((r"(const bool SAVE = )[^;]+", cfg["SAVE"]))
    for pat, val in subs:
        src, n = re.subn(pat, lambda m: m.group(1) + val, src, count=1)
        if n != 1:
            raise RuntimeError("could not patch %s" % pat)
    with tempfile.TemporaryDirectory() as tmp:
        psrc = os.path.join(tmp, "ref2d_patched.cpp")
        with open(psrc, "w") as f:
            f.write(src)
        # <mpi.h> resolves to the shim; flags as the reference CI's g++ line + -ffp-contract=off
        subprocess.check_call(["g++", "-O3", "-ffp-contract=off", "-w", "-I", shim, psrc, "-o", out])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if a.list:
        for c in CONFIGS:
            print(binary_path(c))
        for c in CONFIGS_2D:
            print(binary_path_2d(c))
        return 0
    if not os.path.exists(REF_MAIN):
        print("reference not present (%s); keeping prebuilt oracle/_ref" % REF_MAIN)
        return 0
    for c in CONFIGS:
        print("built", build(c, a.force))
    for c in CONFIGS_2D:
        print("built", build2d(c, a.force))
    return 0


if __name__ == "__main__":
    sys.exit(main())
