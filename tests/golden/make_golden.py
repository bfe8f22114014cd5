#!/usr/bin/env python3
"""tests/golden/make_golden.py -- regenerates the committed fixtures (run in the BUILD container,
where /root/reference exists; the GPU box only reads the committed outputs).

1. solution_n29_re100.npz  <- the reference's only shipped result for this path,
   `3D .../example output/Solution_n=29Re=100_DIAG=0_2RT_M=0.250.dat` (header + 29^3 rows of
   x,y,z,u,v,w,vmag printed with %.10f), parsed to float64 columns.  The header line is kept verbatim.
2. ref_hashes.json  <- sha256 of the raw lattice (NQ doubles, LU4 order) produced by the REAL
   reference (oracle/_ref binaries, see oracle/build_ref.py) after n iterations, per configuration.
   A hash is enough because the parity target is bit-exact.
3. ref_f_N12_Re100_300.npy  <- one small raw lattice, so a mismatch can be localised.
4. solution_n50_re100_default.npz  <- the reference's DEFAULT configuration (BASELINE config 1: N=50,
   Re=100, TRT, THRESH=1e-12) run to convergence with `oracle/_ref/lbm3d_ref_N50_... main` (the
   reference's own main(), 131 000 iterations, ~25 min on 8 cores): the printed values of its
   Solution_*.dat as int32 = round(value * 1e10), the iteration count and the printed residuals.
   Only rebuilt with --default-run (slow).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

DAT = ("/root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/example output/"
       "Solution_n=29Re=100_DIAG=0_2RT_M=0.250.dat")

HASH_CASES = [
    (dict(N=12, Re=100), [1, 2, 3, 10, 300, 1000]),
    (dict(N=12, Re=20, METHOD=0), [1, 2, 10, 300, 800]),
    (dict(N=12, Re=100, DIAG=1), [1, 10, 500]),
    (dict(N=12, Re=100, GAMMA="0.5"), [1, 10, 500]),
    (dict(N=13, Re=20, METHOD=0, DIAG=1), [1, 10, 200]),
    (dict(N=21, Re=400), [1, 10, 600]),
    (dict(N=24, Re=100), [200]),
    (dict(N=29, Re=100), [100]),
    (dict(N=32, Re=65, METHOD=0), [100]),
    (dict(N=50, Re=100), [20, 501]),
    (dict(N=64, Re=253), [30]),
    (dict(N=3, Re=10), [1, 2, 3, 50]),
    (dict(N=4, Re=10, METHOD=0), [1, 2, 3, 50]),
    (dict(N=5, Re=30, DIAG=1, GAMMA="2.0", UMAX="0.05"), [1, 2, 3, 80]),
    (dict(N=16, Re=200, UMAX="0.2", GAMMA="0.8"), [1, 5, 120]),
]


def default_run():
    import subprocess
    import tempfile
    exe = O.ref_binary(N=50, Re=100)
    with tempfile.TemporaryDirectory() as tmp:
        log = subprocess.run([exe, "main"], cwd=tmp, capture_output=True, text=True, check=True).stdout
        dat = os.path.join(tmp, "Solution_n=50Re=100_DIAG=0_2RT_M=0.250.dat")
        md5 = hashlib.md5(open(dat, "rb").read()).hexdigest()
        with open(dat) as fh:
            header = fh.readline().rstrip("\n")
            data = np.loadtxt(fh, delimiter=",")
    its = [l for l in log.splitlines() if l.startswith("Iteration")]
    dfs = [l.split("\t")[1] for l in log.splitlines() if l.startswith("df")]
    q = lambda a: np.round(a * 1e10).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "solution_n50_re100_default.npz"), header=np.array(header), u_e10=q(data[:, 3]),
                        v_e10=q(data[:, 4]), w_e10=q(data[:, 5]), vmag_e10=q(data[:, 6]),
                        iterations=np.array(int(its[-1].split()[1].rstrip(":"))), df_final=np.array(dfs[-1]),
                        df0=np.array(dfs[0]), dat_md5=np.array(md5))


def main():
    if "--default-run" in sys.argv:
        default_run()
        return
    with open(DAT) as fh:
        header = fh.readline().rstrip("\n")
        data = np.loadtxt(fh, delimiter=",")
    assert data.shape == (29 ** 3, 7)
    np.savez_compressed(os.path.join(HERE, "solution_n29_re100.npz"), header=np.array(header), x=data[:, 0],
                        y=data[:, 1], z=data[:, 2], u=data[:, 3], v=data[:, 4], w=data[:, 5], vmag=data[:, 6])
    hashes = {}
    for cfg, steps in HASH_CASES:
        tag = O.ref_tag(**cfg)
        hashes[tag] = {}
        for n in steps:
            f, d, rc, _ = O.ref_steps(n, threads=8, **cfg)
            assert rc == 0
            hashes[tag][str(n)] = {"sha256": hashlib.sha256(f.tobytes()).hexdigest(), "last_diff": d}
            if cfg == dict(N=12, Re=100) and n == 300:
                np.save(os.path.join(HERE, "ref_f_N12_Re100_300.npy"), f)
        print(tag, "done")
    # the one unstable configuration: the reference prints its message and exits 1 (main.c:687-692)
    f, d, rc, out = O.ref_steps(400, threads=8, N=12, Re=100, METHOD=0)
    hashes["unstable_" + O.ref_tag(N=12, Re=100, METHOD=0)] = {"returncode": rc, "stdout_tail": out.strip().splitlines()[-1]}
    with open(os.path.join(HERE, "ref_hashes.json"), "w") as fh:
        json.dump(hashes, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
