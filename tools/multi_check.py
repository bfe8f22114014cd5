#!/usr/bin/env python3
"""tools/multi_check.py -- launched with torchrun (one rank per GPU): z-slab run vs the CPU oracle.
Prints one JSON line on rank 0; exit code 1 on mismatch.  Used by tests/test_gpu_multi.py."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lbmpkg import lbm  # noqa: E402
import cfd_codes_b200.distributed as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=32)
ap.add_argument("--Re", type=int, default=100)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--method", type=int, default=1)
ap.add_argument("--diag", type=int, default=0)
ap.add_argument("--dim", type=int, default=3)
ap.add_argument("--halo", default="nccl", choices=["nccl", "p2p"])
ap.add_argument("--track", type=int, default=1)
ap.add_argument("--M", type=int, default=40)
ap.add_argument("--pair", type=int, default=-1, help="two-iterations-per-pass kernel (k_step_pair): -1 library default, 0 off, 1 on")
ap.add_argument("--pair-variant", type=int, default=-1)
ap.add_argument("--pair-lz", type=int, default=0)
a = ap.parse_args()
rank, world, local = D.env_rank()
dist = D.init_process_group()
nid = D.broadcast_nccl_id() if world > 1 else None
if a.dim == 2:
    # D2Q9: row slabs; compare dist / u / v / rho with the synchronous oracle
    cfg = dict(N=a.N, M=a.M, Re=float(a.Re))
    sim = lbm.Cavity2D(rank=rank, nranks=world, device=local, nccl_id=nid, **cfg)
    sim.set_option("host_local_layout", 1)
    info = sim.info()
    res = []
    for n in (1, 1, a.steps - 2):
        sim.step(n)
        res.append(sim.residual())
    nloc = info["cells_local"]
    d = np.zeros(9 * nloc)
    sim.dist(out=d)
    mv = {k: np.zeros(nloc) for k in ("u", "v", "rho")}
    sim.macro(out=mv)
    # local dist is [rows][9][N] (reference layout, local rows) -> gather as 1 field of rows*9*N
    full = D.gather_slabs(d, a.M, 9 * a.N, 1)
    u = D.gather_slabs(mv["u"], a.M, a.N, 1)
    sim.close()
    ok = True
    if rank == 0:
        from oracle import oracle as O
        o = O.Oracle2D(scheme="sync", **cfg)
        ores = []
        for n in (1, 1, a.steps - 2):
            o.step(n)
            ores.append(o.residual())
        same = np.array_equal(full.reshape(-1).view(np.uint64), o.dist().view(np.uint64))
        u_same = np.array_equal(u.reshape(-1).view(np.uint64), o.macro()["u"].view(np.uint64))
        res_ok = all(abs(x - y) <= 1e-12 * abs(y) for x, y in zip(res, ores))
        ok = same and u_same and res_ok
        print(json.dumps({"world": world, "dim": 2, "N": a.N, "M": a.M, "bit_equal": bool(same), "rho_equal": bool(u_same),
                          "residual_ok": bool(res_ok), "residuals": res, "oracle_residuals": ores}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

cfg = dict(N=a.N, Re=a.Re, METHOD=a.method, DIAG=a.diag)
sim = lbm.Cavity3D(rank=rank, nranks=world, device=local, nccl_id=nid, track_residual=bool(a.track), **cfg)
sim.set_option("host_local_layout", 1)
if a.halo == "p2p" and world > 1:
    D.connect_direct_halo(sim)
if a.pair >= 0:
    sim.set_option("pair", a.pair)
    sim.set_option("pair_variant", a.pair_variant)
    sim.set_option("pair_lz", a.pair_lz)
info = sim.info()
res = []
for n in (1, 1, 1, a.steps - 3):
    sim.step(n)
    res.append(sim.residual() if a.track else 0.0)
loc = np.zeros(19 * info["cells_local"])
sim.f(out=loc)
mv = {k: np.zeros(info["cells_local"]) for k in ("u1", "u2", "u3", "rho", "vmag")}
sim.macrovars(out=mv)
full = D.gather_slabs(loc, a.N, a.N * a.N, 19)
rho = D.gather_slabs(mv["rho"], a.N, a.N * a.N, 1)
sim.close()
ok = True
if rank == 0:
    from oracle import oracle as O
    o = O.Oracle3D(threads=8, **cfg)
    ores = []
    for n in (1, 1, 1, a.steps - 3):
        o.step(n)
        ores.append(o.last_diff)
    g = o.f()
    same = np.array_equal(full.reshape(-1).view(np.uint64), g.view(np.uint64))
    rho_same = np.array_equal(rho.reshape(-1).view(np.uint64), o.macrovars()["rho"].view(np.uint64))
    res_ok = (not a.track) or all(abs(x - y) <= 1e-12 * abs(y) for x, y in zip(res, ores))
    ok = same and rho_same and res_ok
    print(json.dumps({"world": world, "halo": a.halo, "pair": a.pair, "N": a.N, "steps": a.steps, "bit_equal": bool(same), "rho_equal": bool(rho_same),
                      "residual_ok": bool(res_ok), "residuals": res, "oracle_residuals": ores,
                      "sha256": hashlib.sha256(full.tobytes()).hexdigest()}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
