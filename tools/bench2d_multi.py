#!/usr/bin/env python3
"""tools/bench2d_multi.py -- launched with torchrun: MLUPS of the D2Q9 kernel over y-row slabs (NCCL halo of
rows {2,5,6} up / {4,7,8} down, the reference's own exchange, 2D main.cpp:208-270)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lbmpkg import lbm  # noqa: E402
import cfd_codes_b200.distributed as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8192)
ap.add_argument("--iters", type=int, default=1000)
a = ap.parse_args()
rank, world, local = D.env_rank()
dist = D.init_process_group()
nid = D.broadcast_nccl_id() if world > 1 else None
with lbm.Cavity2D(N=a.N, M=a.N, Re=10000.0, rank=rank, nranks=world, device=local, nccl_id=nid) as sim:
    sim.step(200)
    r0 = sim.residual()
    dist.barrier()
    ms = D.allreduce_max(sim.step_timed(a.iters))
    r1 = sim.residual()
if rank == 0:
    mlups = a.N * a.N * a.iters / (ms * 1e-3) / 1e6
    print(json.dumps({"workload": "D2Q9 %d^2" % a.N, "n_gpus": world, "iters": a.iters, "ms_per_iter": ms / a.iters,
                      "mlups": mlups, "GBs_per_gpu": mlups * 144e-3 / world, "residual": r1}), flush=True)
dist.barrier()
dist.destroy_process_group()
