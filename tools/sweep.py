#!/usr/bin/env python3
"""tools/sweep.py -- kernel tuning sweep on one GPU: plain step launches, CUDA-event timed.
usage: python tools/sweep.py [--grids 256,512] [--blocks 256x1,128x2,64x4,32x8] [--iters 200] [--reps 3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbmpkg import lbm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grids", default="256,512")
ap.add_argument("--blocks", default="0x0,128x1,64x2,64x1,32x4,96x1")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--method", type=int, default=1)
ap.add_argument("--Re", type=int, default=1000)
a = ap.parse_args()
for N in [int(x) for x in a.grids.split(",")]:
    for b in a.blocks.split(","):
        bx, by = [int(x) for x in b.split("x")]
        with lbm.Cavity3D(N=N, Re=a.Re if a.method == 1 else 100, METHOD=a.method, track_residual=False) as sim:
            if bx:
                sim.set_option("block_x", bx)
                sim.set_option("block_y", by)
            sim.step(10)
            sim.sync()
            best = None
            times = []
            for _ in range(a.reps):
                ms = sim.step_timed(a.iters) / a.iters
                times.append(ms)
            best = min(times)
            mlups = N ** 3 / (best * 1e-3) / 1e6
            print(json.dumps({"N": N, "block": b, "ms": [round(t, 4) for t in times], "best_mlups": round(mlups, 1),
                              "GBs": round(mlups * 304e-3, 1)}), flush=True)
