#!/usr/bin/env python3
"""tools/sweep2d.py -- D2Q9 kernel timing on one GPU (plain steps, CUDA events)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbmpkg import lbm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grids", default="8192")
ap.add_argument("--blocks", default="0,64,128,256")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
for N in [int(x) for x in a.grids.split(",")]:
    for bx in [int(x) for x in a.blocks.split(",")]:
        with lbm.Cavity2D(N=N, M=N, Re=10000.0) as sim:
            if bx:
                sim.set_option("block_x", bx)
            sim.step(10)
            sim.sync()
            times = [sim.step_timed(a.iters) / a.iters for _ in range(a.reps)]
            best = min(times)
            mlups = N * N / (best * 1e-3) / 1e6
            print(json.dumps({"N": N, "block_x": bx, "ms": [round(t, 4) for t in times], "best_mlups": round(mlups, 1),
                              "GBs": round(mlups * 144e-3, 1)}), flush=True)
