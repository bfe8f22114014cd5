#!/usr/bin/env python3
"""tools/small_grid_sweep.py [N] -- launch-bound grids: device time per iteration of the device-side
convergence loop (one conditional CUDA graph, 500-iteration blocks) for different block shapes and the
two-cells-per-thread kernel.  The reference's default configuration is N = 50."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbmpkg import lbm  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
for opts in ({}, {"block_x": 64, "block_y": 1}, {"block_x": 64, "block_y": 2}, {"block_x": 32, "block_y": 2}, {"block_x": 32, "block_y": 4},
             {"block_x": 32, "block_y": 1}, {"vec2": 1}, {"vec2": 1, "block_x": 32, "block_y": 2}, {"vec2": 1, "block_x": 32, "block_y": 4}):
    with lbm.Cavity3D(N=N, Re=100) as sim:
        for k, v in opts.items():
            sim.set_option(k, v)
        sim.step(1)
        df0 = 1.0 / sim.residual()
        sim.converge(500, 0.0, 4, df0)          # builds the graph, warms up
        t0 = time.perf_counter()
        done, df, recs = sim.converge(500, 0.0, 40, df0)
        dt = time.perf_counter() - t0
    print(json.dumps({"N": N, "opts": opts, "us_per_iteration": round(dt / (done * 500) * 1e6, 3), "blocks": done, "df": df}), flush=True)
