#!/usr/bin/env python3
"""tools/time_lib.py <lib.so> [N] -- sustained per-iteration time of plain step launches with the
given build of the library (A/B testing of kernel variants inside ONE gpurun call)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbmpkg import lbm  # noqa: E402

path = os.path.abspath(sys.argv[1])
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
opts = dict(kv.split("=") for kv in sys.argv[3:])
lbm.LIB_PATH = path
lbm._lib = None
with lbm.Cavity3D(N=N, Re=1000, track_residual=False) as sim:
    for k, v in opts.items():
        sim.set_option(k, int(v))
    burst = sim.step_timed(100) / 100
    sim.step(int(2000 * (256.0 / N) ** 3))  # ~1.6 s: reach the power-capped steady state
    sim.sync()
    t = [sim.step_timed(max(20, int(300 * (256.0 / N) ** 3))) / max(20, int(300 * (256.0 / N) ** 3)) for _ in range(3)]
print(json.dumps({"lib": os.path.basename(path), "opts": opts, "N": N, "burst_ms": round(burst, 4), "sustained_ms": [round(x, 4) for x in t],
                  "sustained_mlups": round(N ** 3 / (sum(t) / len(t)) / 1e3, 1)}), flush=True)
