#!/usr/bin/env python3
"""tools/longrun.py -- sustained behaviour: per-block device time of a long run + nvidia-smi clocks/power.
usage: python tools/longrun.py [--grid 256] [--blocks 20] [--block 500] [--method 1]"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbmpkg import lbm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--blocks", type=int, default=20)
ap.add_argument("--block", type=int, default=500)
ap.add_argument("--method", type=int, default=1)
ap.add_argument("--track", type=int, default=1)
a = ap.parse_args()
N = a.grid
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                        "--format=csv,noheader,nounits", "-lms", "250"], stdout=subprocess.PIPE, text=True)
with lbm.Cavity3D(N=N, Re=1000 if a.method == 1 else 100, METHOD=a.method, track_residual=bool(a.track)) as sim:
    sim.step(5)
    sim.sync()
    t0 = time.time()
    for b in range(a.blocks):
        ms = sim.step_timed(a.block)
        r = sim.residual() if a.track else None
        print(json.dumps({"block": b, "t": round(time.time() - t0, 3), "ms_per_iter": round(ms / a.block, 4),
                          "mlups": round(N ** 3 * a.block / ms / 1e3, 1), "residual": r}), flush=True)
smi.terminate()
out = smi.communicate()[0].strip().splitlines()
print("clocks.sm, clocks.mem, power, temp, sw_power_cap, hw_slowdown, sw_thermal")
for l in out[:: max(1, len(out) // 40)]:
    print(l)
