#!/usr/bin/env python3
"""tools/sanitize_case.py -- a small run that touches every kernel (3D TRT/BGK, all modes, restart,
macrovars, D2Q9) for `compute-sanitizer --tool memcheck python tools/sanitize_case.py`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbmpkg import lbm  # noqa: E402

for cfg in (dict(N=9, Re=20, METHOD=0), dict(N=12, Re=100, DIAG=1), dict(N=33, Re=100)):
    with lbm.Cavity3D(**cfg) as s:
        s.step(1); s.residual(); s.step(5); s.residual()
        f = s.f(); mv = s.macrovars()
        s.set_f(f, f); s.step(3); s.residual()
    with lbm.Cavity3D(track_residual=False, **cfg) as s:
        s.step(4); s.f(); s.step(2); s.macrovars()
with lbm.Cavity2D(N=17, M=9, Re=25.0) as s:
    s.step(7); s.residual(); d = s.dist(); s.set_dist(d); s.step(2); s.macro()
with lbm.Cavity2D(N=300, M=5, Re=100.0) as s:
    s.step(3); s.residual()
print("sanitize case done")
