#!/usr/bin/env python
"""Where the end-to-end time of the drop-in call goes (C2): host wall time vs the device phases of pb_last_timing."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import praga_b200 as pb  # noqa: E402
from praga_b200 import _abi as A  # noqa: E402
from bench import build_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
dem, urban, s, steps = build_workload(name, 8)
eng = pb.Engine(0)
out = np.empty(dem.shape, dtype=np.float32)
rows = []
for k in range(8):
    t0 = time.perf_counter()
    eng.interpolation_raster(steps[k], s, A.airTemperature, dem, (dem, urban), out=out)
    wall = (time.perf_counter() - t0) * 1e3
    t = eng.timing()
    rows.append({"wall_ms": wall, **{k2: t[k2] for k2 in ("h2d_ms", "kernel_ms", "d2h_ms", "total_ms", "launches")}})
print(json.dumps(rows[3:], indent=0))
