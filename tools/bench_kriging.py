#!/usr/bin/env python
"""Timing of ordinary kriging (config 4 shape: n stations, spherical, range 2e5, nugget 0.1, sill 4) on a raster band.
python tools/bench_kriging.py [n rows cols] -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import praga_b200 as pb  # noqa: E402
from praga_b200 import _abi as A, synthetic as syn  # noqa: E402

n, rows, cols = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2000, 1000, 1000)
cs = 1.0e6 / max(rows, cols)
dem = syn.make_dem(rows, cols, cell_size=cs)
rng = np.random.default_rng(1)
pos = np.stack([dem.llx + rng.random(n) * cols * cs, dem.lly + rng.random(n) * rows * cs], axis=1)
val = rng.normal(12.0, 3.0, n)
eng = pb.Engine(0)
t0 = time.perf_counter()
kid = eng.kriging_setup(pos, val, A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0)
setup_s = time.perf_counter() - t0
did = eng.upload(dem)
out = {"n": n, "rows": rows, "cols": cols, "setup_s": setup_s, "tensor_path": eng.kriging_uses_tensor_path(kid)}
for label, want in (("estimate_only", False), ("estimate_and_rmse", True)):
    ms = []
    for it in range(3):
        est, rm, nv = eng.kriging_raster(kid, did, dem.shape, want_rmse=want, device_only=True)
        ms.append(eng.timing()["kernel_ms"])
    out[label] = {"kernel_ms": ms, "cells_per_s": nv / (min(ms) * 1e-3)}
out["valid_cells"] = nv
var_ms = min(out["estimate_and_rmse"]["kernel_ms"]) - min(out["estimate_only"]["kernel_ms"])
out["variance_ms"] = var_ms
out["variance_algorithmic_tflops"] = 2.0 * n * (n + 1) * nv / (var_ms * 1e-3) * 1e-12
print(json.dumps(out))
