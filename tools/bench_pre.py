#!/usr/bin/env python
"""Timing of the batched preInterpolation kernel (K3): T time steps x n stations (config 5: 365 d x 24 h)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import praga_b200 as pb  # noqa: E402
from praga_b200 import _abi as A, synthetic as syn  # noqa: E402

T, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8760, 500)
dem = syn.make_dem(400, 400, cell_size=500.0)
urban = syn.make_urban(400, 400, cell_size=500.0)
st = syn.make_stations(dem, n, proxies=(dem, urban))
s = syn.settings_multiple_detrending()
rng = np.random.default_rng(0)
hours = np.arange(T)
values = (st.value[None, :] + 4.0 * np.sin(2 * np.pi * hours / 24)[:, None] + rng.normal(0, 0.4, (T, n))).astype(np.float32)
eng = pb.Engine(0)
ms = []
for it in range(2):
    det, keep, fitted = eng.pre_interpolation_batch(st, values, s, A.airTemperature)
    ms.append(eng.timing()["kernel_ms"])
print(json.dumps({"kernel": "pre_interpolation_batch_kernel", "T": T, "stations": n, "kernel_ms": ms,
                  "steps_per_s": T / (min(ms) * 1e-3), "ms_per_step_amortised": min(ms) / T}))
