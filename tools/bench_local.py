#!/usr/bin/env python
"""Timing of the local-detrending raster kernel (K2) at config-3 station density (5000 stations / 1000 km x 1000 km):
python tools/bench_local.py [rows cols cell_size stations]  -> one JSON line (device time from pb_last_timing)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import praga_b200 as pb  # noqa: E402
from praga_b200 import _abi as A, synthetic as syn  # noqa: E402

rows, cols, cs, S = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (2000, 2000, 500.0, 5000)
dem = syn.make_dem(rows, cols, cell_size=cs)
urban = syn.make_urban(rows, cols, cell_size=cs)
st = syn.make_stations(dem, S, proxies=(dem, urban))
s = syn.settings_multiple_detrending()
s.useLocalDetrending = 1
s.minPointsLocalDetrending = 20
syn.set_points_range(s, st)
eng = pb.Engine(0)
d, u = eng.upload(dem), eng.upload(urban)
res = []
for it in range(3):
    ok, _, mn, mx, nv = eng.interpolation_raster_resident(st, s, A.airTemperature, d, (d, u), device_only=True, local=True)
    t = eng.timing()
    res.append(t["kernel_ms"])
print(json.dumps({"kernel": "local_detrending_kernel", "rows": rows, "cols": cols, "cell_size": cs, "stations": S,
                  "valid_cells": nv, "kernel_ms": res, "cells_per_s": nv / (min(res) * 1e-3), "min": mn, "max": mx}))
