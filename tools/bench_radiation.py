#!/usr/bin/env python
"""Timing of pb_compute_radiation_dem (radiation::computeRadiationDEM, solarRadiation.cpp:1045-1069) on a synthetic DEM.
python tools/bench_radiation.py [rows cols cell_size] -> one JSON line (device time from pb_last_timing).
The static maps are synthetic inputs here (lat / lon linear in the cell position, slope / aspect from numpy gradients): the
tool measures the kernel, parity lives in tests/test_radiation.py."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import praga_b200 as pb  # noqa: E402
from praga_b200 import _abi as A, synthetic as syn  # noqa: E402

rows, cols, cs = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (4000, 4000, 100.0)
dem = syn.make_dem(rows, cols, cell_size=cs)
z = np.where(dem.data == np.float32(dem.flag), np.float32(500.0), dem.data * np.float32(2.0))
dem.data[dem.data != np.float32(dem.flag)] *= np.float32(2.0)
r, c = np.indices(dem.shape, dtype=np.float32)
lat = (46.0 - r * np.float32(cs / 111000.0)).astype(np.float32)
lon = (9.0 + c * np.float32(cs / 78000.0)).astype(np.float32)
gy, gx = np.gradient(z, cs)
slope = np.degrees(np.arctan(np.hypot(gx, gy))).astype(np.float32)
aspect = (np.degrees(np.arctan2(-gx, gy)) % 360.0).astype(np.float32)
tr = np.full(dem.shape, 0.6, dtype=np.float32)
eng = pb.Engine(0)
mk = lambda a: eng.upload(A.Raster(a, dem.llx, dem.lly, dem.cell_size, dem.flag))  # noqa: E731
ids = [eng.upload(dem), mk(lat), mk(lon), mk(slope), mk(aspect), mk(tr)]
n_valid = int((dem.data != np.float32(dem.flag)).sum())
out = {}
for label, sec in (("noon", 12 * 3600), ("low_sun", 8 * 3600), ("night", 2 * 3600)):
    s = A.default_radiation_settings(2023, 1 if label == "low_sun" else 6, 21, sec, 1)
    s.demMaximum = float(z.max())
    ms = []
    for it in range(3):
        eng.compute_radiation_dem(s, *ids, dem.shape, want_host=False)
        ms.append(eng.timing()["kernel_ms"])
    out[label] = {"kernel_ms": min(ms), "cells_per_s": n_valid / (min(ms) * 1e-3),
                  # algorithmic HBM bytes per cell: 6 input maps read + 5 output maps written, 4 B each
                  "hbm_gbs": 44.0 * rows * cols / (min(ms) * 1e-3) * 1e-9}
print(json.dumps({"kernel": "radiation_dem_kernel", "rows": rows, "cols": cols, "valid_cells": n_valid, "cases": out}))
