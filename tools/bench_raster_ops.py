#!/usr/bin/env python
"""HBM roofline of the raster map kernels (K9) on a 10k x 10k grid: algorithmic bytes / device time vs the measured copy
bandwidth of MEASURED_PEAKS.json. python tools/bench_raster_ops.py [rows cols] -> one JSON line per operator."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import praga_b200 as pb  # noqa: E402
from praga_b200 import _abi as A, synthetic as syn  # noqa: E402

rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 10000)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
dem = syn.make_dem(rows, cols)
flag = np.float32(dem.flag)
rng = np.random.default_rng(1)
t = (15.0 - 0.0065 * np.where(dem.data == flag, 0, dem.data)).astype(np.float32)
td = (t - 3.0).astype(np.float32)
t[dem.data == flag] = flag
td[dem.data == flag] = flag
T = A.Raster(t, dem.llx, dem.lly, dem.cell_size, dem.flag)
TD = A.Raster(td, dem.llx, dem.lly, dem.cell_size, dem.flag)
eng = pb.Engine(0)
d_id, t_id, td_id = eng.upload(dem), eng.upload(T), eng.upload(TD)
n = rows * cols


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        fn()
        best = min(best, eng.timing()["kernel_ms"])
    return best


def line(name, ms, bytes_per_cell, cells, extra=None):
    gbs = bytes_per_cell * cells / (ms * 1e-3) * 1e-9
    out = {"kernel": name, "cells": cells, "kernel_ms": ms, "cells_per_s": cells / (ms * 1e-3),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                        "algorithmic_bytes_per_cell": bytes_per_cell}}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


ms = timed(lambda: eng.relative_humidity_map(t_id, td_id, dem.shape, want_host=False))
line("relative_humidity_map_kernel", ms, 12, n)
ms = timed(lambda: eng.fix_daily_thermal_consistency(t_id, td_id))
line("thermal_consistency_kernel", ms, 8, n, {"note": "reads tmax and tmin; writes only the cells it fixes"})
h = A.pb_raster_header()
f = 5.0
h.cellSize, h.invCellSize = dem.cell_size * f, 1.0 / (dem.cell_size * f)
h.nrRows, h.nrCols, h.llx, h.lly, h.flag = int(rows / f), int(cols / f), dem.llx, dem.lly, dem.flag
ms = timed(lambda: eng.resample_grid(d_id, h, A.PB_AGGR_AVERAGE, 0.0, want_host=False))
line("resample_grid_kernel<average> 100 m -> 500 m", ms, 4.0 * 36 / 25 + 4.0 / 25, n,
     {"note": "36 samples per new cell over 25 old cells: 5.8 B per OLD cell read + 0.16 written"})
small = A.Raster(dem.data[:2000, :2000].copy(), dem.llx, dem.lly, dem.cell_size, dem.flag)
s_id = eng.upload(small)
ms = timed(lambda: eng.topographic_distance_map(s_id, small.shape, dem.llx + 70000.0, dem.lly + 90000.0, 600.0, want_host=False), reps=3)
line("topographic_distance_map_kernel 2000x2000", ms, 8, 2000 * 2000,
     {"note": "ray march: ~d/cellSize DEM samples per cell from L2, not an HBM-streaming kernel; bytes = DEM read + map written"})

# DEM -> meteo-grid aggregation (spatialAggregateMeteoGrid): 5 km cells over the 100 m raster, every DEM cell sampled once
cell = 5000.0
k = int(cell / dem.cell_size)
ncx, ncy = cols // k, rows // k
gx = dem.llx + (np.arange(cols // k * k) + 0.5) * dem.cell_size
gy = dem.lly + (np.arange(rows // k * k) + 0.5) * dem.cell_size
X, Y = np.meshgrid(gx, gy, indexing="xy")
X = X.reshape(ncy, k, ncx, k).transpose(0, 2, 1, 3).reshape(-1)
Y = Y.reshape(ncy, k, ncx, k).transpose(0, 2, 1, 3).reshape(-1)
first = np.arange(ncx * ncy + 1, dtype=np.int64) * (k * k)
aid = eng.aggregation_upload(X, Y, first, np.full(ncx * ncy, k * k, np.int32))
for name, m in (("average", A.PB_AGGR_AVERAGE), ("median", A.PB_AGGR_MEDIAN), ("stdDeviation", A.PB_AGGR_STDDEV)):
    ms = timed(lambda: eng.spatial_aggregate_meteo_grid(t_id, aid, m), reps=3)
    line(f"spatial_aggregate_kernel<{name}> 100 m -> 5 km meteo grid", ms, 4 + 16 + 4, ncx * ncy * k * k,
         {"meteo_grid_cells": ncx * ncy, "samples_per_cell": k * k, "d2h_bytes": 4 * ncx * ncy,
          "note": "24 B per sample (x, y doubles + raster value + scratch); one warp per meteo-grid cell, lane 0 reduces in list "
                  "order (float arithmetic of the reference): latency bound, not an HBM-streaming kernel"})

# topographicIndex: 2000 x 2000 cells @100 m, 5 km window (cellNr = 50: 10 201 offsets per cell, ~7 800 inside the disc)
import time  # noqa: E402
ms = timed(lambda: eng.topographic_index(s_id, small.shape, [5000.0], want_host=False), reps=2)
visits = 4.0e6 * 101 * 101
line("topographic_index_kernel 2000x2000, 5 km window", ms, 8, 2000 * 2000,
     {"window_visits_per_s": visits / (ms * 1e-3),
      "note": "stencil: 10 201 window offsets per cell from L1/L2, float accumulators in the reference's order; compute bound, not HBM"})
try:
    from oracle import binding  # noqa: E402
    orc = binding.Oracle("reference" if binding.have_ref() else "port")
    sub = A.Raster(dem.data[4000:4300, 4000:4300].copy(), dem.llx, dem.lly, dem.cell_size, dem.flag)
    t0 = time.perf_counter()
    orc.topographic_index(sub, [5000.0])
    dt = time.perf_counter() - t0
    print(json.dumps({"kernel": "topographicIndex reference (single thread, as shipped)", "cells": 300 * 300, "seconds": dt,
                      "cells_per_s": 300 * 300 / dt, "note": "300x300 sample, same 5 km window (border cells see smaller windows)"}), flush=True)
except Exception as e:  # noqa: BLE001
    print(json.dumps({"kernel": "topographicIndex reference", "error": str(e)[:100]}))
