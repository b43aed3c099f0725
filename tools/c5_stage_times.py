#!/usr/bin/env python
"""Where one (hour, variable) unit of the hourly batch spends its time: the stages of pb_interpolation_dem called one by one."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import praga_b200 as pb  # noqa: E402
from praga_b200 import _abi as A, synthetic as syn  # noqa: E402
from bench import HourlyBatchWorkload, WORKLOADS  # noqa: E402

eng = pb.Engine(0)
wl = HourlyBatchWorkload(eng, WORKLOADS["c5"], 0, 1, 8)
out = {}
t_st, rh_st, own_t, p_st = wl._hour_inputs(0)
for st, var in ((t_st, A.airTemperature), (rh_st, A.airRelHumidity), (p_st, A.precipitation)):
    acc = {}
    for rep in range(6):
        s, sq = wl._settings(var)
        t0 = time.perf_counter()
        wrong = eng.spatial_quality_control(st.copy(), sq, var) if var != A.precipitation else np.zeros(st.n, bool)
        t1 = time.perf_counter()
        pts = st.subset(~wrong)
        raw = pts.value.copy()
        syn.set_points_range(s, pts)
        keep, _ = eng.pre_interpolation_ex(pts, s, var, current_value=raw)
        t2 = time.perf_counter()
        ip = pts.subset(keep)
        eng.interpolation_raster_resident(ip, s, var, wl.dem_id, (wl.dem_id, wl.urb_id), device_only=True)
        t3 = time.perf_counter()
        meteo = st.subset(~wrong)
        meteo.value[:] = raw
        eng.compute_residuals_exact(meteo, ip, s, var, observed=raw, exclude_supplemental=True)
        t4 = time.perf_counter()
        if rep >= 2:
            for name, dt in (("qc", t1 - t0), ("pre", t2 - t1), ("raster", t3 - t2), ("loo", t4 - t3)):
                acc[name] = acc.get(name, 0.0) + dt * 1e3 / 4
    out[int(var)] = {k2: round(v, 3) for k2, v in acc.items()}
print(json.dumps(out))
