#!/usr/bin/env python
"""Where the device time of one glocal gridding call goes (bench workload c2g): fit (all areas, one launch) vs per-area
gridding (targets + IDW points + blend) vs D2H, from pb_last_timing (h2d_ms = up to the end of the fit)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import praga_b200 as pb  # noqa: E402
from bench import WORKLOADS, GlocalWorkload  # noqa: E402

eng = pb.Engine(0)
wl = GlocalWorkload(eng, WORKLOADS["c2g"], 0, 1, 8)
rows = []
for k in range(8):
    wl.device_step(k)
    t = eng.timing()
    rows.append({k2: t[k2] for k2 in ("h2d_ms", "kernel_ms", "d2h_ms", "total_ms", "launches")})
print(json.dumps({"fit_ms(h2d_ms)": [r["h2d_ms"] for r in rows[3:]], "gridding_ms(kernel_ms)": [r["kernel_ms"] for r in rows[3:]],
                  "total_ms": [r["total_ms"] for r in rows[3:]], "launches": rows[-1]["launches"]}))
