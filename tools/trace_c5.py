#!/usr/bin/env python
"""Device timeline of one step of the hourly batch (c5) through torch.profiler (CUPTI sees every kernel of the process, the
engine's included): per-kernel busy time, how much of the step some kernel is running at all, and how much of it two or more
run at once. python tools/trace_c5.py [lanes] -> one JSON line (+ the chrome trace under gpurun_out/ if the directory exists)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    os.environ["PB_BENCH_LANES"] = sys.argv[1]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import praga_b200 as pb  # noqa: E402
from bench import HourlyBatchWorkload, WORKLOADS  # noqa: E402

eng = pb.Engine(0)
wl = HourlyBatchWorkload(eng, WORKLOADS["c5"], 0, 1, 6)
for k in range(3):
    wl.device_step(k)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    wl.device_step(3)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
spans = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
t0, t1 = spans[0][0], max(s[1] for s in spans)
busy = {}
for a, b, n in spans:
    key = n.split("(")[0][:60]
    busy[key] = busy.get(key, 0.0) + (b - a)
# coverage: sweep
pts = []
for a, b, _ in spans:
    pts.append((a, 1))
    pts.append((b, -1))
pts.sort()
cov1 = cov2 = 0.0
depth, last = 0, pts[0][0]
for t, d in pts:
    if depth >= 1:
        cov1 += t - last
    if depth >= 2:
        cov2 += t - last
    depth += d
    last = t
out = {"lanes": wl.lanes, "step_us": t1 - t0, "some_kernel_running_us": cov1, "two_or_more_running_us": cov2, "kernels": len(spans),
       "busy_us_by_kernel": dict(sorted(busy.items(), key=lambda kv: -kv[1])[:14])}
print(json.dumps(out))
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
    prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", "c5_trace.json"))
# per-kernel launch count and mean duration
import re  # noqa: E402
per = {}
for a, b, n in spans:
    m = re.search(r"(\w+)(<[^>]*>)?\(", n)
    k = (m.group(1) if m else n)[:40]
    c = per.setdefault(k, [0, 0.0])
    c[0] += 1
    c[1] += b - a
print(json.dumps({"mean_us_by_kernel": {k: [v[0], round(v[1] / v[0], 1)] for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:12]}}))
