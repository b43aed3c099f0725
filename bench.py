#!/usr/bin/env python
"""bench.py — gridded cells/s of the interpolationRaster hot path on B200 (driver contract: one JSON line).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3idw] [--impl reference]

Workload (default "c2" = BASELINE.json configs[1]): synthetic 2000x2000 DEM @100 m, 500 stations, IDW with
multipleDetrending (elevation double-piecewise + urban fraction linear), airTemperature. A "step" = one time
step = one interpolationRaster call over the whole raster with that step's station values.

  value : valid cells gridded / device time, DEM + proxy rasters resident in HBM (pb_interpolation_raster_device).
  e2e   : the same through the drop-in C-ABI call with HOST buffers (pb_interpolation_raster): DEM + proxy H2D,
          station H2D, output raster D2H into host memory, all inside the timed region, every step.
  N > 1 : time steps are sharded over ranks (one process per GPU, no data-path collective); weak scaling.
  --impl reference : the reference's own CPU interpolationRaster (oracle/_ref, built from the unmodified
          sources) with all host threads, each step a bounded row-band sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from praga_b200 import _abi as A  # noqa: E402
from praga_b200 import synthetic as syn  # noqa: E402

METRIC = "gridded cells/sec (detrended IDW)"
UNIT = "cells/s"

WORKLOADS = {
    "c2": dict(rows=2000, cols=2000, stations=500, desc="synthetic 2000x2000 DEM @100 m, 500 stations, IDW + "
               "multipleDetrending (elevation double_piecewise + urbanFraction linear), airTemperature"),
    "c3idw": dict(rows=10000, cols=10000, stations=5000, desc="synthetic 10000x10000 DEM @100 m, 5000 stations, IDW + "
                  "multipleDetrending (elevation + urbanFraction), airTemperature"),
    "tiny": dict(rows=256, cols=256, stations=100, desc="tiny debug workload"),
}


def build_workload(name, n_steps, first_step=0):
    """DEM, urban proxy, settings with an already fitted detrending model and per-step detrended station sets.

    interpolationRaster receives points that preInterpolation has already detrended and settings that carry the
    fitted model (project.cpp:3108-3150), so that is what a step is given. The fit itself is station-space work
    of a few ms; here the model parameters are fixed plausible values and the residuals are computed with numpy."""
    w = WORKLOADS[name]
    dem = syn.make_dem(w["rows"], w["cols"])
    urban = syn.make_urban(w["rows"], w["cols"])
    s = syn.settings_multiple_detrending()
    s.currentSignificant[0] = s.currentSignificant[1] = 1
    s.nFit = s.nFitFunctions = 2
    s.fitFunction[0], s.fitNrParams[0] = A.piecewiseTwo, 4
    s.fitFunction[1], s.fitNrParams[1] = A.linear, 2
    el = (400.0, 13.3, 0.0045, -0.0065)
    ur = (2.0, -0.6)
    for j, v in enumerate(el):
        s.fitParams[0][j] = v
    for j, v in enumerate(ur):
        s.fitParams[1][j] = v
    base = syn.make_stations(dem, w["stations"], proxies=(dem, urban))
    z = base.proxy_values[:, 0].astype(np.float64)
    u = base.proxy_values[:, 1].astype(np.float64)
    trend = np.where(z < el[0], el[2], el[3]) * (z - el[0]) + el[1] + ur[0] * u + ur[1]
    steps = []
    for k in range(n_steps):
        rng = np.random.default_rng(1000 + first_step + k)
        st = base.copy()
        st.value[:] = (base.value.astype(np.float64) - trend + rng.normal(0, 0.3, base.n)).astype(np.float32)
        steps.append(st)
    return dem, urban, s, steps


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def run_reference(args, rank, world):
    """--impl reference: the unmodified reference's CPU interpolationRaster on the box's host cores."""
    if rank != 0:
        return
    from oracle import binding
    kind = "reference" if binding.have_ref() else "port"
    orc = binding.Oracle(kind)
    w = WORKLOADS[args.workload]
    # bounded sample: a band of full rows sized for ~1 s per step at ~1e6 cells/s
    sample_rows = max(8, min(w["rows"], int(1.0e6 * 500 / max(1, w["stations"]) / w["cols"])))
    dem, urban, s, steps = build_workload(args.workload, args.steps + args.warmup)
    r0 = (w["rows"] - sample_rows) // 2
    band = A.Raster(dem.data[r0:r0 + sample_rows], dem.llx, dem.lly + (w["rows"] - r0 - sample_rows) * dem.cell_size,
                    dem.cell_size, dem.flag)
    threads = orc.max_threads()
    total_cells, total_s = 0, 0.0
    for k, st in enumerate(steps):
        ok, grid, mn, mx, nv, sec = orc.interpolation_raster(st, s, A.airTemperature, band, (dem, urban), threads=0)
        if k >= args.warmup:
            total_cells += nv
            total_s += sec
    value = total_cells / total_s
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 distances / f64 weights (reference arithmetic)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{sample_rows} full rows ({total_cells // args.steps} valid cells) of the "
                                       f"{w['rows']}x{w['cols']} raster per step, {args.steps} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(args):
    """reference CPU path on a bounded sample (rank 0, N = 1 only): ~10-30 s of CPU work."""
    try:
        from oracle import binding
        kind = "reference" if binding.have_ref() else "port"
        orc = binding.Oracle(kind)
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:100]}
    w = WORKLOADS[args.workload]
    dem, urban, s, steps = build_workload(args.workload, 1)
    threads = orc.max_threads()
    target_cells = 1.2e7 * 500 / max(1, w["stations"])        # ~10-20 s at ~1e6 cells/s (500 stations)
    sample_rows = int(max(8, min(w["rows"], target_cells / w["cols"])))
    r0 = (w["rows"] - sample_rows) // 2
    band = A.Raster(dem.data[r0:r0 + sample_rows], dem.llx, dem.lly + (w["rows"] - r0 - sample_rows) * dem.cell_size,
                    dem.cell_size, dem.flag)
    ok, grid, mn, mx, nv, sec = orc.interpolation_raster(steps[0], s, A.airTemperature, band, (dem, urban), threads=0)
    return {"value": nv / sec, "unit": UNIT, "cores": threads, "kind": kind, "seconds": sec,
            "sample": f"{sample_rows} full rows ({nv} valid cells) of the {w['rows']}x{w['cols']} raster, 1 pass, "
                      f"all {threads} host threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="praga_b200", choices=["praga_b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # torchrun pins OMP_NUM_THREADS=1; the drop-in call gathers/scatters host rasters with the host cores, so every rank
    # gets its share of them (must be set before libgomp starts, i.e. before the engine library runs a parallel region)
    if world > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import praga_b200 as pb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: praga_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = pb.Engine(local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)

    w = WORKLOADS[args.workload]
    n_total = args.warmup + args.steps
    dem, urban, s, steps = build_workload(args.workload, n_total, first_step=rank * n_total)
    var = A.airTemperature

    # one-off upload of the static rasters (reported separately from the per-step numbers)
    t0 = time.perf_counter()
    dem_id = eng.upload(dem)
    urb_id = eng.upload(urban)
    n_valid = eng.valid_cells(dem_id)
    torch.cuda.synchronize()
    upload_s = time.perf_counter() - t0

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp32_peak, mufu_peak = eng.measure_pipe_peaks()
    out_host = np.empty(dem.shape, dtype=np.float32)

    # ---------------- device-resident throughput (`value`) ----------------
    for k in range(args.warmup):
        eng.interpolation_raster_resident(steps[k], s, var, dem_id, (dem_id, urb_id), device_only=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, kernel_ms, launches, cells = 0.0, 0.0, 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    for k in range(args.warmup, n_total):
        flush.zero_()                                   # L2 flush between timed iterations (not timed)
        e0.record()
        ok, _, mn, mx, nv = eng.interpolation_raster_resident(steps[k], s, var, dem_id, (dem_id, urb_id), device_only=True)
        e1.record()
        e1.synchronize()
        dev_ms += e0.elapsed_time(e1)
        t = eng.timing()
        kernel_ms += t["kernel_ms"]
        launches += t["launches"]
        cells += nv
    barrier()
    wall_s = time.perf_counter() - wall0
    clocks = sampler.stop()

    # ---------------- end to end through the drop-in call with host buffers (`e2e`) ----------------
    for k in range(min(2, args.warmup)):
        eng.interpolation_raster(steps[k], s, var, dem, (dem, urban), out=out_host)
    barrier()
    e2e_s, e2e_cells, h2d_b, d2h_b = 0.0, 0, 0, 0
    for k in range(args.warmup, n_total):
        flush.zero_()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ok, _, mn, mx, nv = eng.interpolation_raster(steps[k], s, var, dem, (dem, urban), out=out_host)
        e2e_s += time.perf_counter() - t1
        t = eng.timing()
        h2d_b += t["h2d_bytes"]
        d2h_b += t["d2h_bytes"]
        e2e_cells += nv
    barrier()
    # cached-raster variant (DEM/proxies resident, stations H2D + output D2H per step), reported as extra context
    res_s = 0.0
    for k in range(args.warmup, n_total):
        t1 = time.perf_counter()
        eng.interpolation_raster_resident(steps[k], s, var, dem_id, (dem_id, urb_id), out=out_host)
        res_s += time.perf_counter() - t1
    barrier()

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s, res_s, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, res_s, kernel_ms = tt.tolist()
        cc = torch.tensor([cells, e2e_cells, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        cells, e2e_cells, launches = cc.tolist()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        S = w["stations"]
        kernel_s_per_launch = kernel_ms * 1e-3 / args.steps
        pairs_per_launch = n_valid * S
        achieved_tflops = 11.0 * pairs_per_launch / kernel_s_per_launch * 1e-12
        line = {
            "metric": METRIC, "value": cells / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (f64 cross-tile accumulation and retrend)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "valid_cells_per_step": n_valid,
                       "stations": S, "steps_per_rank": args.steps, "sharding": "time steps over ranks, no collective",
                       "l2": "flushed (256 MiB write) between timed iterations",
                       "static_raster_upload_s": upload_s},
            "e2e": {"value": e2e_cells / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b / args.steps),
                    "d2h_bytes_per_step": int(d2h_b / args.steps), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "what": "pb_interpolation_raster with host DEM/proxy/station buffers and host output grid",
                    "resident_rasters_value": e2e_cells / res_s},
            "gpu_launches": int(launches),
            "wall_s_timed_region": wall_s,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp32_peak, "traffic": None,
                         "kernel": "pb::idw_kernel<2,false>", "kernel_ms_per_launch": kernel_s_per_launch * 1e3,
                         "algorithmic": f"11 flop x {S} stations x {n_valid} cells per launch (SURVEY.md 8d)",
                         "peak_source": "FFMA2 chain measured live on this GPU (MEASURED_PEAKS.json has no FP32 figure; "
                                        f"its hbm/bf16 numbers are '{peak_kind}')",
                         "mufu_bound": {"pairs_per_s": pairs_per_launch / kernel_s_per_launch,
                                        "mufu_peak_per_s": mufu_peak * 1e9,
                                        "frac": pairs_per_launch / kernel_s_per_launch / (mufu_peak * 1e9)},
                         "hbm": {"algorithmic_bytes_per_cell": 12,
                                 "achieved_gbs": 12.0 * n_valid / kernel_s_per_launch * 1e-9,
                                 "peak_gbs": peaks.get("hbm_gbs")}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(args)
        print(json.dumps(line), flush=True)

    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
