#!/usr/bin/env python
"""bench.py — gridded cells/s of the interpolationRaster hot path on B200 (driver contract: one JSON line).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3idw|c3|c4|c5] [--impl reference]

Workload (default "c2" = BASELINE.json configs[1]): synthetic 2000x2000 DEM @100 m, 500 stations, IDW with
multipleDetrending (elevation double-piecewise + urban fraction linear), airTemperature. A "step" = one time
step = one interpolationRaster call over the whole raster with that step's station values.

  value : valid cells gridded / device time, DEM + proxy rasters resident in HBM (pb_interpolation_raster_device).
  e2e   : the same through the drop-in C-ABI call with HOST buffers (pb_interpolation_raster): DEM + proxy H2D,
          station H2D, output raster D2H into host memory, all inside the timed region, every step.
  N > 1 : time steps are sharded over ranks (one process per GPU, no data-path collective); weak scaling.
  --impl reference : the reference's own CPU interpolationRaster (oracle/_ref, built from the unmodified
          sources) with all host threads, each step a bounded row-band sample of the same workload.

Other BASELINE.json configs, same JSON line (not the driver's default; results under profiles/):
  c3 : 10000x10000 DEM, 5000 stations, localDetrending + IDW (K2); a step = one band of rows, bands sharded over ranks.
  c4 : 10000x10000 DEM, 2000 stations, ordinary kriging estimate + RMSE (K5); a step = one band of rows.
  c5 : hourly batch on the 2000x2000 DEM, 500 stations: a step = Project::interpolationDem for one (hour, variable) of
       air temperature / relative humidity / precipitation = spatial QC + preInterpolation + raster + leave-one-out.
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
    "c3": dict(rows=10000, cols=10000, stations=5000, band_rows=400,
               desc="synthetic 10000x10000 DEM @100 m, 5000 stations, localDetrending (minPoints 20) + multipleDetrending + IDW, "
                    "airTemperature; topographic distance is off in local mode as in the reference (getUseTD())"),
    "c4": dict(rows=10000, cols=10000, stations=2000, band_rows=1000,
               desc="synthetic 10000x10000 DEM @100 m, 2000 stations, ordinary kriging (spherical, range 2e5 m, nugget 0.1, "
                    "sill 4), estimate + RMSE grids"),
    "c2g": dict(rows=2000, cols=2000, stations=500, areas_side=3, blend_cells=100, margin_m=30000.0,
                desc="synthetic 2000x2000 DEM @100 m, 500 stations, glocal detrending: 3x3 macro areas with 10 km blended borders, "
                     "per-area multipleDetrending (elevation + urbanFraction) + IDW over the area's stations, airTemperature"),
    "c5": dict(rows=2000, cols=2000, stations=500,
               desc="hourly batch on the synthetic 2000x2000 DEM @100 m, 500 stations: airTemperature (multipleDetrending), "
                    "airRelHumidity, precipitation; per step spatial quality control + preInterpolation + raster + "
                    "leave-one-out cross-validation (Project::interpolationDem / interpolationCv)"),
}
GENERIC = ("c3", "c4", "c5", "c2g")


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

    def wait_first(self, seconds=5.0):
        """nvidia-smi takes 0.1-1 s to print its first line (longer with 8 ranks starting it at once): the timed regions of the
        fast workloads are shorter than that, so the caller waits here BEFORE the warm-up"""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < seconds:
            time.sleep(0.02)

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
                "samples": len(sm), "reasons": sorted(reasons),
                "window": "100 ms samples from the warm-up through the device-timed and end-to-end regions"}


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
    # repeat the pass until ~10 s of CPU work are timed (the reference is fast on many-core hosts); best pass reported
    total, best, passes = 0.0, None, 0
    while total < 10.0 and passes < 40:
        ok, grid, mn, mx, nv, sec = orc.interpolation_raster(steps[0], s, A.airTemperature, band, (dem, urban), threads=0)
        total += sec
        passes += 1
        best = sec if best is None else min(best, sec)
    return {"value": nv / best, "unit": UNIT, "cores": threads, "kind": kind, "seconds": total,
            "sample": f"{sample_rows} full rows ({nv} valid cells) of the {w['rows']}x{w['cols']} raster, best of {passes} "
                      f"passes ({total:.1f} s of CPU work), all {threads} host threads"}


# ---------------------------------------------------------------------------------------------------------------------
# the other BASELINE configs (c3 local detrending, c4 kriging, c5 hourly batch): one harness, three workload objects
# ---------------------------------------------------------------------------------------------------------------------
def _bands(dem, band_rows):
    """interior row bands of band_rows rows (the masked north margin is skipped)."""
    rows = dem.shape[0]
    first = max(1, rows // 100)
    return [(r, min(rows, r + band_rows)) for r in range(first, rows - band_rows + 1, band_rows)]


def _band_raster(dem, r0, r1):
    rows = dem.shape[0]
    return A.Raster(dem.data[r0:r1], dem.llx, dem.lly + (rows - r1) * dem.cell_size, dem.cell_size, dem.flag)


class LocalDetrendingWorkload:
    """c3: K2. unit = a band of rows of the 10k x 10k raster; ranks take different bands (row-band sharding)."""
    kernel = "pb::local_detrending_kernel<4,0,1>"
    # FP64 flop per cell executed by the LM fits + estimate at this station density, counted with ncu
    # (thread-level DADD + DMUL + 2 DFMA of profiles/r01_local_k2e_flops.csv / 453 823 cells, 4 lanes per cell; data dependent;
    # the 16-lane form executed 4.22e5: more lanes replicate the group-uniform parts)
    FP64_FLOP_PER_CELL = 3.64e5

    def __init__(self, eng, w, rank, world, n_total):
        self.eng, self.w, self.var = eng, w, A.airTemperature
        self.dem = syn.make_dem(w["rows"], w["cols"])
        self.urban = syn.make_urban(w["rows"], w["cols"])
        self.base = syn.make_stations(self.dem, w["stations"], proxies=(self.dem, self.urban))
        self.s = syn.settings_multiple_detrending()
        self.s.useLocalDetrending = 1
        self.s.minPointsLocalDetrending = 20
        syn.set_points_range(self.s, self.base)
        self.bands = _bands(self.dem, w["band_rows"])
        self.rank, self.world = rank, world
        self.dem_id, self.urb_id = eng.upload(self.dem), eng.upload(self.urban)
        self.out = np.empty(self.dem.shape, dtype=np.float32)
        self.n_valid_ref = None

    def _step_inputs(self, k):
        r0, r1 = self.bands[(self.rank + k * self.world) % len(self.bands)]
        st = self.base.copy()
        st.value[:] = (self.base.value + np.random.default_rng(500 + k).normal(0, 0.3, st.n)).astype(np.float32)
        return st, r0, r1

    def device_step(self, k):
        st, r0, r1 = self._step_inputs(k)
        ok, _, mn, mx, nv = self.eng.interpolation_raster_resident(st, self.s, self.var, self.dem_id, (self.dem_id, self.urb_id),
                                                                   row_begin=r0, row_end=r1, device_only=True, local=True)
        return nv

    def e2e_step(self, k):
        st, r0, r1 = self._step_inputs(k)
        ok, _, mn, mx, nv = self.eng.interpolation_raster(st, self.s, self.var, self.dem, (self.dem, self.urban), row_begin=r0,
                                                          row_end=r1, out=self.out, local=True)
        return nv

    def roofline(self, cells_per_launch, kernel_s, peaks):
        peak = self.eng.measure_fp64_peak()
        ach = self.FP64_FLOP_PER_CELL * cells_per_launch / kernel_s * 1e-12
        return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "kernel": self.kernel, "kernel_ms_per_launch": kernel_s * 1e3,
                "algorithmic": f"{self.FP64_FLOP_PER_CELL:.3g} FP64 flop per cell (ncu-counted at this station density, data "
                               f"dependent) x {int(cells_per_launch)} cells per launch",
                "peak_source": "DFMA chain measured live on this GPU (MEASURED_PEAKS.json has no FP64 figure)",
                "selection_floor": {"fp32_flop_per_cell": 6 * self.w["stations"],
                                    "note": "6*S FP32 flop per cell for the selection distances (SURVEY 8d)"}}

    def config(self):
        return {"unit_of_work": f"one band of {self.w['band_rows']} rows per step", "bands": len(self.bands),
                "sharding": "row bands over ranks, no collective"}

    def cpu_sample(self, orc, seconds_hint):
        # ~2e-4 s per cell and thread in the reference: a few rows are ~10-30 s
        threads = orc.max_threads()
        rows = max(2, int(seconds_hint * threads / 2.2e-4 / (0.925 * self.w["cols"])))
        r0 = self.bands[len(self.bands) // 2][0]
        st, _, _ = self._step_inputs(0)
        ok, grid, mn, mx, nv, sec = orc.local_detrending_raster(st, self.s, self.var, self.dem, (self.dem, self.urban), threads=0,
                                                                row_begin=r0, row_end=r0 + rows)
        return nv, sec, threads, f"{rows} full rows ({nv} valid cells) of the {self.w['rows']}x{self.w['cols']} raster, 1 pass"


class KrigingWorkload:
    """c4: K5. unit = a band of rows; estimate + RMSE grids."""
    kernel = "pb::kriging_variance_tc_kernel (+ kriging_estimate_kernel)"

    def __init__(self, eng, w, rank, world, n_total):
        self.eng, self.w = eng, w
        self.dem = syn.make_dem(w["rows"], w["cols"])
        n = w["stations"]
        rng = np.random.default_rng(1)
        self.pos = np.stack([self.dem.llx + rng.random(n) * w["cols"] * self.dem.cell_size,
                             self.dem.lly + rng.random(n) * w["rows"] * self.dem.cell_size], axis=1)
        self.val = rng.normal(12.0, 3.0, n)
        self.vario = (A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0)
        t0 = time.perf_counter()
        self.kid = eng.kriging_setup(self.pos, self.val, *self.vario)
        self.setup_s = time.perf_counter() - t0
        self.tensor = eng.kriging_uses_tensor_path(self.kid)
        self.dem_id = eng.upload(self.dem)
        self.bands = _bands(self.dem, w["band_rows"])
        self.rank, self.world = rank, world
        self.est = np.empty(self.dem.shape, dtype=np.float32)
        self.rmse = np.empty(self.dem.shape, dtype=np.float32)

    def _band(self, k):
        return self.bands[(self.rank + k * self.world) % len(self.bands)]

    def device_step(self, k):
        r0, r1 = self._band(k)
        _, _, nv = self.eng.kriging_raster(self.kid, self.dem_id, self.dem.shape, r0, r1, want_rmse=True, device_only=True)
        return nv

    def e2e_step(self, k):
        r0, r1 = self._band(k)
        _, _, nv = self.eng.kriging_raster(self.kid, self.dem_id, self.dem.shape, r0, r1, want_rmse=True, est_out=self.est,
                                           rmse_out=self.rmse)
        return nv

    def roofline(self, cells_per_launch, kernel_s, peaks):
        n = self.w["stations"]
        flop = 2.0 * n * (n + 1) + 4.0 * n
        ach = flop * cells_per_launch / kernel_s * 1e-12
        peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "kernel": self.kernel, "kernel_ms_per_launch": kernel_s * 1e3,
                "algorithmic": f"(2 n (n+1) + 4 n) flop per cell, n = {n} (SURVEY 8d) x {int(cells_per_launch)} cells per launch; "
                               "time = estimate + variance kernels; the variance kernel issues 3 FP16 MMAs per K step "
                               "(hi/lo operand split) on the lower triangle only",
                "peak_source": "dense bf16 GEMM, sustained, MEASURED_PEAKS.json"}

    def config(self):
        return {"unit_of_work": f"one band of {self.w['band_rows']} rows per step", "bands": len(self.bands),
                "sharding": "row bands over ranks, no collective", "kriging_setup_s": self.setup_s,
                "tensor_path": bool(self.tensor == 1),
                "e2e_note": "the DEM is resident by API (pb_kriging_raster takes a raster id); per step: both output grids D2H"}

    def cpu_sample(self, orc, seconds_hint):
        # krigingSetWeight is O(n^2) per cell (~3 ms at n = 2000) and the reference keeps its system in file statics: 1 thread
        m = max(200, int(seconds_hint / 3.2e-3))
        r0, r1 = self.bands[len(self.bands) // 2]
        rng = np.random.default_rng(3)
        xy = np.stack([self.dem.llx + rng.random(m) * self.w["cols"] * self.dem.cell_size,
                       self.dem.lly + (self.w["rows"] - r1 + rng.random(m) * (r1 - r0)) * self.dem.cell_size], axis=1)
        ok, res, rm, t_setup, t_eval = orc.kriging_points(self.pos, self.val, *self.vario, xy)
        return m, t_eval, 1, (f"{m} random cells of one band (krigingSetWeight + krigingResult + krigingRMSE each), single "
                              f"thread: kriging.cpp keeps one system in file-static globals; krigingVariogram setup "
                              f"{t_setup:.1f} s not included")


class HourlyBatchWorkload:
    """c5: one (hour, variable) per step through pb_interpolation_dem (spatial QC + fit + raster + leave-one-out)."""
    kernel = "pb::idw_kernel + station-space kernels (K3/K7, loo_exact, neighbourhood)"
    VARS = (A.airTemperature, A.airRelHumidity, A.precipitation)
    ROOFLINE_ON_STEP_TIME = True

    def __init__(self, eng, w, rank, world, n_total):
        self.eng, self.w = eng, w
        self.dem = syn.make_dem(w["rows"], w["cols"])
        self.urban = syn.make_urban(w["rows"], w["cols"])
        n = w["stations"]
        grids = (self.dem, self.urban)
        self.base = {A.airTemperature: syn.make_stations(self.dem, n, proxies=grids),
                     A.airRelHumidity: syn.make_stations(self.dem, n, proxies=grids, variable="humidity"),
                     A.precipitation: syn.make_stations(self.dem, n, proxies=grids, variable="precipitation")}
        self.dem_id, self.urb_id = eng.upload(self.dem), eng.upload(self.urban)
        self.rank, self.world = rank, world
        self.outs = []
        self.lanes = int(os.environ.get("PB_BENCH_LANES", "16"))

    def _settings(self, var):
        if var == A.airTemperature:
            s = syn.settings_multiple_detrending()
        else:
            s = A.default_settings()
            s.nProxy = 2
            s.proxy[0] = A.default_proxy(A.proxyHeight, 0)
            s.proxy[1] = A.default_proxy(A.proxyUrbanFraction, 1)
        return s, syn.settings_classic_detrending(2)

    def _step_inputs(self, k):
        unit = self.rank + k * self.world                      # global (hour, variable) index: time steps over ranks
        hour, var = unit // 3, self.VARS[unit % 3]
        st = self.base[var].copy()
        rng = np.random.default_rng(9000 + hour)
        anomaly = rng.normal(0, 0.6, st.n)
        if var == A.airTemperature:
            st.value[:] = (st.value + anomaly + 3.0 * np.sin(hour / 24 * 2 * np.pi)).astype(np.float32)
        elif var == A.airRelHumidity:
            st.value[:] = np.clip(st.value + 6 * anomaly, 0, 100).astype(np.float32)
        else:
            st.value[:] = np.maximum(st.value + 2 * anomaly, 0).astype(np.float32)
        return st, var

    UNITS = int(os.environ.get("PB_BENCH_UNITS", "24"))      # (hour, variable) units per bench step (8 hours x 3 variables)

    def _run(self, k, device_only):
        units = []
        for j in range(self.UNITS):
            st, var = self._step_inputs(k * self.UNITS + j)
            s, sq = self._settings(var)
            units.append((st, s, sq, var))
        if not device_only and len(self.outs) < self.UNITS:
            self.outs = [np.empty(self.dem.shape, dtype=np.float32) for _ in range(self.UNITS)]
        res = self.eng.interpolation_dem_batch(units, self.dem_id, self.dem.shape, (self.dem_id, self.urb_id), cross_validate=True,
                                               device_only=device_only, lanes=self.lanes, outs=None if device_only else self.outs)
        return sum(r["n_valid"] for r in res)

    def device_step(self, k):
        return self._run(k, True)

    def e2e_step(self, k):
        return self._run(k, False)

    def roofline(self, cells_per_launch, kernel_s, peaks):
        fp32_peak, mufu_peak = self.eng.measure_pipe_peaks()
        S = self.w["stations"]
        ach = 11.0 * S * cells_per_launch / kernel_s * 1e-12
        return {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak, "traffic": None,
                "kernel": self.kernel, "kernel_ms_per_launch": kernel_s * 1e3,
                "algorithmic": f"11 flop x {S} stations x {int(cells_per_launch)} cells per step (raster only; the station-space "
                               "kernels of the step are counted in the time, not in the flops; time = device time of the whole step)",
                "peak_source": "FFMA2 chain measured live on this GPU"}

    def config(self):
        return {"unit_of_work": f"{self.UNITS} (hour, variable) units per step through pb_interpolation_dem_batch ({self.lanes} lanes); "
                                "each unit: QC + preInterpolation + raster + leave-one-out",
                "sharding": "time steps over ranks, no collective",
                "e2e_note": "DEM / proxy rasters resident by API (pb_interpolation_dem takes raster ids); per step: station "
                            "arrays H2D, output raster D2H"}

    def cpu_sample(self, orc, seconds_hint):
        threads = orc.max_threads()
        cells, sec = 0, 0.0
        for k in range(3):                                      # one step per variable
            st, var = self._step_inputs(k)
            s, sq = self._settings(var)
            t0 = time.perf_counter()
            r = orc.interpolation_dem(st, s, var, self.dem, (self.dem, self.urban), quality_settings=sq, cross_validate=True)
            sec += time.perf_counter() - t0
            cells += r["n_valid"]
        return cells, sec, threads, "3 whole steps (one per variable) of the same chain out of the reference's functions"


class GlocalWorkload:
    """c2g: G1. unit = one glocal gridding call on the 2000 x 2000 raster (fit of every macro area + per-area IDW + blend);
    ranks take different time steps."""
    kernel = "pb::idw_kernel<1,true> per macro area (+ pre_interpolation_batch_kernel, glocal_targets / glocal_blend)"
    ROOFLINE_ON_STEP_TIME = True

    def __init__(self, eng, w, rank, world, n_total):
        self.eng, self.w, self.var = eng, w, A.airTemperature
        self.dem = syn.make_dem(w["rows"], w["cols"])
        self.urban = syn.make_urban(w["rows"], w["cols"])
        self.base = syn.make_stations(self.dem, w["stations"], proxies=(self.dem, self.urban))
        self.s = syn.settings_multiple_detrending()
        self.s.useGlocalDetrending = 1
        self.areas = syn.make_macro_areas(self.dem, self.base, w["areas_side"], w["blend_cells"], w["margin_m"])
        self.pairs = float(sum(len(m) * (c.size // 2) for m, c in zip(self.areas.meteo_points, self.areas.area_cells)))
        self.area_cells = int(sum(c.size // 2 for c in self.areas.area_cells))
        self.rank, self.world = rank, world
        if eng is not None and getattr(eng, "h", None):
            self.dem_id, self.urb_id = eng.upload(self.dem), eng.upload(self.urban)
            self.cells_id = eng.macro_areas_upload(self.areas)       # static like the DEM: the glocal weight maps
            self.out = np.empty(self.dem.shape, dtype=np.float32)    # the host's output grid, page-locked once by its owner
            try:
                eng.host_register(self.out)
            except Exception:  # noqa: BLE001 - pageable output: slower download, same result
                pass

    def _step_inputs(self, k):
        st = self.base.copy()
        st.value[:] = (self.base.value + np.random.default_rng(700 + self.rank + k * self.world).normal(0, 0.3, st.n)).astype(np.float32)
        s = A.pb_settings.from_buffer_copy(bytes(self.s))
        syn.set_points_range(s, st)
        return st, s

    def _run(self, k, device_only):
        st, s = self._step_inputs(k)
        ok, grid, mn, mx, nv = self.eng.glocal_detrending_raster(st, s, self.var, self.areas, self.dem_id, self.dem.shape,
                                                                 (self.dem_id, self.urb_id), device_only=device_only,
                                                                 cells_id=self.cells_id, out=None if device_only else self.out)
        assert ok
        return nv

    def device_step(self, k):
        return self._run(k, True)

    def e2e_step(self, k):
        return self._run(k, False)

    def roofline(self, cells_per_launch, step_s, peaks):
        fp32, mufu = self.eng.measure_pipe_peaks()
        ach = 11.0 * self.pairs / step_s * 1e-12
        return {"bound": "fp32", "achieved": ach, "peak": fp32, "unit": "TFLOP/s", "frac": ach / fp32, "traffic": None,
                "kernel": self.kernel, "kernel_ms_per_launch": step_s * 1e3,
                "algorithmic": f"11 flop x {self.pairs:.4g} (cell, station) pairs per step = sum over areas of stations x cells "
                               f"({self.area_cells} (cell, area) entries for {int(cells_per_launch)} cells); taken against the whole "
                               "device time of the step (fit + 3 kernels per area + min/max)",
                "peak_source": "FFMA2 chain measured live on this GPU (MEASURED_PEAKS.json has no FP32 figure)",
                "mufu_bound": {"pairs_per_s": self.pairs / step_s, "mufu_peak_per_s": mufu * 1e9, "frac": self.pairs / step_s / (mufu * 1e9)}}

    def config(self):
        return {"unit_of_work": "one glocal gridding call (all macro areas) per step", "macro_areas": self.areas.n,
                "cell_area_entries": self.area_cells, "stations_per_area": [int(len(m)) for m in self.areas.meteo_points],
                "sharding": "time steps over ranks, no collective"}

    def cpu_sample(self, orc, seconds_hint):
        threads = orc.max_threads()
        st, s = self._step_inputs(0)
        ok, grid, mn, mx, nv, sec = orc.glocal_detrending_raster(st, s, self.var, self.areas, self.dem, (self.dem, self.urban), threads=0)
        return nv, sec, threads, (f"the whole {self.w['rows']}x{self.w['cols']} raster ({nv} valid cells), 1 pass; the reference "
                                  f"parallelises over the {self.areas.n} macro areas (project.cpp:3317)")


GENERIC_CLASSES = {"c3": LocalDetrendingWorkload, "c4": KrigingWorkload, "c5": HourlyBatchWorkload, "c2g": GlocalWorkload}
GENERIC_DTYPE = {"c3": "f64 (LM fits, IDW weights) / f32 distances", "c4": "f64 estimate / fp16x2-split tensor-core variance, f32 accumulate",
                 "c5": "f32 (f64 cross-tile accumulation, f64 fits)", "c2g": "f32 (f64 fits, f64 cross-tile accumulation and blend)"}
GENERIC_METRIC = {"c3": "gridded cells/sec (localDetrending + IDW)", "c4": "gridded cells/sec (ordinary kriging, estimate + RMSE)",
                  "c5": "gridded cells/sec (hourly batch: QC + detrending + IDW + LOO)",
                  "c2g": "gridded cells/sec (glocal detrending + IDW)"}


def run_generic(args, rank, local_rank, world):
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
    t0 = time.perf_counter()
    wl = GENERIC_CLASSES[args.workload](eng, w, rank, world, n_total)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for k in range(args.warmup):
        wl.device_step(k)
    barrier()
    dev_ms, kernel_ms, launches, cells = 0.0, 0.0, 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    for k in range(args.warmup, n_total):
        flush.zero_()
        e0.record()
        nv = wl.device_step(k)
        e1.record()
        e1.synchronize()
        dev_ms += e0.elapsed_time(e1)
        tm = eng.timing()
        kernel_ms += tm["kernel_ms"]
        launches += tm["launches"]
        cells += nv
    barrier()
    wall_s = time.perf_counter() - wall0

    for k in range(min(1, args.warmup)):
        wl.e2e_step(k)
    barrier()
    e2e_s, e2e_cells, h2d_b, d2h_b = 0.0, 0, 0, 0
    for k in range(args.warmup, n_total):
        flush.zero_()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nv = wl.e2e_step(k)
        e2e_s += time.perf_counter() - t1
        tm = eng.timing()
        h2d_b += tm["h2d_bytes"]
        d2h_b += tm["d2h_bytes"]
        e2e_cells += nv
    barrier()
    clocks = sampler.stop()

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, kernel_ms = tt.tolist()
        cc = torch.tensor([cells, e2e_cells, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        cells, e2e_cells, launches = cc.tolist()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        line = {
            "metric": GENERIC_METRIC[args.workload], "value": cells / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": GENERIC_DTYPE[args.workload], "data": "synthetic",
            "config": dict({"workload": f"{args.workload}: {w['desc']}", "stations": w["stations"],
                            "valid_cells_per_step": cells / (args.steps * world), "steps_per_rank": args.steps,
                            "l2": "flushed (256 MiB write) between timed iterations", "setup_s": setup_s}, **wl.config()),
            "e2e": {"value": e2e_cells / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b / args.steps),
                    "d2h_bytes_per_step": int(d2h_b / args.steps), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches), "wall_s_timed_region": wall_s, "clocks": clocks,
            # c5: many kernels on several lanes per step; its roofline is taken against the whole device time of the step
            "roofline": wl.roofline(cells / (args.steps * world),
                                    (dev_ms if getattr(wl, "ROOFLINE_ON_STEP_TIME", False) else kernel_ms) * 1e-3 / args.steps, peaks),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import binding
                kind = "reference" if binding.have_ref() else "port"
                orc = binding.Oracle(kind)
                n_cpu, sec, threads, sample = wl.cpu_sample(orc, 15.0)
                line["cpu_baseline"] = {"value": n_cpu / sec, "unit": UNIT, "cores": threads, "kind": kind, "seconds": sec,
                                        "sample": sample}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:120]}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference_generic(args, rank):
    """--impl reference for c3 / c4 / c5: the reference's CPU functions on a bounded sample per step."""
    if rank != 0:
        return
    from oracle import binding
    kind = "reference" if binding.have_ref() else "port"
    orc = binding.Oracle(kind)
    w = WORKLOADS[args.workload]

    class _NoEngine:                                           # the workload objects only need the engine for GPU steps
        def upload(self, r):
            return -1

        def kriging_setup(self, *a, **k):
            return -1

        def kriging_uses_tensor_path(self, k):
            return -1

    wl = GENERIC_CLASSES[args.workload](_NoEngine(), w, 0, 1, args.steps + args.warmup)
    cells, sec, threads, sample = 0, 0.0, 1, ""
    budget = max(2.0, 60.0 / max(1, args.steps + args.warmup))
    for k in range(args.warmup + args.steps):
        n_cpu, s_cpu, threads, sample = wl.cpu_sample(orc, budget)
        if k >= args.warmup:
            cells += n_cpu
            sec += s_cpu
    value = cells / sec
    line = {"impl": "reference", "metric": GENERIC_METRIC[args.workload], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "reference arithmetic (f32 / f64 mix)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample + " per step"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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
        # the reference arm runs on rank 0 alone and gets every host core
        share = 1 if args.impl == "reference" else world
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // share))

    if args.impl == "reference":
        if args.workload in GENERIC:
            run_reference_generic(args, rank)
        else:
            run_reference(args, rank, world)
        return
    if args.workload in GENERIC:
        run_generic(args, rank, local_rank, world)
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
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for k in range(args.warmup):
        eng.interpolation_raster_resident(steps[k], s, var, dem_id, (dem_id, urb_id), device_only=True)
    barrier()
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

    # ---------------- end to end through the drop-in call with host buffers (`e2e`) ----------------
    # first with pageable host buffers (every byte passes through the engine's pinned staging ring: extra context), then — the
    # `e2e` figure — with the three long-lived host buffers (DEM, proxy grid, output grid) page-locked once by their owner
    # (pb_host_register), which is what a host keeps loaded across time steps; the copies stay inside the timed region
    for k in range(min(2, args.warmup)):
        eng.interpolation_raster(steps[k], s, var, dem, (dem, urban), out=out_host)
    barrier()
    pageable_s = 0.0
    for k in range(args.warmup, n_total):
        flush.zero_()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        eng.interpolation_raster(steps[k], s, var, dem, (dem, urban), out=out_host)
        pageable_s += time.perf_counter() - t1
    barrier()
    registered = []
    try:
        for a in (dem.data, urban.data, out_host):
            eng.host_register(a)
            registered.append(a)
    except Exception as e:  # noqa: BLE001 - a box that refuses to page-lock: the call then stages every byte itself
        print(f"[bench] pb_host_register failed ({e}); e2e measured on pageable host buffers", file=sys.stderr)
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
    clocks = sampler.stop()
    for a in registered:
        eng.host_unregister(a)
    # cached-raster variant (DEM/proxies resident, stations H2D + output D2H per step), reported as extra context
    res_s = 0.0
    for k in range(args.warmup, n_total):
        t1 = time.perf_counter()
        eng.interpolation_raster_resident(steps[k], s, var, dem_id, (dem_id, urb_id), out=out_host)
        res_s += time.perf_counter() - t1
    barrier()

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s, res_s, kernel_ms, pageable_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, res_s, kernel_ms, pageable_s = tt.tolist()
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
                    "what": "pb_interpolation_raster with host DEM/proxy/station buffers and host output grid; the three raster "
                            "buffers are page-locked once by their owner (pb_host_register), every step copies them H2D / D2H"
                            if len(registered) == 3 else "pb_interpolation_raster with pageable host buffers (page-locking refused)",
                    "pageable_host_buffers_value": e2e_cells / pageable_s,
                    "resident_rasters_value": e2e_cells / res_s},
            "gpu_launches": int(launches),
            "wall_s_timed_region": wall_s,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp32_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at C2 from the ncu --set full capture
                         # profiles/r01_idw_c2b.ncu_summary.json (46.9 MB + 10.2 MB; algorithmic 12 B/cell = 44.4 MB + the
                         # 14.8 MB valid-cell list); other workloads: not captured
                         "traffic": 57.09e6 if args.workload == "c2" else None,
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
