#!/usr/bin/env python
"""bench.py — gridded cells/s of the interpolationRaster hot path on B200 (driver contract: one JSON line).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3idw|c3|c4|c5|c2g] [--impl reference] [--no-extras]

Primary workload (default "c2" = BASELINE.json configs[1]): synthetic 2000x2000 DEM @100 m, 500 stations, IDW with
multipleDetrending (elevation double-piecewise + urban fraction linear), airTemperature. A "step" = one time step = one
interpolationRaster call over the whole raster with that step's station values.

  value : valid cells gridded / device time, DEM + proxy rasters resident in HBM (pb_interpolation_raster_device).
  e2e   : the same through the drop-in C-ABI call with HOST buffers (pb_interpolation_raster): DEM + proxy H2D,
          station H2D, output raster D2H into host memory, all inside the timed region, every step.
  N > 1 : time steps are sharded over ranks (one process per GPU, no data-path collective); weak scaling.
  parity: before the line is printed a sample of the workload is gridded by the reference's own CPU code (oracle/_ref) and
          compared with the GPU result of the same inputs: {"max_abs", "cells", "mask_equal", "ok"} (bar 1e-4).
  --impl reference : the reference's own CPU interpolationRaster (oracle/_ref, built from the unmodified sources) with all
          host threads, each step a bounded row-band sample of the same workload.

With the default workload the line also carries an "extra" block with the other BASELINE.json configs, each measured by the
same harness (value / e2e / roofline / parity, cpu_baseline at N = 1) on fewer steps:
  c3idw : 10000x10000 DEM, 5000 stations, IDW + multipleDetrending (the shape the >= 50x target is quoted on);
  c3    : 10000x10000 DEM, 5000 stations, localDetrending + IDW (K2); a step = one band of rows, ROW BANDS sharded over ranks;
  c4    : 10000x10000 DEM, 2000 stations, ordinary kriging estimate + RMSE (K5); a step = one band of rows (N = 1 only);
  c5    : hourly batch on the 2000x2000 DEM, 500 stations: a step = 8 hours of PragaProject::interpolationMeteoGrid (per hour air
          temperature, relative humidity through the dew point, precipitation; each gridding = spatial QC + preInterpolation +
          raster + leave-one-out; every map aggregated onto a 5 km meteo grid on the device), TIME STEPS sharded over ranks.
At N > 1 the extra block holds c3 and c5: the two partitions BASELINE.json's north_star names.
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

UNIT = "cells/s"
TOL = 1e-4

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
                    "airRelHumidity through the dew point, precipitation; per gridding spatial quality control + preInterpolation + "
                    "raster + leave-one-out cross-validation, every map aggregated onto a 5 km meteo grid "
                    "(PragaProject::interpolationMeteoGrid / Project::interpolationDem / interpolationCv)"),
}
METRIC = {"c2": "gridded cells/sec (detrended IDW)", "c3idw": "gridded cells/sec (detrended IDW)",
          "tiny": "gridded cells/sec (detrended IDW)",
          "c3": "gridded cells/sec (localDetrending + IDW)", "c4": "gridded cells/sec (ordinary kriging, estimate + RMSE)",
          "c5": "gridded cells/sec (hourly batch: QC + detrending + IDW + LOO + meteo-grid aggregation)",
          "c2g": "gridded cells/sec (glocal detrending + IDW)"}
DTYPE = {"c2": "f32 (f64 cross-tile accumulation and retrend)", "c3idw": "f32 (f64 cross-tile accumulation and retrend)",
         "tiny": "f32 (f64 cross-tile accumulation and retrend)",
         "c3": "f64 (LM fits, IDW weights) / f32 distances", "c4": "f64 estimate / fp16x2-split tensor-core variance, f32 accumulate",
         "c5": "f32 (f64 cross-tile accumulation, f64 fits)", "c2g": "f32 (f64 fits, f64 cross-tile accumulation and blend)"}
SHARDING = {"c2": "time steps over ranks, no collective", "c3idw": "time steps over ranks, no collective",
            "tiny": "time steps over ranks, no collective", "c3": "row bands over ranks, no collective",
            "c4": "row bands over ranks, no collective", "c5": "time steps over ranks, no collective",
            "c2g": "time steps over ranks, no collective"}


def static_config(name):
    """the `config` object of the JSON line: a pure function of the workload name, printed identically by both arms"""
    w = WORKLOADS[name]
    cfg = {"workload": f"{name}: {w['desc']}", "raster": f"{w['rows']}x{w['cols']}", "stations": w["stations"],
           "sharding": SHARDING[name], "l2": "flushed (256 MiB write) between timed iterations"}
    if "band_rows" in w:
        cfg["unit_of_work"] = f"one band of {w['band_rows']} rows per step"
    return cfg


_RASTER_CACHE = {}


def cached(kind, rows, cols):
    """the 10k x 10k rasters are shared by c3idw / c3 / c4 (4 s of numpy each)"""
    key = (kind, rows, cols)
    if key not in _RASTER_CACHE:
        _RASTER_CACHE[key] = syn.make_dem(rows, cols) if kind == "dem" else syn.make_urban(rows, cols)
    return _RASTER_CACHE[key]


def build_workload(name, n_steps, first_step=0):
    """DEM, urban proxy, settings with an already fitted detrending model and per-step detrended station sets.

    interpolationRaster receives points that preInterpolation has already detrended and settings that carry the
    fitted model (project.cpp:3108-3150), so that is what a step is given. The fit itself is station-space work
    of a few ms; here the model parameters are fixed plausible values and the residuals are computed with numpy."""
    w = WORKLOADS[name]
    dem = cached("dem", w["rows"], w["cols"])
    urban = cached("urban", w["rows"], w["cols"])
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
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe). ONE sampler per node (rank 0)
    watches every GPU of the job: eight ranks each starting their own nvidia-smi loop cost host time inside the timed regions."""

    def __init__(self, n_gpus):
        self.n = n_gpus
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first(self, seconds=5.0):
        """nvidia-smi takes 0.1-1 s to print its first line: the timed regions of the fast workloads are shorter than that,
        so the caller waits here BEFORE the warm-up"""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < seconds:
            time.sleep(0.02)

    def mark(self):
        return len(self.lines)

    def stats(self, since=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[since:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                if int(f[0]) >= self.n:
                    continue
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "gpus_watched": self.n, "reasons": sorted(reasons),
                "window": "100 ms samples of every GPU of the job from this workload's warm-up through its device-timed and "
                          "end-to-end regions (one sampler per node)"}

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def _band_raster(dem, r0, r1):
    rows = dem.shape[0]
    return A.Raster(dem.data[r0:r1], dem.llx, dem.lly + (rows - r1) * dem.cell_size, dem.cell_size, dem.flag)


def _bands(dem, band_rows):
    """interior row bands of band_rows rows (the masked north margin is skipped)."""
    rows = dem.shape[0]
    first = max(1, rows // 100)
    return [(r, min(rows, r + band_rows)) for r in range(first, rows - band_rows + 1, band_rows)]


def grid_parity(gpu, cpu, flag):
    """{"max_abs", "cells", "mask_equal", "ok"} of two grids (bar: 1e-4 on every valid cell, NODATA masks equal)"""
    f = np.float32(flag)
    mg, mc = gpu == f, cpu == f
    mask_equal = bool(np.array_equal(mg, mc))
    both = ~mg & ~mc
    d = np.abs(gpu[both].astype(np.float64) - cpu[both].astype(np.float64))
    mx = float(d.max()) if d.size else 0.0
    return {"max_abs": mx, "cells": int(both.sum()), "mask_equal": mask_equal, "ok": bool(mask_equal and mx <= TOL),
            "tolerance": TOL}


class _NoEngine:
    """the reference arm builds the same workload objects without a GPU"""
    h = None

    def upload(self, r):
        return -1

    def free(self, rid):
        pass

    def kriging_setup(self, *a, **k):
        return -1

    def kriging_free(self, k):
        pass

    def kriging_uses_tensor_path(self, k):
        return -1

    def kriging_mma_count(self, k):
        return 0


# ---------------------------------------------------------------------------------------------------------------------
# workload objects: device_step / e2e_step / roofline / cpu_sample / parity, one harness (run_workload) for all of them
# ---------------------------------------------------------------------------------------------------------------------
class IdwWorkload:
    """c2 / c3idw: K1. unit = one time step over the whole raster; ranks take different time steps."""
    kernel = "pb::idw_persistent_kernel<2,1024>"

    def __init__(self, eng, w, rank, world, n_total, name="c2"):
        self.eng, self.w, self.name, self.var = eng, w, name, A.airTemperature
        self.rank, self.world = rank, world
        self.dem, self.urban, self.s, self.steps = build_workload(name, n_total, first_step=rank * n_total)
        self.gpu = getattr(eng, "h", None) is not None
        self.registered = []
        if self.gpu:
            t0 = time.perf_counter()
            self.dem_id, self.urb_id = eng.upload(self.dem), eng.upload(self.urban)
            self.n_valid = eng.valid_cells(self.dem_id)
            self.upload_s = time.perf_counter() - t0
            self.out_host = np.empty(self.dem.shape, dtype=np.float32)
            self.fp32_peak, self.mufu_peak = eng.measure_pipe_peaks()

    def device_step(self, k):
        ok, _, mn, mx, nv = self.eng.interpolation_raster_resident(self.steps[k], self.s, self.var, self.dem_id,
                                                                   (self.dem_id, self.urb_id), device_only=True)
        return nv

    def e2e_begin(self):
        """the three long-lived host buffers (DEM, proxy grid, output grid) are page-locked once by their owner
        (pb_host_register), which is what a host keeps loaded across time steps; the copies stay inside the timed region"""
        try:
            for a in (self.dem.data, self.urban.data, self.out_host):
                self.eng.host_register(a)
                self.registered.append(a)
        except Exception as e:  # noqa: BLE001 - a box that refuses to page-lock: the call then stages every byte itself
            print(f"[bench] pb_host_register failed ({e}); e2e measured on pageable host buffers", file=sys.stderr)

    def e2e_step(self, k):
        ok, _, mn, mx, nv = self.eng.interpolation_raster(self.steps[k], self.s, self.var, self.dem, (self.dem, self.urban),
                                                          out=self.out_host)
        return nv

    def e2e_end(self):
        for a in self.registered:
            self.eng.host_unregister(a)
        self.registered = []

    def e2e_what(self):
        if len(self.registered) == 3:
            return ("pb_interpolation_raster with host DEM/proxy/station buffers and host output grid; the three raster buffers are "
                    "page-locked once by their owner (pb_host_register), every step copies them H2D / D2H")
        return "pb_interpolation_raster with pageable host buffers (page-locking refused)"

    # ---- the drop-in as a PRAGA host calls it: pragab200::interpolationRaster (praga_b200/host/praga_b200_host.cpp) on
    # gis::Crit3DRasterGrid objects (float** rows) that live across time steps. oracle/_ref/libpraga_shim_test.so is only the
    # harness that owns those reference objects (it links the reference's gis / interpolation classes, which cannot be built
    # outside oracle/_ref); the code on the clock is the shim + libpraga_b200.so.
    def shim_open(self):
        from oracle import binding
        if not binding.have_shim():
            return False
        import ctypes as C
        self._C, self.shim = C, binding.load_shim()
        if self.shim.shim_session_set_device(int(os.environ.get("LOCAL_RANK", "0"))) != A.PB_OK:      # one GPU per rank
            return False
        d = self.dem.as_pb()
        g = (A.pb_raster * 2)()
        g[0], g[1] = d, self.urban.as_pb()
        self.shim_h = self.shim.shim_session_open(C.byref(self.s), C.byref(d), g, 2)
        return bool(self.shim_h)

    def shim_step(self, k):
        C = self._C
        stp = self.steps[k].as_pb()
        sec, mn, mx = C.c_double(0), C.c_float(0), C.c_float(0)
        rc = self.shim.shim_session_step(self.shim_h, C.byref(stp), self.var, 0, C.byref(sec), C.byref(mn), C.byref(mx))
        if rc != A.PB_OK:
            raise RuntimeError(f"pragab200::interpolationRaster failed ({rc})")
        t = A.pb_timing()
        self.shim.shim_session_timing(C.byref(t))
        return self.n_valid, sec.value, t.h2d_bytes, t.d2h_bytes

    def shim_parity(self, orc):
        """the session's output grid (float** rows) of step 0 against the reference CPU interpolationRaster on a row band"""
        C = self._C
        self.shim_step(0)
        o = A.pb_raster_out()
        o.data = A.fptr(self.out_host)
        self.shim.shim_session_output(self.shim_h, C.byref(o))
        rows = 16 if self.w["stations"] <= 1000 else 2
        r0, r1 = self._sample_band(rows)
        okc, gc, *_ = orc.interpolation_raster(self.steps[0], self.s, self.var, _band_raster(self.dem, r0, r1), (self.dem, self.urban), threads=0)
        return grid_parity(self.out_host[r0:r1], gc, self.dem.flag)

    def shim_close(self):
        self.shim.shim_session_close(self.shim_h)

    def e2e_context(self, k_range, sync):
        """extra end-to-end variants of the same step (context for the e2e figure): pageable caller buffers (every byte passes
        through the engine's pinned staging ring) and resident rasters (stations H2D + output D2H per step)"""
        out = {}
        for label, fn in (("pageable_host_buffers", lambda k: self.eng.interpolation_raster(self.steps[k], self.s, self.var, self.dem,
                                                                                           (self.dem, self.urban), out=self.out_host)),
                          ("resident_rasters", lambda k: self.eng.interpolation_raster_resident(self.steps[k], self.s, self.var,
                                                                                               self.dem_id, (self.dem_id, self.urb_id),
                                                                                               out=self.out_host))):
            fn(k_range[0])
            sync()
            t = 0.0
            for k in k_range:
                t1 = time.perf_counter()
                fn(k)
                t += time.perf_counter() - t1
            out[label] = t
        return out

    def roofline(self, cells_per_launch, kernel_s, step_s, peaks):
        S = self.w["stations"]
        pairs = cells_per_launch * S
        ach = 11.0 * pairs / kernel_s * 1e-12
        return {"bound": "fp32", "achieved": ach, "peak": self.fp32_peak, "unit": "TFLOP/s", "frac": ach / self.fp32_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at C2 from the ncu --set full capture
                # profiles/r01_idw_c2b.ncu_summary.json (46.9 MB + 10.2 MB; algorithmic 12 B/cell = 44.4 MB + the 14.8 MB
                # valid-cell list); other workloads: not captured
                "traffic": 56.70e6 if self.name == "c2" else None,   # dram read + write, ncu --set full (profiles/r02_idw_c2.ncu_summary.json)
                "kernel": self.kernel, "kernel_ms_per_launch": kernel_s * 1e3,
                "algorithmic": f"11 flop x {S} stations x {int(cells_per_launch)} cells per launch (SURVEY.md 8d)",
                "peak_source": "FFMA2 chain measured live on this GPU by the builder's own micro-kernel (pb_measure_pipe_peaks): "
                               "MEASURED_PEAKS.json (driver-written) has no FP32 figure",
                "mufu_bound": {"pairs_per_s": pairs / kernel_s, "mufu_peak_per_s": self.mufu_peak * 1e9,
                               "frac": pairs / kernel_s / (self.mufu_peak * 1e9)},
                "hbm": {"algorithmic_bytes_per_cell": 12, "achieved_gbs": 12.0 * cells_per_launch / kernel_s * 1e-9,
                        "peak_gbs": peaks.get("hbm_gbs")}}

    def details(self):
        return {"valid_cells_per_step": self.n_valid, "static_raster_upload_s": self.upload_s}

    def _sample_band(self, rows):
        r0 = (self.w["rows"] - rows) // 2
        return r0, r0 + rows

    def cpu_sample(self, orc, seconds_hint):
        """reference CPU path on a bounded row band: the pass is repeated until ~seconds_hint of CPU work are timed"""
        threads = orc.max_threads()
        target_cells = 1.2e6 * seconds_hint * 500 / max(1, self.w["stations"])   # ~1e6 cells/s at 500 stations on 16 threads
        rows = int(max(8, min(self.w["rows"], target_cells / self.w["cols"])))
        r0, r1 = self._sample_band(rows)
        band = _band_raster(self.dem, r0, r1)
        total, best, passes, nv = 0.0, None, 0, 0
        while total < seconds_hint and passes < 40:
            ok, grid, mn, mx, nv, sec = orc.interpolation_raster(self.steps[0], self.s, self.var, band, (self.dem, self.urban), threads=0)
            total += sec
            passes += 1
            best = sec if best is None else min(best, sec)
        return nv, best, threads, (f"{rows} full rows ({nv} valid cells) of the {self.w['rows']}x{self.w['cols']} raster, best of "
                                   f"{passes} passes ({total:.1f} s of CPU work), all {threads} host threads")

    def parity(self, orc):
        rows = 16 if self.w["stations"] <= 1000 else 2
        r0, r1 = self._sample_band(rows)
        band = _band_raster(self.dem, r0, r1)
        okc, gc, mnc, mxc, nvc, _ = orc.interpolation_raster(self.steps[0], self.s, self.var, band, (self.dem, self.urban), threads=0)
        self.out_host[r0:r1] = 0
        okg, _, mn, mx, nvg = self.eng.interpolation_raster_resident(self.steps[0], self.s, self.var, self.dem_id,
                                                                     (self.dem_id, self.urb_id), row_begin=r0, row_end=r1,
                                                                     out=self.out_host)
        p = grid_parity(self.out_host[r0:r1], gc, self.dem.flag)
        p["ok"] = bool(p["ok"] and nvc == nvg)
        p["sample"] = f"rows {r0}-{r1} of step 0 against the {orc.kind} CPU interpolationRaster"
        return p

    def close(self):
        if self.gpu:
            self.eng.free(self.urb_id)
            self.eng.free(self.dem_id)


class LocalDetrendingWorkload:
    """c3: K2. unit = a band of rows of the 10k x 10k raster; ranks take different bands (row-band sharding)."""
    prewarm_timed_units = True
    kernel = "pb::local_descent_kernel<0,11,1> (+ pb::local_select_kernel, pb::local_finish_kernel<0>: the three-kernel pipeline)"
    # FP64 flop per cell executed by the LM fits + estimate at this station density, counted with ncu
    # (thread-level DADD + DMUL + 2 DFMA of profiles/r01_local_k2e_flops.csv / 453 823 cells; data dependent)
    FP64_FLOP_PER_CELL = 3.64e5

    def __init__(self, eng, w, rank, world, n_total, name="c3"):
        self.eng, self.w, self.var = eng, w, A.airTemperature
        self.dem = cached("dem", w["rows"], w["cols"])
        self.urban = cached("urban", w["rows"], w["cols"])
        self.base = syn.make_stations(self.dem, w["stations"], proxies=(self.dem, self.urban))
        self.s = syn.settings_multiple_detrending()
        self.s.useLocalDetrending = 1
        self.s.minPointsLocalDetrending = 20
        syn.set_points_range(self.s, self.base)
        self.bands = _bands(self.dem, w["band_rows"])
        self.rank, self.world = rank, world
        self.gpu = getattr(eng, "h", None) is not None
        if self.gpu:
            self.dem_id, self.urb_id = eng.upload(self.dem), eng.upload(self.urban)
            self.out = np.empty(self.dem.shape, dtype=np.float32)

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

    def e2e_what(self):
        return "pb_local_detrending_raster with host DEM / proxy grids (pageable) and host output grid, one band of rows per call"

    def roofline(self, cells_per_launch, kernel_s, step_s, peaks):
        peak = self.eng.measure_fp64_peak()
        ach = self.FP64_FLOP_PER_CELL * cells_per_launch / kernel_s * 1e-12
        return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "kernel": self.kernel, "kernel_ms_per_launch": kernel_s * 1e3,
                "algorithmic": f"{self.FP64_FLOP_PER_CELL:.3g} FP64 flop per cell (ncu-counted at this station density, data "
                               f"dependent) x {int(cells_per_launch)} cells per launch",
                "peak_source": "DFMA chain measured live on this GPU by the builder's own micro-kernel (MEASURED_PEAKS.json has no "
                               "FP64 figure); the TU is built with -fmad=false like the reference's x86-64 build, so no DFMA is "
                               "formed: the reachable ceiling is half of this peak",
                "selection_floor": {"fp32_flop_per_cell": 6 * self.w["stations"],
                                    "note": "6*S FP32 flop per cell for the selection distances (SURVEY 8d)"}}

    def details(self):
        return {"bands": len(self.bands)}

    def cpu_sample(self, orc, seconds_hint):
        # ~2e-4 s per cell and thread in the reference: a few rows are ~10-30 s
        threads = orc.max_threads()
        rows = max(1, int(seconds_hint * threads / 2.2e-4 / (0.925 * self.w["cols"])))
        r0 = self.bands[len(self.bands) // 2][0]
        st, _, _ = self._step_inputs(0)
        ok, grid, mn, mx, nv, sec = orc.local_detrending_raster(st, self.s, self.var, self.dem, (self.dem, self.urban), threads=0,
                                                                row_begin=r0, row_end=r0 + rows)
        self._cpu_grid = (r0, r0 + rows, st, grid[r0:r0 + rows].copy())
        return nv, sec, threads, f"{rows} full rows ({nv} valid cells) of the {self.w['rows']}x{self.w['cols']} raster, 1 pass"

    def parity(self, orc):
        if getattr(self, "_cpu_grid", None) is not None:      # N = 1: the rows the CPU baseline leg just gridded
            r0, r1, st, gc = self._cpu_grid
        else:
            r0 = self.bands[len(self.bands) // 2][0]
            r1 = r0 + 1
            st, _, _ = self._step_inputs(0)
            ok, grid, mn, mx, nv, sec = orc.local_detrending_raster(st, self.s, self.var, self.dem, (self.dem, self.urban), threads=0,
                                                                    row_begin=r0, row_end=r1)
            gc = grid[r0:r1]
        self.out[r0:r1] = 0
        self.eng.interpolation_raster_resident(st, self.s, self.var, self.dem_id, (self.dem_id, self.urb_id), row_begin=r0,
                                               row_end=r1, out=self.out, local=True)
        p = grid_parity(self.out[r0:r1], gc, self.dem.flag)
        p["bit_identical"] = bool(np.array_equal(self.out[r0:r1], gc))
        p["sample"] = f"rows {r0}-{r1} against the {orc.kind} CPU loop of interpolationDemLocalDetrending"
        return p

    def close(self):
        if self.gpu:
            self.eng.free(self.urb_id)
            self.eng.free(self.dem_id)


class KrigingWorkload:
    """c4: K5. unit = a band of rows; estimate + RMSE grids."""
    prewarm_timed_units = True
    kernel = "pb::kriging_variance_pair_kernel<1> (+ kriging_estimate_fast_kernel<1>)"

    def __init__(self, eng, w, rank, world, n_total, name="c4"):
        self.eng, self.w = eng, w
        self.dem = cached("dem", w["rows"], w["cols"])
        n = w["stations"]
        rng = np.random.default_rng(1)
        self.pos = np.stack([self.dem.llx + rng.random(n) * w["cols"] * self.dem.cell_size,
                             self.dem.lly + rng.random(n) * w["rows"] * self.dem.cell_size], axis=1)
        self.val = rng.normal(12.0, 3.0, n)
        self.vario = (A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0)
        self.bands = _bands(self.dem, w["band_rows"])
        self.rank, self.world = rank, world
        self.gpu = getattr(eng, "h", None) is not None
        self.setup_s, self.tensor = 0.0, -1
        if self.gpu:
            t0 = time.perf_counter()
            self.kid = eng.kriging_setup(self.pos, self.val, *self.vario)
            self.setup_s = time.perf_counter() - t0
            self.tensor = eng.kriging_uses_tensor_path(self.kid)
            self.dem_id = eng.upload(self.dem)
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

    def e2e_what(self):
        return ("pb_kriging_raster: the DEM is resident by API (the call takes a raster id); per step both output grids of the band "
                "come down to host memory")

    def roofline(self, cells_per_launch, kernel_s, step_s, peaks):
        n = self.w["stations"]
        flop = 2.0 * n * (n + 1) + 4.0 * n
        ach = flop * cells_per_launch / kernel_s * 1e-12
        peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")
        out = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
               "kernel": self.kernel, "kernel_ms_per_launch": kernel_s * 1e3,
               "algorithmic": f"(2 n (n+1) + 4 n) flop per cell, n = {n} (SURVEY 8d: the reference's dense weight solve) x "
                              f"{int(cells_per_launch)} cells per launch; time = estimate + variance kernels",
               "peak_source": "dense bf16 GEMM, sustained, MEASURED_PEAKS.json"}
        # what the tensor cores really executed: the variance kernel issues 3 FP16 MMAs per K step (hi/lo operand split) on the
        # lower triangle of L^-1 and, with the bounded spherical model, only for the K stages whose 32 stations can be inside
        # the variogram range of the tile's cells (c = 0 exactly beyond it): counted on the device for one band
        self.eng.kriging_mma_count(self.kid)
        nv = self.device_step(0)
        mma = self.eng.kriging_mma_count(self.kid)
        if mma and nv:
            exe = mma * 2.0 * 256 * 256 * 16 / nv * cells_per_launch / kernel_s * 1e-12
            out["executed"] = {"tflops": exe, "frac_of_peak": exe / peak, "mma_256x256x16_per_256_cells": mma / (nv / 256.0),
                               "note": "tensor flops issued (tcgen05 cta_group::2 MMAs counted by the kernel); `achieved` counts "
                                       "the reference formulation's flops, most of which the kernel proves to be zero and skips, "
                                       "so `frac` can exceed what a dense evaluation could reach"}
        return out

    def details(self):
        return {"bands": len(self.bands), "kriging_setup_s": self.setup_s, "tensor_path": bool(self.tensor == 1)}

    def _sample_xy(self, m, seed=3):
        r0, r1 = self.bands[len(self.bands) // 2]
        rng = np.random.default_rng(seed)
        return np.stack([self.dem.llx + rng.random(m) * self.w["cols"] * self.dem.cell_size,
                         self.dem.lly + (self.w["rows"] - r1 + rng.random(m) * (r1 - r0)) * self.dem.cell_size], axis=1)

    def cpu_sample(self, orc, seconds_hint):
        # krigingSetWeight is O(n^2) per cell (~3 ms at n = 2000) and the reference keeps its system in file statics: 1 thread
        m = max(200, int(seconds_hint / 3.2e-3))
        xy = self._sample_xy(m)
        ok, res, rm, t_setup, t_eval = orc.kriging_points(self.pos, self.val, *self.vario, xy)
        self._cpu_points = (xy, res, rm)
        return m, t_eval, 1, (f"{m} random cells of one band (krigingSetWeight + krigingResult + krigingRMSE each), single "
                              f"thread: kriging.cpp keeps one system in file-static globals; krigingVariogram setup "
                              f"{t_setup:.1f} s not included")

    def parity(self, orc):
        if getattr(self, "_cpu_points", None) is not None:
            xy, res, rm = self._cpu_points
        else:
            xy = self._sample_xy(48)
            ok, res, rm, _, _ = orc.kriging_points(self.pos, self.val, *self.vario, xy)
        got, got_rm = self.eng.kriging_points(self.kid, xy)
        e1, e2 = float(np.abs(got - res).max()), float(np.abs(got_rm - rm).max())
        return {"max_abs": max(e1, e2), "max_abs_estimate": e1, "max_abs_rmse": e2, "cells": int(xy.shape[0]), "mask_equal": True,
                "ok": bool(max(e1, e2) <= TOL), "tolerance": TOL,
                "sample": f"{xy.shape[0]} random cells of one band against the {orc.kind} CPU krigingSetWeight / krigingResult / "
                          f"krigingRMSE (n = {self.w['stations']}, tensor path = {bool(self.tensor == 1)})"}

    def close(self):
        if self.gpu:
            self.eng.kriging_free(self.kid)
            self.eng.free(self.dem_id)


class HourlyBatchWorkload:
    """c5: the hourly driver, PragaProject::interpolationMeteoGrid with getMeteoGridUpscaleFromDem() (pragaProject.cpp:2951-2990),
    through pb_meteo_grid_step_batch: per hour air temperature (multipleDetrending), relative humidity through the dew point
    (temperature map sampled at the humidity stations -> dew point -> dew-point map -> relative humidity map) and precipitation;
    every gridded map is aggregated onto a 5 km meteo grid on the device and only those cell values come back; each gridding
    = spatial QC + preInterpolation + raster + leave-one-out. A step = HOURS hours = 3 * HOURS gridded rasters."""
    kernel = "pb::idw_persistent_kernel<2,1024> + station-space kernels (K3/K7, loo_exact, neighbourhood) + map operators"
    ROOFLINE_ON_STEP_TIME = True
    HOURS = int(os.environ.get("PB_BENCH_HOURS", "16"))
    GRID_CELL_M = 5000.0

    def __init__(self, eng, w, rank, world, n_total, name="c5"):
        self.eng, self.w = eng, w
        self.dem = cached("dem", w["rows"], w["cols"])
        self.urban = cached("urban", w["rows"], w["cols"])
        n = w["stations"]
        grids = (self.dem, self.urban)
        self.base = {A.airTemperature: syn.make_stations(self.dem, n, proxies=grids),
                     A.airRelHumidity: syn.make_stations(self.dem, n, proxies=grids, variable="humidity"),
                     A.precipitation: syn.make_stations(self.dem, n, proxies=grids, variable="precipitation")}
        self.rank, self.world = rank, world
        # 4 lanes measured best on one B200 (the raster kernels of the lanes run one after the other; more lanes only queue)
        # with several ranks on one node every rank's lane threads + its own thread should fit the rank's share of the cores
        # (they wait for the device by spinning; oversubscribed spinning threads were what collapsed c5 at 8 GPUs in r01)
        share = max(1, (os.cpu_count() or 1) // max(1, world))
        self.lanes = int(os.environ.get("PB_BENCH_LANES", str(max(1, min(4, share - 1)))))
        self.geom = self._geometry()
        self.n_cells = self.geom[3].size
        self.gpu = getattr(eng, "h", None) is not None
        if self.gpu:
            self.dem_id, self.urb_id = eng.upload(self.dem), eng.upload(self.urban)
            self.agg_id = eng.aggregation_upload(*self.geom)

    def _geometry(self):
        """aggregation points of a regular 5 km meteo grid over the DEM: the DEM cell centres inside each grid cell (what
        Crit3DMeteoGrid::findGridAggregationPoints builds, meteoGrid.cpp:655-760)"""
        rows, cols, cs = self.w["rows"], self.w["cols"], self.dem.cell_size
        k = int(self.GRID_CELL_M / cs)
        ny, nx = rows // k, cols // k
        off = (np.arange(k) + 0.5) * cs
        xs = self.dem.llx + (np.arange(nx)[:, None] * k * cs + off[None, :])           # nx x k
        ys = self.dem.lly + (np.arange(ny)[:, None] * k * cs + off[None, :])           # ny x k
        px = np.broadcast_to(xs[None, :, :, None], (ny, nx, k, k)).reshape(ny * nx, k * k)
        py = np.broadcast_to(ys[:, None, None, :], (ny, nx, k, k)).reshape(ny * nx, k * k)
        first = (np.arange(ny * nx + 1, dtype=np.int64) * (k * k))
        max_nr = np.full(ny * nx, k * k, dtype=np.int32)
        return np.ascontiguousarray(px.ravel()), np.ascontiguousarray(py.ravel()), first, max_nr

    def _settings(self, var):
        if var == A.airTemperature:
            s = syn.settings_multiple_detrending()
        else:
            s = A.default_settings()
            s.nProxy = 2
            s.proxy[0] = A.default_proxy(A.proxyHeight, 0)
            s.proxy[1] = A.default_proxy(A.proxyUrbanFraction, 1)
        return s, syn.settings_classic_detrending(2)

    def _hour_inputs(self, hour):
        """global hour index -> the three station sets of that hour + the humidity stations' own air temperature"""
        rng = np.random.default_rng(9000 + hour)
        anomaly = rng.normal(0, 0.6, self.w["stations"])
        t = self.base[A.airTemperature].copy()
        t.value[:] = (t.value + anomaly + 3.0 * np.sin(hour / 24 * 2 * np.pi)).astype(np.float32)
        rh = self.base[A.airRelHumidity].copy()
        rh.value[:] = np.clip(rh.value + 6 * anomaly, 1, 100).astype(np.float32)
        own_t = (15.0 - 0.0065 * rh.z + anomaly).astype(np.float32)
        own_t[rng.random(own_t.size) < 0.3] = np.float32(A.NODATA)          # humidity stations without a thermometer
        p = self.base[A.precipitation].copy()
        p.value[:] = np.maximum(p.value + 2 * anomaly, 0).astype(np.float32)
        return t, rh, own_t, p

    def _hour_batch(self, hour):
        t, rh, own_t, p = self._hour_inputs(hour)
        (s_t, q_t), (s_h, q_h), (s_p, q_p) = self._settings(A.airTemperature), self._settings(A.airDewTemperature), self._settings(A.precipitation)
        return [dict(meteo=t, settings=s_t, quality_settings=q_t, variable=A.airTemperature, aggr_method=A.PB_AGGR_AVERAGE),
                dict(meteo=rh, settings=s_h, quality_settings=q_h, variable=A.airRelHumidity, use_dew_point=True, air_temperature=own_t,
                     aggr_method=A.PB_AGGR_AVERAGE),
                dict(meteo=p, settings=s_p, quality_settings=q_p, variable=A.precipitation, aggr_method=A.PB_AGGR_AVERAGE)]

    def _run(self, k):
        # time steps over ranks: rank r takes the steps r, r + world, ... (a step = HOURS consecutive hours)
        h0 = (self.rank + k * self.world) * self.HOURS
        batch = [self._hour_batch(h0 + j) for j in range(self.HOURS)]
        res = self.eng.meteo_grid_step_batch(batch, self.dem_id, self.dem.shape, (self.dem_id, self.urb_id), aggregation_id=self.agg_id,
                                             n_cells=self.n_cells, lanes=self.lanes, cross_validate=True)
        return sum(v["n_valid"] for hour in res for v in hour)

    def device_step(self, k):
        return self._run(k)

    def e2e_step(self, k):
        return self._run(k)

    def e2e_what(self):
        return ("pb_meteo_grid_step_batch: DEM / proxy rasters and the aggregation geometry resident by API; per gridding the station "
                "arrays go up, the meteo-grid cell values, quality flags and leave-one-out residuals come down; the rasters stay in HBM "
                "(the same call as `value`: the hourly driver has no other form)")

    def roofline(self, cells_per_launch, kernel_s, step_s, peaks):
        fp32_peak, mufu_peak = self.eng.measure_pipe_peaks()
        S = self.w["stations"]
        ach = 11.0 * S * cells_per_launch / step_s * 1e-12
        return {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak, "traffic": None,
                "kernel": self.kernel, "kernel_ms_per_launch": step_s * 1e3,
                "algorithmic": f"11 flop x {S} stations x {int(cells_per_launch)} gridded cells per step (rasters only; the station-space "
                               "kernels and the map operators of the step are counted in the time, not in the flops; time = device "
                               "time of the whole step)",
                "peak_source": "FFMA2 chain measured live on this GPU by the builder's own micro-kernel"}

    def details(self):
        return {"hours_per_step": self.HOURS, "griddings_per_step": 3 * self.HOURS, "lanes": self.lanes, "meteo_grid_cells": int(self.n_cells),
                "aggregation_points": int(self.geom[0].size),
                "unit_of_work": f"{self.HOURS} hours per step through pb_meteo_grid_step_batch ({self.lanes} lanes); per hour: air "
                                "temperature, relative humidity via the dew point, precipitation; each gridding: QC + preInterpolation "
                                "+ raster + leave-one-out, then the map is aggregated onto the 5 km meteo grid"}

    def _reference_hour(self, orc, hour):
        """the same hour out of the reference's own functions (oracle/_ref)"""
        from oracle import binding
        import ctypes as C
        dem, grids = self.dem, (self.dem, self.urban)
        x, y, first, mx = self.geom
        t, rh, own_t, p = self._hour_inputs(hour)
        (s_t, q_t), (s_h, q_h), (s_p, q_p) = self._settings(A.airTemperature), self._settings(A.airDewTemperature), self._settings(A.precipitation)
        r_t = orc.interpolation_dem(t, s_t, A.airTemperature, dem, grids, quality_settings=q_t, cross_validate=True)
        map_t = A.Raster(r_t["grid"], dem.llx, dem.lly, dem.cell_size, dem.flag)
        cells_t = orc.spatial_aggregate_meteo_grid(map_t, x, y, first, mx, A.PB_AGGR_AVERAGE)
        air = own_t.copy()
        need = np.flatnonzero((rh.value != np.float32(A.NODATA)) & (own_t == np.float32(A.NODATA)))
        lib = binding.load_ref_sample()
        rr = map_t.as_pb()
        sx, sy = np.ascontiguousarray(rh.x[need]), np.ascontiguousarray(rh.y[need])
        val, ing = np.empty(need.size, dtype=np.float32), np.empty(need.size, dtype=np.uint8)
        lib.ref_raster_sample_points(C.byref(rr), A.dptr(sx), A.dptr(sy), need.size, A.fptr(val), A.u8ptr(ing))
        air[need[ing.astype(bool)]] = val[ing.astype(bool)]
        tdew = orc.tdew_from_relhum(rh.value, air)
        have = tdew != np.float32(A.NODATA)
        dew = rh.subset(have)
        dew.value[:] = tdew[have]
        r_d = orc.interpolation_dem(dew, s_h, A.airDewTemperature, dem, grids, quality_settings=q_h, cross_validate=True)
        ok, map_rh, mn, mxv = orc.relative_humidity_map(map_t, A.Raster(r_d["grid"], dem.llx, dem.lly, dem.cell_size, dem.flag))
        cells_h = orc.spatial_aggregate_meteo_grid(A.Raster(map_rh, dem.llx, dem.lly, dem.cell_size, dem.flag), x, y, first, mx, A.PB_AGGR_AVERAGE)
        r_p = orc.interpolation_dem(p, s_p, A.precipitation, dem, grids, quality_settings=q_p, cross_validate=True)
        cells_p = orc.spatial_aggregate_meteo_grid(A.Raster(r_p["grid"], dem.llx, dem.lly, dem.cell_size, dem.flag), x, y, first, mx,
                                                   A.PB_AGGR_AVERAGE)
        n_valid = r_t["n_valid"] + r_d["n_valid"] + r_p["n_valid"]
        return n_valid, dict(cells=(cells_t, cells_h, cells_p), have=have, residual=(r_t["residual"], r_d["residual"], r_p["residual"]))

    def cpu_sample(self, orc, seconds_hint):
        threads = orc.max_threads()
        t0 = time.perf_counter()
        n_valid, self._cpu_hour = self._reference_hour(orc, 0)
        sec = time.perf_counter() - t0
        return n_valid, sec, threads, ("1 whole hour (3 griddings of the 2000x2000 raster + humidity chain + 3 aggregations) of the same "
                                       "chain out of the reference's functions")

    def parity(self, orc):
        ref_hour = getattr(self, "_cpu_hour", None)
        if ref_hour is None:
            _, ref_hour = self._reference_hour(orc, 0)
        res = self.eng.meteo_grid_step_batch([self._hour_batch(0)], self.dem_id, self.dem.shape, (self.dem_id, self.urb_id),
                                             aggregation_id=self.agg_id, n_cells=self.n_cells, lanes=1, cross_validate=True)[0]
        nod = np.float32(A.NODATA)
        worst, cells, ok = 0.0, 0, True
        for k in range(3):
            a, b = res[k]["cells"], ref_hour["cells"][k]
            ok = ok and bool(np.array_equal(a == nod, b == nod)) and res[k]["ok"]
            m = b != nod
            if m.any():
                worst = max(worst, float(np.abs(a[m].astype(np.float64) - b[m]).max()))
            cells += int(m.sum())
        worst_res = 0.0
        for k in range(3):
            a, b = res[k]["residual"], ref_hour["residual"][k]
            if k == 1:
                a = a[ref_hour["have"]]
            ok = ok and bool(np.array_equal(a == nod, b == nod))
            m = b != nod
            if m.any():
                worst_res = max(worst_res, float(np.abs(a[m].astype(np.float64) - b[m]).max()))
        return {"max_abs": worst, "cells": cells, "mask_equal": ok, "ok": bool(ok and worst <= TOL and worst_res <= TOL), "tolerance": TOL,
                "max_abs_residual": worst_res,
                "sample": f"1 whole hour against the {orc.kind} CPU chain: the meteo-grid cell values of the three maps "
                          "(2000x2000 rasters gridded, humidity through the dew point) and the leave-one-out residuals"}

    def close(self):
        if self.gpu:
            self.eng.aggregation_free(self.agg_id)
            self.eng.free(self.urb_id)
            self.eng.free(self.dem_id)


class GlocalWorkload:
    """c2g: G1. unit = one glocal gridding call on the 2000 x 2000 raster (fit of every macro area + per-area IDW + blend);
    ranks take different time steps."""
    kernel = "pb::idw_kernel<1,true> per macro area (+ pre_interpolation_coop_kernel, glocal_targets / glocal_blend)"
    ROOFLINE_ON_STEP_TIME = True

    def __init__(self, eng, w, rank, world, n_total, name="c2g"):
        self.eng, self.w, self.var = eng, w, A.airTemperature
        self.dem = cached("dem", w["rows"], w["cols"])
        self.urban = cached("urban", w["rows"], w["cols"])
        self.base = syn.make_stations(self.dem, w["stations"], proxies=(self.dem, self.urban))
        self.s = syn.settings_multiple_detrending()
        self.s.useGlocalDetrending = 1
        self.areas = syn.make_macro_areas(self.dem, self.base, w["areas_side"], w["blend_cells"], w["margin_m"])
        self.pairs = float(sum(len(m) * (c.size // 2) for m, c in zip(self.areas.meteo_points, self.areas.area_cells)))
        self.area_cells = int(sum(c.size // 2 for c in self.areas.area_cells))
        self.rank, self.world = rank, world
        self.gpu = getattr(eng, "h", None) is not None
        if self.gpu:
            self.dem_id, self.urb_id = eng.upload(self.dem), eng.upload(self.urban)
            self.cells_id = eng.macro_areas_upload(self.areas)       # static like the DEM: the glocal weight maps
            self.out = np.empty(self.dem.shape, dtype=np.float32)    # the host's output grid, page-locked once by its owner
            self.registered = False
            try:
                eng.host_register(self.out)
                self.registered = True
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

    def e2e_what(self):
        return "pb_glocal_detrending_raster_resident: rasters and area cells resident, stations up, output grid down per step"

    def roofline(self, cells_per_launch, kernel_s, step_s, peaks):
        fp32, mufu = self.eng.measure_pipe_peaks()
        ach = 11.0 * self.pairs / step_s * 1e-12
        return {"bound": "fp32", "achieved": ach, "peak": fp32, "unit": "TFLOP/s", "frac": ach / fp32, "traffic": None,
                "kernel": self.kernel, "kernel_ms_per_launch": step_s * 1e3,
                "algorithmic": f"11 flop x {self.pairs:.4g} (cell, station) pairs per step = sum over areas of stations x cells "
                               f"({self.area_cells} (cell, area) entries for {int(cells_per_launch)} cells); taken against the whole "
                               "device time of the step (fit + kernels per area + min/max)",
                "peak_source": "FFMA2 chain measured live on this GPU by the builder's own micro-kernel",
                "mufu_bound": {"pairs_per_s": self.pairs / step_s, "mufu_peak_per_s": mufu * 1e9, "frac": self.pairs / step_s / (mufu * 1e9)}}

    def details(self):
        return {"macro_areas": self.areas.n, "cell_area_entries": self.area_cells,
                "stations_per_area": [int(len(m)) for m in self.areas.meteo_points]}

    def cpu_sample(self, orc, seconds_hint):
        threads = orc.max_threads()
        st, s = self._step_inputs(0)
        ok, grid, mn, mx, nv, sec = orc.glocal_detrending_raster(st, s, self.var, self.areas, self.dem, (self.dem, self.urban), threads=0)
        self._cpu_grid = grid
        return nv, sec, threads, (f"the whole {self.w['rows']}x{self.w['cols']} raster ({nv} valid cells), 1 pass; the reference "
                                  f"parallelises over the {self.areas.n} macro areas (project.cpp:3317)")

    def parity(self, orc):
        gc = getattr(self, "_cpu_grid", None)
        st, s = self._step_inputs(0)
        if gc is None:
            ok, gc, mn, mx, nv, sec = orc.glocal_detrending_raster(st, s, self.var, self.areas, self.dem, (self.dem, self.urban), threads=0)
            st, s = self._step_inputs(0)
        out = np.empty(self.dem.shape, dtype=np.float32)
        self.eng.glocal_detrending_raster(st, s, self.var, self.areas, self.dem_id, self.dem.shape, (self.dem_id, self.urb_id),
                                          cells_id=self.cells_id, out=out)
        p = grid_parity(out, gc, self.dem.flag)
        p["sample"] = f"the whole raster of step 0 against the {orc.kind} CPU loop of interpolationDemGlocalDetrending"
        return p

    def close(self):
        if self.gpu:
            if self.registered:
                self.eng.host_unregister(self.out)
            self.eng.free(self.urb_id)
            self.eng.free(self.dem_id)


CLASSES = {"c2": IdwWorkload, "c3idw": IdwWorkload, "tiny": IdwWorkload, "c3": LocalDetrendingWorkload, "c4": KrigingWorkload,
           "c5": HourlyBatchWorkload, "c2g": GlocalWorkload}
# extras run on fewer steps (warm-up stays >= 3): they are context for the primary line, not the driver's headline
EXTRA_STEPS = {"c3idw": 4, "c3": 3, "c4": 4, "c5": 4}
CPU_SECONDS = {"c2": 10.0, "c3idw": 8.0, "tiny": 1.0, "c3": 10.0, "c4": 8.0, "c5": 0.0, "c2g": 0.0}


def get_oracle():
    from oracle import binding
    kind = "reference" if binding.have_ref() else "port"
    return binding.Oracle(kind), kind


BLOCKING_SYNC_FOR = set()          # workloads that run with blocking synchronisation (filled by main)


def run_workload(name, eng, torch, dist, rank, world, steps, warmup, sampler, flush, want_cpu):
    eng.set_blocking_sync(name in BLOCKING_SYNC_FOR)
    try:
        return _run_workload(name, eng, torch, dist, rank, world, steps, warmup, sampler, flush, want_cpu)
    finally:
        eng.set_blocking_sync(False)


def _run_workload(name, eng, torch, dist, rank, world, steps, warmup, sampler, flush, want_cpu):
    """one workload through the contract: W untimed warm-up steps, K timed steps (CUDA events on the launching stream, L2
    flushed between iterations), barrier + synchronize on both sides, max over ranks; then the same K steps end to end."""
    n_total = warmup + steps
    t0 = time.perf_counter()
    wl = CLASSES[name](eng, WORKLOADS[name], rank, world, n_total, name=name)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mark = sampler.mark() if sampler else 0
    for k in range(warmup):
        wl.device_step(k)
    if getattr(wl, "prewarm_timed_units", False):
        # row-band workloads: every band has its own valid-cell count, so the engine's per-call buffers grow (cudaMalloc) the
        # first time a larger band comes by; the bands of the timed steps are gridded once, untimed, like a session's second pass
        for k in range(warmup, n_total):
            wl.device_step(k)
    barrier()
    dev_ms, kernel_ms, launches, cells = 0.0, 0.0, 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    for k in range(warmup, n_total):
        flush.zero_()                                   # L2 flush between timed iterations (not timed)
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

    # ---------------- end to end through the C-ABI call with host buffers ----------------
    if hasattr(wl, "e2e_begin"):
        wl.e2e_begin()
    for k in range(min(2, warmup)):
        wl.e2e_step(k)
    barrier()
    e2e_s, e2e_cells, h2d_b, d2h_b = 0.0, 0, 0, 0
    for k in range(warmup, n_total):
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
    e2e_what = wl.e2e_what()
    if hasattr(wl, "e2e_end"):
        wl.e2e_end()
    ctx_s = {}
    if name == "c2":
        ctx_s = wl.e2e_context(range(warmup, n_total), torch.cuda.synchronize)
        barrier()
    # ---------------- end to end through the C++ drop-in itself (reference objects, float** rows, resident raster cache) ----
    shim_s, shim_cells, shim_h2d, shim_d2h, shim_first_s, shim_ok = 0.0, 0, 0, 0, 0.0, 0.0
    if hasattr(wl, "shim_open") and not os.environ.get("PB_BENCH_NO_SHIM"):
        opened = False
        try:
            if wl.shim_open():
                _, shim_first_s, _, _ = wl.shim_step(0)            # first call of the session: uploads DEM + proxy grid
                for k in range(1, min(3, warmup)):
                    wl.shim_step(k)
                opened = True
        except Exception as e:  # noqa: BLE001
            print(f"[bench] shim leg unavailable: {e}", file=sys.stderr)
        barrier()
        if opened:
            try:
                for k in range(warmup, n_total):
                    flush.zero_()
                    torch.cuda.synchronize()
                    nv, sec, hb, db = wl.shim_step(k)
                    shim_s += sec
                    shim_cells += nv
                    shim_h2d += hb
                    shim_d2h += db
                shim_ok = 1.0
            except Exception as e:  # noqa: BLE001
                print(f"[bench] shim leg failed: {e}", file=sys.stderr)
        barrier()
    clocks = sampler.stats(mark) if sampler else None

    if world > 1:
        keys = sorted(ctx_s)
        tt = torch.tensor([dev_ms, e2e_s, kernel_ms, shim_s, shim_first_s, -shim_ok] + [ctx_s[k] for k in keys], dtype=torch.float64,
                          device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        vals = tt.tolist()
        dev_ms, e2e_s, kernel_ms, shim_s, shim_first_s = vals[:5]
        shim_ok = -vals[5]                              # 1 only if the leg ran on every rank
        ctx_s = dict(zip(keys, vals[6:]))
        cc = torch.tensor([cells, e2e_cells, launches, h2d_b, d2h_b, shim_cells, shim_h2d, shim_d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        cells, e2e_cells, launches, h2d_b, d2h_b, shim_cells, shim_h2d, shim_d2h = cc.tolist()
        h2d_b /= world                                  # bytes per step of ONE rank's call
        d2h_b /= world
        shim_h2d /= world
        shim_d2h /= world

    res = None
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        per_launch_cells = cells / (steps * world)
        on_step = getattr(wl, "ROOFLINE_ON_STEP_TIME", False)
        res = {
            "metric": METRIC[name], "value": cells / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE[name], "data": "synthetic", "config": static_config(name),
            "details": dict({"valid_cells_per_step": per_launch_cells, "steps_per_rank": steps, "setup_s": setup_s,
                             "peaks_file": peak_kind}, **wl.details()),
            "e2e": {"value": e2e_cells / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b / steps),
                    "d2h_bytes_per_step": int(d2h_b / steps), "ms_per_step": 1e3 * e2e_s / steps, "what": e2e_what},
            "gpu_launches": int(launches), "wall_s_timed_region": wall_s, "clocks": clocks,
            "roofline": wl.roofline(per_launch_cells, kernel_ms * 1e-3 / steps, dev_ms * 1e-3 / steps, peaks),
        }
        if on_step:
            res["roofline"]["note"] = "taken against the device time of the whole step (many kernels on several lanes per step)"
        for k, v in ctx_s.items():
            res["e2e"][k + "_value"] = e2e_cells / v
        if shim_ok > 0 and shim_s > 0:
            # the headline end-to-end figure is the drop-in as the reference's host calls it; the C-ABI call that uploads the
            # rasters every step stays beside it
            c_abi = res["e2e"]
            res["e2e"] = {
                "value": shim_cells / shim_s, "unit": UNIT, "h2d_bytes_per_step": int(shim_h2d / steps),
                "d2h_bytes_per_step": int(shim_d2h / steps), "ms_per_step": 1e3 * shim_s / steps,
                "what": "pragab200::interpolationRaster (the C++ drop-in with the reference's signature) on gis::Crit3DRasterGrid "
                        "objects (float** rows) that live across time steps, as in a PRAGA session: the DEM / proxy grids are kept "
                        "resident by the shim's raster cache after the session's first call, every step flattens that step's "
                        "Crit3DInterpolationDataPoint vector + settings, copies the stations up and the output grid down into "
                        "the caller's float** rows (pinned staging ring + host scatter), all inside the timed region",
                "first_call_ms": 1e3 * shim_first_s,
                "c_abi_host_rasters": {k: v for k, v in c_abi.items() if k != "unit"}}
        # ---------------- CPU baseline (N = 1 only) and the parity gate (every N, rank 0) ----------------
        try:
            orc, kind = get_oracle()
            if want_cpu and world == 1:
                n_cpu, sec, threads, sample = wl.cpu_sample(orc, CPU_SECONDS.get(name, 8.0))
                res["cpu_baseline"] = {"value": n_cpu / sec, "unit": UNIT, "cores": threads, "kind": kind, "seconds": sec,
                                       "sample": sample}
            res["parity"] = wl.parity(orc)
            if shim_ok > 0 and hasattr(wl, "shim_parity"):
                ps = wl.shim_parity(orc)
                res["parity"]["shim_max_abs"] = ps["max_abs"]
                res["parity"]["ok"] = bool(res["parity"]["ok"] and ps["ok"])
        except Exception as e:  # noqa: BLE001
            res.setdefault("cpu_baseline", {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:160]})
            res["parity"] = {"ok": False, "error": f"{type(e).__name__}: {e}"[:200]}
    if world > 1:
        dist.barrier()                                  # the other ranks wait for rank 0's CPU legs before the next workload
    if getattr(wl, "shim_h", None):
        wl.shim_close()
    wl.close()
    return res


def run_reference(args, rank):
    """--impl reference: the unmodified reference's CPU code on the box's host cores, rank 0 only, each step a bounded
    sample of the workload (the same sample the GPU arm's cpu_baseline leg times)."""
    if rank != 0:
        return
    orc, kind = get_oracle()
    name = args.workload
    wl = CLASSES[name](_NoEngine(), WORKLOADS[name], 0, 1, 1, name=name)
    cells, sec, threads, sample = 0, 0.0, 1, ""
    # the whole run ends within a few minutes: ~60 s of CPU work spread over the steps
    budget = min(CPU_SECONDS.get(name, 8.0) or 8.0, max(1.0, 60.0 / max(1, args.steps + args.warmup)))
    for k in range(args.warmup + args.steps):
        n_cpu, s_cpu, threads, sample = wl.cpu_sample(orc, budget)
        if k >= args.warmup:
            cells += n_cpu
            sec += s_cpu
    value = cells / sec
    line = {"impl": "reference", "metric": METRIC[name], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "reference arithmetic (f32 / f64 mix)", "data": "synthetic",
            "config": static_config(name),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample + ", per step"},
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
    ap.add_argument("--no-extras", action="store_true", help="primary workload only (no `extra` block)")
    ap.add_argument("--extras", default=None, help="comma-separated workloads for the `extra` block (default: c3idw,c3,c4,c5 at "
                                                   "N = 1; c3,c5 at N > 1)")
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
        if args.impl != "reference":
            os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # idle OpenMP workers sleep: the cores belong to the lane threads

    if args.impl == "reference":
        run_reference(args, rank)
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
    # c5 only (run_workload): with more waiting host threads (ranks x lanes of the batch driver + the ranks' own) than cores
    # the lane threads sleep in their synchronisations instead of spinning; the single-threaded workloads keep spinning
    # (a blocking wake-up sits inside every device-timed step: it cost c2 0.35 ms per step at 8 ranks)
    lanes_c5 = int(os.environ.get("PB_BENCH_LANES", str(max(1, min(4, max(1, (os.cpu_count() or 1) // world) - 1)))))
    blocking_for_c5 = world * (lanes_c5 + 1) > (os.cpu_count() or 1)
    if blocking_for_c5:
        BLOCKING_SYNC_FOR.add("c5")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    sampler = None
    if rank == 0:
        sampler = ClockSampler(world)
        sampler.start()
        sampler.wait_first()
    if world > 1:
        dist.barrier()

    line = run_workload(args.workload, eng, torch, dist, rank, world, args.steps, args.warmup, sampler, flush,
                        want_cpu=not args.no_cpu_baseline)
    if args.workload == "c2" and not args.no_extras:
        names = args.extras.split(",") if args.extras else (["c3idw", "c3", "c4", "c5"] if world == 1 else ["c3", "c5"])
        extra = {}
        for nme in names:
            nme = nme.strip()
            if nme not in CLASSES or nme == "c2":
                continue
            try:
                r = run_workload(nme, eng, torch, dist, rank, world, EXTRA_STEPS.get(nme, 4), max(3, min(args.warmup, 3)), sampler,
                                 flush, want_cpu=not args.no_cpu_baseline)
            except Exception as e:  # noqa: BLE001 - an extra must never cost the primary line
                if world > 1:
                    raise
                r = {"error": f"{type(e).__name__}: {e}"[:300]}
            if rank == 0:
                extra[nme] = r
        if rank == 0:
            line["extra"] = extra
    if rank == 0:
        if sampler:
            sampler.stop()
        print(json.dumps(line), flush=True)

    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
