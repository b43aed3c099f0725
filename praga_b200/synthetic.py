"""Seeded synthetic scenarios of the shapes BASELINE.json names: DEM, proxy rasters, stations.

Pure numpy data generation, shared by the GPU harness, the tests and the CPU oracle so both sides see the
same bytes. Nothing here computes an interpolation.
"""
import numpy as np

from . import _abi as A

LLX, LLY = 500000.0, 4800000.0
FLAG = -9999.0


def make_dem(rows, cols, cell_size=100.0, seed=7, nodata_frac=0.05, border=True):
    """DEM(r,c) = 200 + 1500*(0.5+0.5*sin(0.003 r)*cos(0.004 c)) + multi-octave noise, ~5 % flag cells and a
    masked border polygon (exercises the NODATA mask)."""
    rng = np.random.default_rng(seed)
    r = np.arange(rows, dtype=np.float32)[:, None]
    c = np.arange(cols, dtype=np.float32)[None, :]
    z = 200.0 + 1500.0 * (0.5 + 0.5 * np.sin(0.003 * r) * np.cos(0.004 * c))
    for k, (fr, fc, amp) in enumerate(((0.021, 0.017, 120.0), (0.053, 0.047, 60.0), (0.131, 0.119, 25.0))):
        ph = rng.uniform(0, 2 * np.pi, size=2)
        z = z + amp * np.sin(fr * r + ph[0]) * np.sin(fc * c + ph[1])
    z = z.astype(np.float32)
    z += rng.normal(0.0, 5.0, size=(rows, cols)).astype(np.float32)
    if nodata_frac > 0:
        z[rng.random((rows, cols), dtype=np.float32) < nodata_frac] = FLAG
    if border:
        # diagonal corner cut + a margin, like a region boundary inside its bounding box
        rr = np.arange(rows)[:, None]
        cc = np.arange(cols)[None, :]
        z[(rr + cc) < 0.12 * (rows + cols) / 2] = FLAG
        m = max(1, rows // 100)
        z[:m, :] = FLAG
        z[:, -m:] = FLAG
    return A.Raster(z, LLX, LLY, cell_size, FLAG)


def make_urban(rows, cols, cell_size=100.0, seed=11, llx=LLX, lly=LLY, nodata_frac=0.0):
    """Urban-fraction proxy: smooth field in [0, 1]; pass a coarser cell size for a foreign-geometry grid."""
    rng = np.random.default_rng(seed)
    r = np.arange(rows, dtype=np.float32)[:, None] * np.float32(cell_size / 100.0)
    c = np.arange(cols, dtype=np.float32)[None, :] * np.float32(cell_size / 100.0)
    ph = rng.uniform(0, 2 * np.pi, size=4)
    u = 0.5 + 0.3 * np.sin(0.0041 * r + ph[0]) * np.cos(0.0037 * c + ph[1]) + 0.2 * np.sin(0.019 * r + ph[2]) * np.sin(0.023 * c + ph[3])
    u = np.clip(u, 0.0, 1.0).astype(np.float32)
    if nodata_frac > 0:
        u[rng.random((rows, cols), dtype=np.float32) < nodata_frac] = FLAG
    return A.Raster(u, llx, lly, cell_size, FLAG)


def sample_grid(raster, x, y):
    """gis::getValueFromXY (gis.cpp:533-546) vectorised: nearest cell by truncation, flag outside."""
    rows, cols = raster.shape
    inv = 1.0 / raster.cell_size
    r = ((y - raster.lly) * inv).astype(np.int64)
    row = (rows - 1) - r
    col = ((x - raster.llx) * inv).astype(np.int64)
    ok = (row >= 0) & (row < rows) & (col >= 0) & (col < cols)
    out = np.full(x.shape, raster.flag, dtype=np.float32)
    out[ok] = raster.data[row[ok], col[ok]]
    return out


def make_stations(dem, n, proxies=(), variable="temperature", seed=42, supplemental_frac=0.0):
    """Stations uniform in the DEM bounding box; z = DEM at the station (500 m where the DEM is flag);
    proxy values sampled from the proxy rasters (first proxy = elevation -> z). Value models:
    temperature: 15 - 0.0065 z + 2 urban + N(0, 0.5); humidity: clip(70 + N(0,10)); precipitation: gamma-like, ~40 % zeros."""
    rng = np.random.default_rng(seed)
    rows, cols = dem.shape
    x = dem.llx + rng.random(n) * cols * dem.cell_size
    y = dem.lly + rng.random(n) * rows * dem.cell_size
    z = sample_grid(dem, x, y).astype(np.float64)
    z[z == dem.flag] = 500.0
    pv = np.full((n, len(proxies)), float(A.NODATA), dtype=np.float32)
    urban = np.zeros(n)
    for k, p in enumerate(proxies):
        if p is dem:
            pv[:, k] = z
        else:
            v = sample_grid(p, x, y)
            pv[:, k] = np.where(v == p.flag, float(A.NODATA), v)
            urban = np.where(v == p.flag, 0.0, v)
    if variable == "temperature":
        value = 15.0 - 0.0065 * z + 2.0 * urban + rng.normal(0, 0.5, n)
    elif variable == "humidity":
        value = np.clip(70.0 + rng.normal(0, 10.0, n), 0, 100)
    elif variable == "precipitation":
        value = np.where(rng.random(n) < 0.4, 0.0, rng.gamma(0.8, 6.0, n))
    else:
        raise ValueError(variable)
    codes = None
    if supplemental_frac > 0:
        u = rng.random(n)
        codes = np.where(u < supplemental_frac, A.supplemental, np.where(u < 2 * supplemental_frac, A.secondary, A.primary))
    return A.Stations(x, y, z, value, pv, codes)


def settings_multiple_detrending(use_urban=True):
    """C2 settings: idw + useMultipleDetrending, proxies elevation (double_piecewise, ranges
    {0,-30,0.002,-0.01 | 2500,50,0.007,-0.0015}, first guess {0,1,1,1}) + urbanFraction (linear {-10,-50 | 11,50})."""
    s = A.default_settings()
    s.useMultipleDetrending = 1
    s.nProxy = 2 if use_urban else 1
    el = A.default_proxy(A.proxyHeight, 0)
    el.fittingFunction = A.piecewiseTwo
    el.nFitParams = 4
    for j, (lo, hi, fg) in enumerate(zip((0, -30, 0.002, -0.01), (2500, 50, 0.007, -0.0015), (0, 1, 1, 1))):
        el.fitMin[j], el.fitMax[j], el.fitFirstGuess[j] = lo, hi, fg
    s.proxy[0] = el
    s.selectedActive[0] = s.currentActive[0] = 1
    if use_urban:
        ur = A.default_proxy(A.proxyUrbanFraction, 1)
        ur.fittingFunction = A.linear
        ur.nFitParams = 2
        for j, (lo, hi) in enumerate(zip((-10, -50), (11, 50))):
            ur.fitMin[j], ur.fitMax[j] = lo, hi
            ur.fitFirstGuess[j] = 1
        s.proxy[1] = ur
        s.selectedActive[1] = s.currentActive[1] = 1
    s.selectedUseThermalInversion = s.currentUseThermalInversion = 1
    return s


def settings_classic_detrending(n_proxy=1):
    """Classic detrending (detrending(), interpolation.cpp:1380): elevation (+ optional urban) linear regressions."""
    s = A.default_settings()
    s.nProxy = n_proxy
    s.proxy[0] = A.default_proxy(A.proxyHeight, 0)
    s.selectedActive[0] = s.currentActive[0] = 1
    if n_proxy > 1:
        s.proxy[1] = A.default_proxy(A.proxyUrbanFraction, 1)
        s.selectedActive[1] = s.currentActive[1] = 1
    s.selectedUseThermalInversion = s.currentUseThermalInversion = 1
    return s


def set_points_range(settings, stations):
    """Crit3DInterpolationSettings::setPointsRange(min, max) as passDataToInterpolation does (spatialControl.cpp:564)."""
    settings.hasPointsRange = 1
    settings.pointsRangeMin = float(stations.value.min())
    settings.pointsRangeMax = float(stations.value.max())


def make_macro_areas(dem, stations, n_side=2, blend_cells=8, margin_m=None):
    """Synthetic macro areas for glocal detrending: the DEM split into n_side x n_side tiles whose weights ramp linearly
    across a band of `blend_cells` cells around every internal border (weights of a cell sum to 1, like the reference's
    glocalWeight_<i> maps, project.cpp:2592-2700); an area's cells are the flat (row * nrCols + col, weight) float pairs of
    every DEM-valid cell with weight > 0; its stations are those inside the tile grown by `margin_m` (default: the blend
    band), in station order. Returns _abi.MacroAreas."""
    rows, cols = dem.shape
    cs = dem.cell_size
    if margin_m is None:
        margin_m = blend_cells * cs

    def ramps(n):
        edges = np.linspace(0, n, n_side + 1)
        c = np.arange(n) + 0.5
        w = np.zeros((n_side, n))
        for k in range(n_side):
            lo, hi = edges[k], edges[k + 1]
            up = np.ones(n) if k == 0 else np.clip((c - (lo - blend_cells / 2)) / blend_cells, 0, 1)
            dn = np.ones(n) if k == n_side - 1 else np.clip(((hi + blend_cells / 2) - c) / blend_cells, 0, 1)
            w[k] = np.minimum(up, dn)
        return w / w.sum(axis=0), edges

    wr, er = ramps(rows)
    wc, ec = ramps(cols)
    valid = dem.data != np.float32(dem.flag)
    meteo, cells = [], []
    for i in range(n_side):
        for j in range(n_side):
            w = np.outer(wr[i], wc[j]).astype(np.float32)
            sel = valid & (w > 0)
            idx = np.flatnonzero(sel.ravel())
            pairs = np.empty((idx.size, 2), dtype=np.float32)
            pairs[:, 0] = idx.astype(np.float32)
            pairs[:, 1] = w.ravel()[idx]
            cells.append(pairs.ravel())
            y_hi = dem.lly + (rows - er[i]) * cs + margin_m
            y_lo = dem.lly + (rows - er[i + 1]) * cs - margin_m
            x_lo = dem.llx + ec[j] * cs - margin_m
            x_hi = dem.llx + ec[j + 1] * cs + margin_m
            inside = (stations.x >= x_lo) & (stations.x <= x_hi) & (stations.y >= y_lo) & (stations.y <= y_hi)
            meteo.append(np.flatnonzero(inside).astype(np.int32))
    return A.MacroAreas(meteo, cells)
