"""ESRI float grids (.hdr + .flt), the on-disk format at either end of the path (gis/gisIO.cpp): files written by the
reference are read straight into HBM, files written from HBM are byte-identical to the reference's (header text included),
16-bit and 8-bit data files and the header rules (case-insensitive keys, NODATA / NODATA_value, missing keys) follow
gis::readEsriGridHeader / readRasterFloatData."""
import os

import numpy as np
import pytest

import praga_b200 as pb
from praga_b200 import _abi as A
from praga_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_read_what_the_reference_writes_and_write_the_same_bytes(engine, ref, tmp_path):
    dem = syn.make_dem(300, 333, cell_size=450.5)
    dem = A.Raster(dem.data, 512345.678, 4871234.5, 450.5, dem.flag)
    base_r, base_g = str(tmp_path / "ref_dem"), str(tmp_path / "gpu_dem")
    assert ref.write_esri(dem, base_r)
    for name in (base_r, base_r + ".flt"):                       # with and without the extension (readEsriGrid :525-545)
        rid, h, mn, mx = engine.read_esri(name)
        try:
            hr, gr, mnr, mxr = ref.read_esri(name)
            assert (h.nrRows, h.nrCols, h.cellSize, h.invCellSize, h.llx, h.lly, h.flag) == \
                   (hr.nrRows, hr.nrCols, hr.cellSize, hr.invCellSize, hr.llx, hr.lly, hr.flag)
            assert np.array_equal(engine.download(rid, (h.nrRows, h.nrCols)), gr)
            assert mn == mnr and mx == mxr
            engine.write_esri(rid, base_g)
        finally:
            engine.free(rid)
        assert open(base_g + ".hdr", "rb").read() == open(base_r + ".hdr", "rb").read()
        assert open(base_g + ".flt", "rb").read() == open(base_r + ".flt", "rb").read()
    # and the reference reads what the engine wrote
    hr, gr, _, _ = ref.read_esri(base_g)
    assert np.array_equal(gr, dem.data)


def test_interpolation_output_round_trips_through_the_file_format(engine, ref, tmp_path):
    """file -> HBM -> interpolationRaster (resident) -> result kept on the device -> file: no host raster in between."""
    dem = syn.make_dem(260, 280)
    urban = syn.make_urban(260, 280)
    ref.write_esri(dem, str(tmp_path / "dem"))
    ref.write_esri(urban, str(tmp_path / "urban"))
    did, hd, _, _ = engine.read_esri(str(tmp_path / "dem"))
    uid, _, _, _ = engine.read_esri(str(tmp_path / "urban.flt"))
    try:
        st = syn.make_stations(dem, 90, proxies=(dem, urban))
        s = syn.settings_multiple_detrending()
        syn.set_points_range(s, st)
        okp, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
        st = st.subset(keep)
        out = np.empty(dem.shape, dtype=np.float32)
        ok, g, mn, mx, nv = engine.interpolation_raster_resident(st, s, A.airTemperature, did, (did, uid), out=out)
        okc, gc, _, _, nvc, _ = ref.interpolation_raster(st, s, A.airTemperature, dem, (dem, urban))
        assert ok == okc and nv == nvc
        m = gc != np.float32(dem.flag)
        assert np.abs(g[m].astype(np.float64) - gc[m]).max() <= 1e-4
        res = A.Raster(g, dem.llx, dem.lly, dem.cell_size, dem.flag)
        rid = engine.upload(res)
        engine.write_esri(rid, str(tmp_path / "tavg"))
        engine.free(rid)
        hr, gr, mnr, mxr = ref.read_esri(str(tmp_path / "tavg"))
        assert np.array_equal(gr, g) and abs(mnr - mn) <= 1e-6 and abs(mxr - mx) <= 1e-6
    finally:
        engine.free(did)
        engine.free(uid)


def test_header_rules_and_narrow_data_types(engine, tmp_path):
    rng = np.random.default_rng(3)
    vals = rng.integers(-300, 3000, size=(40, 50)).astype(np.int16)
    base = str(tmp_path / "short")
    vals.tofile(base + ".flt")
    open(base + ".hdr", "w").write("NCOLS 50\nnRows 40\nXLLCORNER 10.5\nyllcorner 20.25\nCellSize 30\nnodata -300\ndatatype 2\n")
    rid, h, mn, mx = engine.read_esri(base)
    try:
        got = engine.download(rid, (40, 50))
    finally:
        engine.free(rid)
    assert (h.nrRows, h.nrCols, h.cellSize, h.llx, h.lly, h.flag) == (40, 50, 30.0, 10.5, 20.25, -300.0)
    assert np.array_equal(got, vals.astype(np.float32))
    keep = vals[vals != -300]
    assert mn == float(keep.min()) and mx == float(keep.max())
    b8 = rng.integers(0, 255, size=(12, 9)).astype(np.uint8)
    base8 = str(tmp_path / "byte")
    b8.tofile(base8 + ".flt")
    open(base8 + ".hdr", "w").write("ncols 9\nnrows 12\nxllcorner 0\nyllcorner 0\ncellsize 1\nNODATA_value 255\nDATATYPE 1\n")
    rid, h, _, _ = engine.read_esri(base8 + ".flt")
    try:
        assert np.array_equal(engine.download(rid, (12, 9)), b8.astype(np.float32))
    finally:
        engine.free(rid)
    # fewer than six of the required keys: "Missing keys in header file."
    open(str(tmp_path / "bad.hdr"), "w").write("ncols 9\nnrows 12\ncellsize 1\n")
    with pytest.raises(pb.PragaB200Error, match="Missing keys"):
        engine.read_esri(str(tmp_path / "bad"))
    with pytest.raises(pb.PragaB200Error, match="Missing file"):
        engine.read_esri(str(tmp_path / "nothing_here"))
    # a data file shorter than the header says: "Error reading raster data."
    open(str(tmp_path / "cut.hdr"), "w").write("ncols 9\nnrows 12\nxllcorner 0\nyllcorner 0\ncellsize 1\nNODATA_value -9999\n")
    np.zeros(50, dtype=np.float32).tofile(str(tmp_path / "cut.flt"))
    with pytest.raises(pb.PragaB200Error, match="Error reading raster data"):
        engine.read_esri(str(tmp_path / "cut"))
