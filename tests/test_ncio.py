"""NetCDF classic container of the host mirror (mpassit_b200/host/ncio.cpp) against scipy.io.netcdf_file, an
independent implementation of the same format; and the CPU-side behaviour of the file driver (mpassit_run)."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from mpassit_b200 import defaults
from tests import mpas_files


@pytest.fixture(scope="module")
def host(engine_lib):
    from mpassit_b200 import build, host

    build.build_host()
    host.load()
    return host


def _sample(path, version):
    rng = np.random.default_rng(7)
    data = {}
    with netcdf_file(path, "w", version=version) as f:
        f.createDimension("Time", None)
        f.createDimension("nCells", 37)
        f.createDimension("nVertLevels", 5)
        f.createDimension("StrLen", 7)  # odd sizes: padding to 4 bytes matters
        f.title = b"sample"
        f.config_dt = np.float64(18.0)
        f.levels = np.int32(5)
        v = f.createVariable("fixed", "f8", ("nCells",))
        v.units = b"m"
        data["fixed"] = rng.normal(size=37)
        v[:] = data["fixed"]
        v = f.createVariable("idx", "i4", ("nCells", "nVertLevels"))
        data["idx"] = rng.integers(-5, 1 << 20, size=(37, 5)).astype(np.int32)
        v[:] = data["idx"]
        v = f.createVariable("theta", "f4", ("Time", "nCells", "nVertLevels"))
        v.long_name = b"potential temperature"
        data["theta"] = rng.normal(300, 5, size=(3, 37, 5)).astype(np.float32)
        v = f.createVariable("xtime", "S1", ("Time", "StrLen"))
        data["xtime"] = np.frombuffer(b"abcdefghijklmnopqrstu", "S1").reshape(3, 7)
        v2 = f.createVariable("s", "i2", ("Time",))
        data["s"] = np.array([3, -2, 9], np.int16)
        for r in range(3):
            f.variables["theta"][r] = data["theta"][r]
            f.variables["xtime"][r] = data["xtime"][r]
            v2[r] = data["s"][r]
    return data


@pytest.mark.parametrize("version", [1, 2])
def test_reader_against_scipy_files(host, tmp_path, version):
    p = str(tmp_path / f"s{version}.nc")
    data = _sample(p, version)
    d = host.nc_describe(p)
    assert d[0] == ["version", str(version), "numrecs", "3"]
    dims = {r[1]: int(r[2]) for r in d if r[0] == "dim"}
    assert dims == {"Time": 0, "nCells": 37, "nVertLevels": 5, "StrLen": 7}
    assert {r[1] for r in d if r[0] == "gatt"} == {"title", "config_dt", "levels"}
    assert sorted(r[1] for r in d if r[0] == "var") == sorted(["fixed", "idx", "theta", "xtime", "s"])  # scipy orders variables by size
    assert np.array_equal(host.nc_get(p, "fixed", 37), data["fixed"])
    assert np.array_equal(host.nc_get(p, "idx", 37 * 5), data["idx"].reshape(-1))
    for r in range(3):
        assert np.array_equal(host.nc_get(p, "theta", 37 * 5, rec=r), data["theta"][r].reshape(-1).astype(np.float64))
        assert host.nc_get(p, "s", 1, rec=r)[0] == data["s"][r]
    assert np.array_equal(host.nc_get(p, "theta", 4, rec=1, first=11), data["theta"][1].reshape(-1)[11:15].astype(np.float64))
    with pytest.raises(host.HostError, match="Variable not found"):
        host.nc_get(p, "nope", 1)
    with pytest.raises(host.HostError, match="exceeds"):
        host.nc_get(p, "fixed", 38)
    with pytest.raises(host.HostError, match="exceeds"):
        host.nc_get(p, "theta", 1, rec=3)


@pytest.mark.parametrize("src_version,dst_version", [(1, 1), (1, 2), (2, 2), (2, 1), (2, 5), (1, 0)])
def test_writer_round_trip(host, tmp_path, src_version, dst_version):
    """reader -> writer -> (scipy for CDF-1/2, the reader itself for CDF-5): same dimensions, attributes and values."""
    p, q = str(tmp_path / "a.nc"), str(tmp_path / "b.nc")
    data = _sample(p, src_version)
    host.nc_copy(p, q, dst_version)
    want_version = dst_version or 2
    dq = host.nc_describe(q)
    assert dq[0] == ["version", str(want_version), "numrecs", "3"]
    strip = lambda d: [[x for i, x in enumerate(r) if not (r[0] == "var" and i == 3)] for r in d[1:]]  # noqa: E731  (offsets differ)
    assert strip(dq) == strip(host.nc_describe(p))
    if want_version in (1, 2):
        with netcdf_file(q, "r", mmap=False) as f:
            assert f.version_byte == want_version
            assert f.title == b"sample" and f.config_dt == 18.0 and f.levels == 5
            assert f.variables["fixed"].units == b"m"
            assert f.variables["theta"].long_name == b"potential temperature"
            for k, a in data.items():
                assert np.array_equal(f.variables[k][:], a), k
    for r in range(3):
        assert np.array_equal(host.nc_get(q, "theta", 37 * 5, rec=r), data["theta"][r].reshape(-1).astype(np.float64))
        assert host.nc_get(q, "s", 1, rec=r)[0] == data["s"][r]
    assert np.array_equal(host.nc_get(q, "idx", 37 * 5), data["idx"].reshape(-1))


def test_reader_rejects_what_it_cannot_read(host, tmp_path):
    h5 = tmp_path / "h5.nc"
    h5.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(host.HostError, match="NetCDF-4 / HDF5"):
        host.nc_describe(str(h5))
    junk = tmp_path / "junk.nc"
    junk.write_bytes(b"not a netcdf file at all")
    with pytest.raises(host.HostError, match="magic"):
        host.nc_describe(str(junk))
    p = str(tmp_path / "ok.nc")
    _sample(p, 2)
    raw = open(p, "rb").read()
    cut = tmp_path / "cut.nc"
    cut.write_bytes(raw[: len(raw) - 40])  # data section shorter than the header promises
    with pytest.raises(host.HostError, match="past the end"):
        host.nc_describe(str(cut))
    cut.write_bytes(raw[:60])  # header itself truncated
    with pytest.raises(host.HostError):
        host.nc_describe(str(cut))
    with pytest.raises(host.HostError):
        host.nc_describe(str(tmp_path / "missing.nc"))


def test_mpas_file_helper_matches_the_reference_input_contract(host, tmp_path):
    """The synthetic init file carries what define_input_grid reads (model_grid.F90:286-419), in file order."""
    from mpassit_b200 import synth

    mesh = synth.regional_delaunay_mesh(n_points=300)
    p = str(tmp_path / "init.nc")
    ter = np.linspace(0, 1500, mesh.lonCell.size).astype(np.float32)
    mpas_files.write_grid_file(p, mesh, nz=6, nsoil=4, ter=ter)
    dims = {r[1]: int(r[2]) for r in host.nc_describe(p) if r[0] == "dim"}
    assert dims["nCells"] == mesh.lonCell.size and dims["nVertLevelsP1"] == 7 and dims["nSoilLevels"] == 4
    assert np.array_equal(host.nc_get(p, "lonCell", mesh.lonCell.size), mesh.lonCell)
    assert np.array_equal(host.nc_get(p, "verticesOnCell", mesh.verticesOnCell.size), mesh.verticesOnCell.reshape(-1))
    assert np.allclose(host.nc_get(p, "zs", 4), [0.05, 0.25, 0.7, 1.5])
    assert np.array_equal(host.nc_get(p, "ter", ter.size), ter.astype(np.float64))


def test_xytoll_and_map_factor(host, tmp_path):
    """xytoll (llxy_module.F90:166-216) agrees with the coordinate arrays; the Lambert map factor is 1 on the true
    latitude and follows Saucier's formula elsewhere (model_grid.F90:2229-2365)."""
    from mpassit_b200 import lib as L

    cfg = host.read_setup_namelist(defaults.write_namelist(str(tmp_path / "namelist.input"), nx=61, ny=41, dx=30000.0))
    lat, lon = host.target_coords(cfg, L.CENTER)
    la, lo = host.xytoll(cfg, 7.0, 5.0, 1)
    assert (la, lo) == (lat[4, 6], lon[4, 6])
    latu, lonu = host.target_coords(cfg, L.EDGE1)
    assert host.xytoll(cfg, 7.0, 5.0, 2) == (latu[4, 6], lonu[4, 6])
    mf = host.get_map_factor(cfg, np.array([38.5, 30.0, 50.0]))
    assert abs(mf[0] - 1.0) < 1e-12
    colat0, n = np.deg2rad(90 - 38.5), np.cos(np.deg2rad(90 - 38.5))
    for k, la in ((1, 30.0), (2, 50.0)):
        c = np.deg2rad(90 - la)
        assert abs(mf[k] - np.sin(colat0) / np.sin(c) * (np.tan(c / 2) / np.tan(colat0 / 2)) ** n) < 1e-12
    assert mf[1] > 1.0 and mf[2] > 1.0


def test_run_fails_loudly(host, tmp_path):
    """error_handler behaviour of the driver: bad namelist, nothing to do, and -- without a GPU -- no CPU fallback."""
    import torch

    with pytest.raises(host.HostError, match="NAMELIST"):
        host.run(str(tmp_path / "nope.nml"))
    nl = tmp_path / "namelist.input"
    nl.write_text("&config\n target_grid_type='lambert'\n nx=61\n ny=41\n dx=30000.\n dy=30000.\n ref_lat=38.5\n ref_lon=-97.5\n"
                  " truelat1=38.5\n truelat2=38.5\n stand_lon=-97.5\n interp_diag=.false.\n interp_hist=.false.\n/\n")
    with pytest.raises(host.HostError, match="INTERP_DIAG AND/OR INTERP_HIST"):
        host.run(str(nl))
    if not torch.cuda.is_available():
        p = defaults.write_namelist(str(tmp_path / "nl2"), nx=61, ny=41, dx=30000.0)
        with pytest.raises(host.HostError, match="CUDA"):
            host.run(p, str(tmp_path))


def _cpu_fields(wl):
    from mpassit_b200 import workload

    F = {g: [(nm, workload._field_values_torch(wl, g, nm, wl.levels_of(g, nm), k, "cpu").numpy())
             for k, (nm, _) in enumerate(wl.lists[g])] for g in ("diag", "hist_2d", "hist_3d", "soil")}
    ter = workload._field_values_torch(wl, "hist_2d", "ter", 1, 99, "cpu").numpy()
    return F, ter


def test_output_file_contract_header_only(host, tmp_path):
    """mpassit_run(device = -1) lays the output file out without a GPU: dimensions, global attributes, variable
    names / order / attributes and the grid description of write_to_file (write_data.F90:170-1210), quirks included."""
    from mpassit_b200 import lib as L
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    F, ter = _cpu_fields(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), F, ter)
    st = host.run(nl, str(tmp_path), device=-1)
    assert st.output_version == 2 and st.n_cells == wl.mesh.lonCell.size
    out, g, va, dims, order = mpas_files.read_output(paths["out"])
    it, jt, nz = wl.cfg.i_target, wl.cfg.j_target, wl.nz
    assert dims == {"Time": None, "west_east": it, "west_east_stag": it + 1, "south_north": jt, "south_north_stag": jt + 1,
                    "bottom_top": nz, "bottom_top_stag": nz + 1, "soil_layers_stag": wl.nsoil, "StrLen": 19}
    assert list(dims) == ["Time", "west_east", "west_east_stag", "south_north", "south_north_stag", "bottom_top",
                          "bottom_top_stag", "soil_layers_stag", "StrLen"]
    # global attributes, write_data.F90:199-300
    assert g["WEST-EAST_GRID_DIMENSION"] == it + 1 and g["BOTTOM-TOP_GRID_DIMENSION"] == nz + 1
    assert g["SIMULATION_START_DATE"] == mpas_files.START_TIME and g["START_DATE"] == mpas_files.START_TIME
    assert g["DX"] == 30000.0 and g["DY"] == 30000.0 and g["DT"] == 18.0
    assert (g["SF_SURFACE_PHYSICS"], g["MP_PHYSICS"], g["CU_PHYSICS"]) == (2, 18, 3)
    clat, clon = host.xytoll(wl.cfg, it / 2.0, jt / 2.0, 1)
    assert g["CEN_LAT"] == clat and g["CEN_LON"] == clon and g["MOAD_CEN_LAT"] == clat
    assert g["TRUELAT1"] == 38.5 and g["STAND_LON"] == -97.5 and g["POLE_LAT"] == 90.0 and g["POL_ELAT"] == 90.0
    assert g["MAP_PROJ"] == 1 and g["MAP_PROJ_CHAR"] == "Lambert Conformal" and g["PREC_ACC_DT"] == 3600
    assert g["WEST-EAST_PATCH_END_STAG"] == it + 1 and g["SOUTH-NORTH_PATCH_END_UNSTAG"] == jt
    # definition order, write_data.F90:303-893
    head = ["XLONG", "XLONG_U", "XLONG_V", "XLAT", "XLAT_U", "XLAT_V", "MAPFAC_M", "MAPFAC_U", "MAPFAC_V", "SINALPHA",
            "COSALPHA", "Z_C", "ZS", "HGT", "Times", "ITIMESTEP", "XTIME"]
    diag = [t for _, t in wl.lists["diag"]]  # 2-D and 3-D diag variables are defined as the list meets them (:571-613)
    want = (head + diag + ["SNOW", "SNOWH"] + ["PSFC", "TSK", "SST"] + ["XLAND"] + ["TSLB", "SMOIS", "SH2O"] +
            ["T", "QVAPOR", "QCLOUD", "QRAIN", "QICE", "QSNOW", "QGRAUP", "QNICE", "QNRAIN", "P_HYD", "P_TOP", "MUB", "MU"] +
            ["U", "V"] + ["PHB", "PH", "W"] + ["P", "PB"])
    assert order == want
    # shapes: NetCDF order is the Fortran dimension list reversed
    assert out["XLONG_U"].shape == (jt, it + 1) and out["XLAT_V"].shape == (jt + 1, it)
    assert out["U"].shape == (nz, jt, it + 1) and out["V"].shape == (nz, jt + 1, it) and out["PHB"].shape == (nz + 1, jt, it)
    assert out["Z_C"].shape == (nz + 1, jt, it) and out["TSLB"].shape == (wl.nsoil, jt, it) and out["P_TOP"].shape == ()
    # attributes and the writer's quirks
    assert va["XLONG"] == {"description": "LONGITUDE, WEST IS NEGATIVE", "units": "degree_east", "MemoryOrder": "XY ",
                           "coordinates": "XLONG XLAT", "stagger": "", "FieldType": 104}
    assert va["SINALPHA"]["description"] == "COSINE OF GRID ROTATION ANGLE ALPHA" and va["COSALPHA"] == {}
    assert va["MAPFAC_U"]["description"] == "LATITUDE, SOUTH IS NEGATIVE" and va["MAPFAC_U"]["stagger"] == "X"
    assert va["T"] == {"MemoryOrder": "XYZ ", "coordinates": "XLONG XLAT XTIME", "units": "unit_of_theta",
                       "description": "long name of theta", "stagger": "", "FieldType": 104}
    assert va["PHB"]["units"] == "gpm" and va["PHB"]["description"] == "Base Geopotential Height" and va["PHB"]["stagger"] == "Z"
    assert va["PH"]["description"] == "Perturbation Geopotential Height" and va["W"]["units"] == "unit_of_w"
    assert va["MU"]["description"] == "Perturbation long name of rho" and va["P_TOP"]["description"] == "PRESSURE TOP OF THE MODEL"
    assert va["U"]["coordinates"] == "XLONG_U XLAT_U XTIME" and va["U"]["units"] == "m s^{-1}" and va["V"]["stagger"] == "Y"
    assert va["PB"]["description"] == "BASE STATE PRESSURE (pfull)" and va["ITIMESTEP"]["FieldType"] == 106
    assert va["XTIME"]["units"] == "minutes since " + mpas_files.START_TIME
    # grid description
    for nm, s in (("", L.CENTER), ("_U", L.EDGE1), ("_V", L.EDGE2)):
        lat, lon = host.target_coords(wl.cfg, s)
        assert np.array_equal(out["XLAT" + nm], lat.astype(np.float32)) and np.array_equal(out["XLONG" + nm], lon.astype(np.float32))
        assert np.array_equal(out["MAPFAC" + (nm or "_M")], host.get_map_factor(wl.cfg, lat).astype(np.float32))
    assert np.array_equal(out["COSALPHA"], wl.cosa.astype(np.float32)) and np.array_equal(out["SINALPHA"], wl.sina.astype(np.float32))
    assert np.allclose(out["ZS"], [0.05, 0.25, 0.7, 1.5]) and out["Times"].tobytes().decode() == mpas_files.VALID_TIME
    assert out["XTIME"] == -540.0 and out["ITIMESTEP"] == int(-540 * 60 / 18.0)  # (start - valid), write_data.F90:1198
    assert not out["T"].any() and not out["P"].any()  # nothing regridded in this mode


def test_run_reports_input_file_errors(host, tmp_path):
    """netcdf_err-style failures (utils.F90:35-60) of read_input_data: missing variable, missing attribute, NetCDF-4."""
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    F, ter = _cpu_fields(wl)
    F2 = dict(F, hist_3d=[x for x in F["hist_3d"] if x[0] != "qv"])
    nl, paths = mpas_files.write_case(wl, str(tmp_path), F2, ter)
    with pytest.raises(host.HostError, match="reading field id - qv"):
        host.run(nl, str(tmp_path), device=-1)
    open(paths["history"], "wb").write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(host.HostError, match="opening: .*NetCDF-4"):
        host.run(nl, str(tmp_path), device=-1)
    (tmp_path / "histlist_soil").unlink()
    with pytest.raises(host.HostError, match="VARLIST FILE"):
        host.run(nl, str(tmp_path), device=-1)


def test_large_blocks_are_written_in_parallel_chunks(host, tmp_path):
    """Blocks of 8 MiB and more are cut across threads (disjoint pwrite regions): same bytes in the same place."""
    p, q = str(tmp_path / "big.nc"), str(tmp_path / "big_copy.nc")
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1 << 31, size=(2, 1500, 1001)).astype(np.int32)  # 12 MB per record, not a multiple of the chunk
    with netcdf_file(p, "w", version=2) as f:
        f.createDimension("Time", None)
        f.createDimension("y", 1500)
        f.createDimension("x", 1001)
        v = f.createVariable("a", "i4", ("Time", "y", "x"))
        w = f.createVariable("b", "i4", ("y", "x"))
        w[:] = a[1]
        v[0] = a[0]
        v[1] = a[1]
    host.nc_copy(p, q, 2)
    with netcdf_file(q, "r", mmap=False) as f:
        assert np.array_equal(f.variables["a"][:], a) and np.array_equal(f.variables["b"][:], a[1])


def test_output_contract_variants_header_only(host, tmp_path):
    """(a) lat-lon global target, hist only (BASELINE configs[0] shape): no rotation fields, MAP_PROJ 0, no PREC_ACC_DT;
    (b) wrf_mod_vars = .false.: winds are plain 3d_nz variables on the mass grid, no MU / P_TOP / PH / P / PB;
    (c) the same inputs in the 64-bit-data format (CDF-5) are read the same way."""
    from mpassit_b200 import workload

    # (a)
    da = tmp_path / "a"
    da.mkdir()
    wl = workload.make("c1", rundir=str(da))
    wl.nz = 3  # a thin column keeps the synthetic files small; the grid and lists are those of configs[0]
    F, ter = _cpu_fields(wl)
    nl, paths = mpas_files.write_case(wl, str(da), F, ter)
    host.run(nl, str(da), device=-1)
    out, g, va, dims, order = mpas_files.read_output(paths["out"])
    assert dims["west_east"] == 360 and dims["south_north"] == 180 and dims["bottom_top"] == 3
    assert g["MAP_PROJ"] == 0 and g["MAP_PROJ_CHAR"] == "Lat/Lon" and "PREC_ACC_DT" not in g
    assert "SINALPHA" not in order and "COSALPHA" not in order and "RAINC" not in order
    assert order.index("Z_C") == order.index("MAPFAC_V") + 1 and not out["MAPFAC_M"].any()
    assert order[-2:] == ["P", "PB"] and out["U"].shape == (3, 180, 361) and out["V"].shape == (3, 181, 360)
    # (b)
    db = tmp_path / "b"
    db.mkdir()
    wl = workload.make("mini", rundir=str(db))
    wl.cfg.wrf_mod_vars = 0
    wl.cfg.interp_diag = 0
    F, ter = _cpu_fields(wl)
    nl, paths = mpas_files.write_case(wl, str(db), F, ter)
    host.run(nl, str(db), device=-1)
    out, g, va, dims, order = mpas_files.read_output(paths["out"])
    for absent in ("MU", "P_TOP", "PH", "P", "PB", "RAINC", "REFL_10CM"):
        assert absent not in order
    it, jt = wl.cfg.i_target, wl.cfg.j_target
    assert out["U"].shape == (wl.nz, jt, it) and va["U"]["units"] == "unit_of_uReconstructZonal" and va["U"]["stagger"] == ""
    assert va["PHB"]["units"] == "unit_of_zgrid"  # the 'gpm' relabelling is a wrf_mod_vars feature (write_data.F90:814)
    i3 = [order.index(n) for n in ("T", "U", "V", "QVAPOR", "MUB", "PHB", "W")]
    assert i3 == sorted(i3)  # 3d_nz in list order (winds included), then 3d_nzp1
    # (c)
    for k in ("init", "diag", "history"):
        host.nc_copy(paths[k], paths[k] + ".cdf5", 5)
        import os
        os.replace(paths[k] + ".cdf5", paths[k])
    assert host.nc_describe(paths["history"])[0][:2] == ["version", "5"]
    raw = open(paths["out"], "rb").read()
    host.run(nl, str(db), device=-1)
    assert open(paths["out"], "rb").read() == raw


def test_esmf_weight_files(host, tmp_path):
    """The ESMF weight-file bridge (host/weights.cpp): a file laid out like ESMF_RegridWeightGen's (entries in no
    particular order, 1-based, extra variables present) reads back as a sorted 0-based CSR with ESMF's order kept
    inside each row; the engine-side writer produces a file scipy reads with the same matrix."""
    rng = np.random.default_rng(5)
    n_a, n_b = 500, 300
    lens = rng.integers(0, 5, n_b)
    rp = np.zeros(n_b + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    col = rng.integers(0, n_a, rp[-1]).astype(np.int32)
    w = rng.random(rp[-1])
    row = np.repeat(np.arange(n_b), lens)
    perm = rng.permutation(rp[-1])            # ESMF promises no order
    p = str(tmp_path / "esmf_w.nc")
    with netcdf_file(p, "w", version=2) as f:
        for n, l in (("n_a", n_a), ("n_b", n_b), ("n_s", int(rp[-1])), ("nv_a", 3), ("src_grid_rank", 1)):
            f.createDimension(n, l)
        f.title = b"ESMF Offline Regridding Weight Generator"
        f.createVariable("frac_b", "f8", ("n_b",))[:] = 1.0
        f.createVariable("col", "i4", ("n_s",))[:] = col[perm] + 1
        f.createVariable("row", "i4", ("n_s",))[:] = row[perm] + 1
        f.createVariable("S", "f8", ("n_s",))[:] = w[perm]
    na, nb, rp2, col2, w2 = host.read_esmf_weights(p)
    assert (na, nb) == (n_a, n_b) and np.array_equal(rp2, rp)
    for i in range(n_b):                      # same entries per row, in the file's order
        k = np.flatnonzero(row[perm] == i)
        assert np.array_equal(col2[rp[i]:rp[i + 1]], col[perm][k]) and np.array_equal(w2[rp[i]:rp[i + 1]], w[perm][k])
    q = str(tmp_path / "ours_w.nc")
    host.write_esmf_weights(q, n_a, rp, col, w, "Bilinear")
    with netcdf_file(q, "r", mmap=False) as f:
        assert f.dimensions["n_a"] == n_a and f.dimensions["n_b"] == n_b and f.dimensions["n_s"] == rp[-1]
        assert np.array_equal(f.variables["col"][:], col + 1) and np.array_equal(f.variables["row"][:], row + 1)
        assert np.array_equal(f.variables["S"][:], w)
    na, nb, rp3, col3, w3 = host.read_esmf_weights(q)
    assert np.array_equal(rp3, rp) and np.array_equal(col3, col) and np.array_equal(w3, w)
    # an empty matrix survives the trip as a single zero weight; junk is refused
    host.write_esmf_weights(q, 7, np.zeros(4, np.int32), np.zeros(0, np.int32), np.zeros(0))
    na, nb, rp4, col4, w4 = host.read_esmf_weights(q)
    assert (na, nb) == (7, 3) and rp4.tolist() == [0, 1, 1, 1] and w4.tolist() == [0.0]
    with netcdf_file(p, "w", version=2) as f:
        f.createDimension("n_s", 2)
        f.createVariable("S", "f8", ("n_s",))[:] = 1.0
    with pytest.raises(host.HostError, match="not an ESMF weight file"):
        host.read_esmf_weights(p)


def test_esmf_kit_mesh_grid_files_and_compare(host, tmp_path):
    """tools/esmf_kit.py: the ESMF unstructured-mesh file follows what the reference hands to ESMF_MeshCreate
    (model_grid.F90:446-497), the SCRIP file carries the CENTER points with their 4 CORNER-stagger corners, and
    `compare` tells identical / perturbed / restructured matrices apart."""
    from mpassit_b200 import workload
    from tools import esmf_kit

    wl = workload.make("mini", rundir=str(tmp_path))
    paths = esmf_kit.export("mini", str(tmp_path))
    m = wl.mesh
    with netcdf_file(paths["mesh"], "r", mmap=False) as f:
        assert f.gridType == b"unstructured" and f.dimensions["elementCount"] == m.lonCell.size
        nc, conn, num = f.variables["nodeCoords"][:], f.variables["elementConn"][:], f.variables["numElementConn"][:]
        cc = f.variables["centerCoords"][:]
        assert f.variables["elementConn"].start_index == 1 and f.variables["nodeCoords"].units == b"degrees"
    lonV, latV = esmf_kit.mesh_degrees(m.lonVertex, m.latVertex)
    assert np.array_equal(nc[:, 0], lonV) and np.array_equal(nc[:, 1], latV) and nc[:, 0].max() <= 180.0
    assert np.array_equal(num, (m.verticesOnCell > 0).sum(axis=1))
    for c in (0, 17, m.lonCell.size - 1):
        v = m.verticesOnCell[c][m.verticesOnCell[c] > 0]
        assert np.array_equal(conn[c, :v.size], v) and (conn[c, v.size:] == -1).all()
    assert np.array_equal(cc[:, 1], esmf_kit.mesh_degrees(m.lonCell, m.latCell)[1])
    latM, lonM = wl.grids["M"]
    latQ, lonQ = wl.grids["CORNER"]
    with netcdf_file(paths["grid"], "r", mmap=False) as f:
        assert f.variables["grid_dims"][:].tolist() == [latM.shape[1], latM.shape[0]]
        assert np.array_equal(f.variables["grid_center_lat"][:], latM.reshape(-1))
        cl = f.variables["grid_corner_lat"][:].reshape(latM.shape + (4,))
        co = f.variables["grid_corner_lon"][:].reshape(latM.shape + (4,))
    j, i = 5, 9
    assert cl[j, i].tolist() == [latQ[j, i], latQ[j, i + 1], latQ[j + 1, i + 1], latQ[j + 1, i]]
    assert co[j, i].tolist() == [lonQ[j, i], lonQ[j, i + 1], lonQ[j + 1, i + 1], lonQ[j + 1, i]]
    # counter-clockwise: positive signed area in the (lon, lat) plane
    x, y = co[j, i], cl[j, i]
    assert sum(x[k] * y[(k + 1) % 4] - x[(k + 1) % 4] * y[k] for k in range(4)) > 0
    # compare
    rng = np.random.default_rng(2)
    rp = np.arange(0, 3 * 200 + 1, 3).astype(np.int32)
    col = rng.integers(0, 900, 600).astype(np.int32)
    w = rng.random(600)
    a, b = str(tmp_path / "wa.nc"), str(tmp_path / "wb.nc")
    host.write_esmf_weights(a, 900, rp, col, w)
    shuffled = col.reshape(200, 3)[:, ::-1].reshape(-1)        # same sources, another order inside the rows
    host.write_esmf_weights(b, 900, rp, shuffled, w.reshape(200, 3)[:, ::-1].reshape(-1))
    r = esmf_kit.compare(a, b)
    assert r["rows_same_sources"] == 200 and r["max_weight_difference"] == 0.0 and r["verdict"].startswith("identical structure, weights")
    w2 = w.copy()
    w2[301] += 1e-6
    host.write_esmf_weights(b, 900, rp, col, w2)
    r = esmf_kit.compare(a, b)
    assert r["verdict"] == "identical structure" and abs(r["max_weight_difference"] - 1e-6) < 1e-12 and r["worst_row"] == 100
    col2 = col.copy()
    col2[0] = (col2[0] + 1) % 900
    host.write_esmf_weights(b, 900, rp, col2, w)
    assert esmf_kit.compare(a, b)["verdict"] == "structures differ"


def _corners_numpy(lat, lon, dx):
    """get_cell_corners (model_grid.F90:1902-1972) restated with numpy: bearings 135 / 225 / 45 / 315, truncated pi."""
    pi, R = 3.14159265359, 6370000.0
    d = np.sqrt(dx ** 2.0 / 2.0)
    nj, ni = lat.shape

    def step(la, lo, b):
        lat1, lon1, br = la * (pi / 180.0), lo * (pi / 180.0), b * pi / 180.0
        lat2 = np.arcsin(np.sin(lat1) * np.cos(d / R) + np.cos(lat1) * np.sin(d / R) * np.cos(br))
        lon2 = lon1 + np.arctan2(np.sin(br) * np.sin(d / R) * np.cos(lat1), np.cos(d / R) - np.sin(lat1) * np.sin(lat2))
        return lat2 * 180.0 / pi, lon2 * 180.0 / pi

    clat, clon = np.empty((nj + 1, ni + 1)), np.empty((nj + 1, ni + 1))
    clat[:nj, :ni], clon[:nj, :ni] = step(lat, lon, 135.0)
    clat[:nj, ni], clon[:nj, ni] = step(lat[:, -1], lon[:, -1], 225.0)
    clat[nj, :ni], clon[nj, :ni] = step(lat[-1, :], lon[-1, :], 45.0)
    clat[nj, ni], clon[nj, ni] = step(lat[-1, -1], lon[-1, -1], 315.0)
    return clat, clon


def test_target_grid_from_a_wrf_style_file(host, tmp_path):
    """target_grid_type = 'file' (define_target_grid_file, model_grid.F90:1203-1890): the output of a parameter-mode
    run is itself a valid target file; a run that reads it reproduces its grid description (stored as float, so
    bit-identical), takes sizes / projection from its attributes, keeps its terrain when no hist data is regridded, and
    synthesises the corners the way get_cell_corners does."""
    from mpassit_b200 import lib as L
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    F, ter = _cpu_fields(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), F, ter)
    host.run(nl, str(tmp_path), device=-1)
    out1, g1, va1, dims1, order1 = mpas_files.read_output(paths["out"])
    target = str(tmp_path / "wrf_target.nc")
    import os
    os.replace(paths["out"], target)
    wl.cfg.interp_hist = 0
    nl2, paths2 = mpas_files.write_case(wl, str(tmp_path), F, ter, target_file=target)
    host.run(nl2, str(tmp_path), device=-1)
    out2, g2, va2, dims2, order2 = mpas_files.read_output(paths2["out"])
    for k in ("west_east", "south_north", "west_east_stag", "south_north_stag"):
        assert dims2[k] == dims1[k]
    for k in ("XLAT", "XLONG", "XLAT_U", "XLONG_U", "XLAT_V", "XLONG_V", "MAPFAC_M", "MAPFAC_U", "MAPFAC_V", "SINALPHA", "COSALPHA"):
        assert np.array_equal(out2[k], out1[k]), k
    for k in ("DX", "DY", "CEN_LAT", "CEN_LON", "MOAD_CEN_LAT", "TRUELAT1", "TRUELAT2", "STAND_LON", "POLE_LAT", "MAP_PROJ",
              "MAP_PROJ_CHAR", "WEST-EAST_GRID_DIMENSION"):
        assert g2[k] == g1[k], k
    assert "PSFC" not in order2 and "RAINC" in order2          # diag only this time
    assert np.array_equal(out2["HGT"], out1["HGT"])            # hgt_target_grid as read from the target file
    # corners: the C mirror against the numpy restatement, on the float-rounded centres the file holds
    lat = out1["XLAT"].astype(np.float64)
    lon = out1["XLONG"].astype(np.float64)
    clat, clon = host.get_cell_corners(lat, lon, 30000.0)
    wlat, wlon = _corners_numpy(lat, lon, 30000.0)
    assert np.abs(clat - wlat).max() < 1e-12 and np.abs(clon - wlon).max() < 1e-12
    # what the reference's bearings mean on the ground: interior corners sit south-EAST of their centre (bearing 135),
    # half a cell diagonal away
    assert (clat[:-1, :-1] < lat).all() and (clon[:-1, :-1] > lon).all()
    d = 6370000.0 * np.hypot(np.deg2rad(clat[:-1, :-1] - lat), np.deg2rad(clon[:-1, :-1] - lon) * np.cos(np.deg2rad(lat)))
    assert np.abs(d / np.sqrt(30000.0 ** 2 / 2) - 1).max() < 1e-3
    # ... so the quad ESMF builds from corners (i,j), (i+1,j), (i+1,j+1), (i,j+1) is centred on mass point (i+1, j):
    # in this mode the reference's conservative cells are shifted one column east
    qlat = 0.25 * (clat[:-2, :-2] + clat[:-2, 1:-1] + clat[1:-1, 1:-1] + clat[1:-1, :-2])
    qlon = 0.25 * (clon[:-2, :-2] + clon[:-2, 1:-1] + clon[1:-1, 1:-1] + clon[1:-1, :-2])
    def km(la1, lo1, la2, lo2):
        return 6370.0 * np.hypot(np.deg2rad(la1 - la2), np.deg2rad(lo1 - lo2) * np.cos(np.deg2rad(la2)))
    assert km(qlat, qlon, lat[:-1, 1:], lon[:-1, 1:]).max() < 0.15 * 30.0     # (true bearings on a rotated Lambert grid)
    assert km(qlat, qlon, lat[:-1, :-1], lon[:-1, :-1]).min() > 0.8 * 30.0     # a whole cell away from its own centre
    # a target file without the staggered coordinates is refused like netcdf_err does
    from scipy.io import netcdf_file as ncf
    bad = str(tmp_path / "bad_target.nc")
    with ncf(bad, "w", version=2) as f:
        f.createDimension("west_east", 4)
        f.createDimension("south_north", 3)
        f.DX = 30000.0
    nl3, _ = mpas_files.write_case(wl, str(tmp_path), F, ter, target_file=bad)
    with pytest.raises(host.HostError, match="reading CEN_LAT"):
        host.run(nl3, str(tmp_path), device=-1)


@pytest.mark.parametrize("start,valid,dt,minutes", [
    ("2024-02-28_18:00:00", "2024-03-01_06:00:00", 20.0, -36 * 60.0),                 # across a leap day
    ("2023-12-31_23:30:00", "2024-01-01_00:15:30", 0.0, -45.5),                        # across a year, dt unknown
    ("2024-03-25_09:00:00", "2024-03-25_09:00:00", 18.0, 0.0),
    ("2100-02-28_00:00:00", "2100-03-01_00:00:00", 60.0, -1440.0),                     # 2100 is not a leap year
])
def test_xtime_and_itimestep(host, tmp_path, start, valid, dt, minutes):
    """XTIME = (start - valid) in minutes and ITIMESTEP = int(seconds / config_dt), 0 when config_dt is not positive
    (write_data.F90:1184-1210, datetime arithmetic of datetime_module)."""
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    wl.cfg.interp_hist = 0
    F, ter = _cpu_fields(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), {"diag": F["diag"]}, ter, start=start, valid=valid, config_dt=dt)
    host.run(nl, str(tmp_path), device=-1)
    out, g, va, dims, order = mpas_files.read_output(paths["out"])
    assert out["XTIME"] == np.float32(minutes) and out["Times"].tobytes().decode() == valid
    assert out["ITIMESTEP"] == (int(minutes * 60.0 / dt) if dt > 0 else 0)
    assert g["START_DATE"] == start and va["XTIME"]["description"] == "minutes since " + start


def test_command_line_entry(host, tmp_path, monkeypatch, capsys):
    """`python -m mpassit_b200 <namelist>`: missing namelist and error_handler-style failures give non-zero exits."""
    from mpassit_b200 import __main__ as cli

    monkeypatch.chdir(tmp_path)
    assert cli.main([]) == 1                      # no fort.41 here (mpassit.F90:60-65)
    assert "fort.41" in capsys.readouterr().err
    (tmp_path / "fort.41").write_text("&config\n target_grid_type='lambert'\n nx=61\n ny=41\n dx=30000.\n dy=30000.\n"
                                      " ref_lat=38.5\n ref_lon=-97.5\n truelat1=38.5\n truelat2=38.5\n stand_lon=-97.5\n"
                                      " interp_diag=.false.\n interp_hist=.false.\n/\n")
    assert cli.main([]) == 999
    assert "INTERP_DIAG AND/OR INTERP_HIST" in capsys.readouterr().err


@pytest.mark.parametrize("version", [1, 2])
def test_lone_record_variable_is_not_padded(host, tmp_path, version):
    """The classic format pads every variable to 4 bytes EXCEPT a single record variable of a 1- or 2-byte type,
    whose records are packed (format specification, 'vsize' note): read scipy's packing, write the same."""
    p, q = str(tmp_path / "one.nc"), str(tmp_path / "one_copy.nc")
    a = np.arange(15, dtype=np.int16).reshape(5, 3) - 7            # 6 bytes per record
    with netcdf_file(p, "w", version=version) as f:
        f.createDimension("t", None)
        f.createDimension("x", 3)
        v = f.createVariable("s", "i2", ("t", "x"))
        for r in range(5):
            v[r] = a[r]
    for r in range(5):
        assert np.array_equal(host.nc_get(p, "s", 3, rec=r), a[r].astype(np.float64))
    host.nc_copy(p, q, version)
    with netcdf_file(q, "r", mmap=False) as f:
        assert np.array_equal(f.variables["s"][:], a)
    # with a second record variable the padding is back
    with netcdf_file(p, "w", version=version) as f:
        f.createDimension("t", None)
        f.createDimension("x", 3)
        v = f.createVariable("s", "i2", ("t", "x"))
        u = f.createVariable("b", "i1", ("t",))
        for r in range(5):
            v[r] = a[r]
            u[r] = r - 2
    host.nc_copy(p, q, version)
    for r in range(5):
        assert np.array_equal(host.nc_get(p, "s", 3, rec=r), a[r].astype(np.float64)) and host.nc_get(p, "b", 1, rec=r)[0] == r - 2
    with netcdf_file(q, "r", mmap=False) as f:
        assert np.array_equal(f.variables["s"][:], a) and f.variables["b"][:].tolist() == [-2, -1, 0, 1, 2]


@pytest.mark.parametrize("seed", range(12))
def test_random_files_round_trip(host, tmp_path, seed):
    """Seeded random classic files (dimension counts, record / fixed variables of every classic type, odd sizes,
    text / numeric attributes): scipy writes, the reader + writer copy, scipy reads the copy back unchanged."""
    rng = np.random.default_rng(100 + seed)
    p, q = str(tmp_path / "r.nc"), str(tmp_path / "r_copy.nc")
    vs = int(rng.integers(1, 3))
    types = ["i1", "i2", "i4", "f4", "f8", "S1"]
    data, atts = {}, {}
    nrec = int(rng.integers(0, 4))
    with netcdf_file(p, "w", version=vs) as f:
        f.createDimension("rec", None)
        dims = {}
        for k in range(int(rng.integers(1, 4))):
            dims[f"d{k}"] = int(rng.integers(1, 8))
            f.createDimension(f"d{k}", dims[f"d{k}"])
        f.note = ("x" * int(rng.integers(1, 9))).encode()
        f.scale = np.float64(rng.normal())
        nvar = int(rng.integers(1, 6))
        for k in range(nvar):
            t = types[int(rng.integers(0, len(types)))]
            names = list(rng.choice(list(dims), size=int(rng.integers(0, len(dims) + 1)), replace=False))
            isrec = bool(rng.integers(0, 2))
            if not isrec and not names:
                names = [list(dims)[0]]  # (scipy cannot assign 0-d fixed variables)
            shape = tuple(dims[n] for n in names)
            v = f.createVariable(f"v{k}", t, (("rec",) if isrec else ()) + tuple(names))
            if rng.integers(0, 2):
                v.units = b"K"
                v.fill = np.int32(rng.integers(-5, 5))
            def draw(sh):
                if t == "S1":
                    return rng.integers(97, 123, size=sh).astype(np.uint8).view("S1").reshape(sh)
                if t[0] == "i":
                    return rng.integers(-100, 100, size=sh).astype(t)
                return rng.normal(size=sh).astype(t)
            if isrec:
                a = draw((nrec,) + shape)
                for r in range(nrec):
                    v[r] = a[r]
                if nrec == 0:
                    a = a.reshape((0,) + shape)
            else:
                a = draw(shape)
                v[:] = a
            data[f"v{k}"] = a
    out_version = int(rng.choice([1, 2]))
    host.nc_copy(p, q, out_version)
    with netcdf_file(q, "r", mmap=False) as f:
        assert f.version_byte == out_version and f.note == getattr(netcdf_file(p, "r", mmap=False), "note")
        for k, a in data.items():
            got = f.variables[k][:] if f.variables[k].shape else f.variables[k].getValue()
            assert np.array_equal(np.asarray(got).reshape(a.shape), a), k
    host.nc_copy(p, q, 5)                                   # and through the 64-bit-data format with the reader alone
    desc = host.nc_describe(q)
    any_rec = any(r[0] == "var" and r[4:5] == ["rec"] for r in desc)
    assert desc[0] == ["version", "5", "numrecs", str(nrec if any_rec else 0)]
    for k, a in data.items():
        if a.dtype.kind == "S" or a.size == 0:
            continue
        isrec = [r for r in desc if r[0] == "var" and r[1] == k][0][4:5] == ["rec"]
        flat = a.reshape(nrec, -1) if isrec else a.reshape(1, -1)
        for r in range(flat.shape[0]):
            assert np.array_equal(host.nc_get(q, k, flat.shape[1], rec=r), flat[r].astype(np.float64)), k
