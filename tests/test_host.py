"""Host-side mirror (C++) of the reference's Fortran stages: namelist, var-lists, regrid
classes, decomposition helpers, target-grid coordinates.  No GPU needed."""
import os

import numpy as np
import pytest

from mpassit_b200 import defaults


@pytest.fixture(scope="module")
def host(engine_lib):
    from mpassit_b200 import build, host

    build.build_host()
    host.load()
    return host


def test_namelist_conus_defaults_and_derived(host, tmp_path):
    p = defaults.write_namelist(str(tmp_path / "namelist.input"))
    c = host.read_setup_namelist(p)
    assert (c.nx, c.ny, c.i_target, c.j_target) == (1801, 1061, 1800, 1060)   # mass grid = (nx-1) x (ny-1)
    assert c.proj_code == host.PROJ_LC and c.map_proj_char == b"Lambert Conformal"
    assert c.interp_diag and c.interp_hist and c.wrf_mod_vars and not c.esmf_log
    assert c.is_regional == 1 and c.interp_as_bundle == 1                     # defaults, program_setup.F90:41,70
    assert c.known_x == 1801 / 2.0 and c.known_y == 1061 / 2.0                  # (i_target+1)/2
    assert c.truelat2 == 38.5 and c.dxkm == 3000.0 and c.block_decomp_file == b"NULL"
    assert c.pole_lat == 90.0 and c.pole_lon == 0.0


def test_namelist_syntax_variants_and_errors(host, tmp_path):
    f = tmp_path / "n1"
    f.write_text("! comment\n &CONFIG target_grid_type='LAT-LON', nx=361, ny = 181 , is_regional=.false.,\n"
                 " stand_lon = -180.d0 ! trailing\n interp_hist = T\n /\n")
    c = host.read_setup_namelist(str(f))
    assert c.proj_code == host.PROJ_LATLON and (c.i_target, c.j_target) == (360, 180)
    assert c.dlondeg == 1.0 and c.dlatdeg == 1.0 and c.known_lon == -179.5 and c.known_lat == -89.5
    assert c.known_x == 1.0 and c.known_y == 1.0 and c.interp_hist == 1
    # reference's shipped sample uses '-' for '=' and '. true.' : list-directed namelist read fails
    g = tmp_path / "n2"
    g.write_text('&config\n grid_file_input_grid-"x"\n wrf_mod_vars=. true.\n/\n')
    with pytest.raises(host.HostError) as e:
        host.read_setup_namelist(str(g))
    assert "READING SETUP NAMELIST" in str(e.value)
    with pytest.raises(host.HostError) as e:
        host.read_setup_namelist(str(tmp_path / "missing"))
    assert "OPENING SETUP NAMELIST" in str(e.value)
    h = tmp_path / "n3"
    h.write_text("&config target_grid_type='utm' nx=10 ny=10 /")
    with pytest.raises(host.HostError) as e:
        host.read_setup_namelist(str(h))
    assert "invalid target_grid_type" in str(e.value)
    k = tmp_path / "n4"
    k.write_text("&config target_grid_type='lat-lon' nx=10 ny=10 /")        # global lat-lon but is_regional default
    with pytest.raises(host.HostError):
        host.read_setup_namelist(str(k))
    m = tmp_path / "n5"
    m.write_text("&config target_grid_type='lambert' nx=10 ny=10 dx=3000. ref_lat=30. ref_lon=-90. stand_lon=-90. ref_x=3. /")
    with pytest.raises(host.HostError) as e:
        host.read_setup_namelist(str(m))
    assert "TRUELAT1" in str(e.value) or "ref_x, ref_y" in str(e.value)


def test_varlists_and_classes(host, tmp_path):
    paths = defaults.write_varlists(str(tmp_path))
    d = host.read_varlist(paths["diaglist"])
    assert d == defaults.DIAGLIST and len(d) == 19
    (tmp_path / "ragged").write_text("\n  a   A\n\n\tb\tB  \n'c c' \"C\"\n")
    assert host.read_varlist(str(tmp_path / "ragged")) == [("a", "A"), ("b", "B"), ("c c", "C")]
    (tmp_path / "bad").write_text("onlyone\n")
    with pytest.raises(host.HostError):
        host.read_varlist(str(tmp_path / "bad"))
    with pytest.raises(host.HostError) as e:
        host.read_varlist(str(tmp_path / "nope"))
    assert "not exist" in str(e.value)
    k2 = {n: host.classify_hist_2d(n) for n, _ in defaults.HISTLIST_2D}
    assert k2 == {"surface_pressure": host.CLASS_2D_PATCH, "xland": host.CLASS_2D_NSTD, "skintemp": host.CLASS_2D_PATCH,
                  "snow": host.CLASS_2D_CONS, "snowh": host.CLASS_2D_CONS, "sst": host.CLASS_2D_PATCH}
    for n in ("ivgtyp", "isltyp", "landmask"):
        assert host.classify_hist_2d(n) == host.CLASS_2D_NSTD
    k3 = [host.classify_hist_3d(n, True) for n, _ in defaults.HISTLIST_3D]
    assert k3.count(host.CLASS_3D_NZ) == 11 and k3.count(host.CLASS_3D_NZP1) == 2
    assert host.classify_hist_3d("uReconstructZonal", True) == host.CLASS_U
    assert host.classify_hist_3d("uReconstructZonal", False) == host.CLASS_3D_NZ   # only peeled off with wrf_mod_vars
    assert host.classify_hist_3d("vorticity", True) == host.CLASS_3D_VERT
    assert host.classify_diag("refl10cm") == host.CLASS_DIAG_3D and host.classify_diag("t2m") == host.CLASS_DIAG_2D


def test_para_range_and_block_decomp(host, tmp_path):
    # model_grid.F90:2428-2441: balanced, contiguous, 1-based inclusive
    for n, p in ((10, 3), (1060, 8), (7, 8), (1061, 8)):
        spans = [host.para_range(1, n, p, r) for r in range(p)]
        assert spans[0][0] == 1 and spans[-1][1] == n
        assert all(spans[r + 1][0] == spans[r][1] + 1 for r in range(p - 1))
        lens = [b - a + 1 for a, b in spans]
        assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)
    f = tmp_path / "part.3"
    f.write_text("0\n1\n\n2\n1\n0\n")
    assert host.read_block_decomp_file(str(f), 5, 3).tolist() == [0, 1, 2, 1, 0]
    with pytest.raises(host.HostError) as e:
        host.read_block_decomp_file(str(f), 6, 3)
    assert "CONTAINS MORE CELLS" in str(e.value)
    with pytest.raises(host.HostError) as e:
        host.read_block_decomp_file(str(f), 5, 4)
    assert "PROCESSES BUT" in str(e.value)


@pytest.mark.parametrize("stag,code", [("M", 0), ("U", 1), ("V", 2), ("CORNER", 3)])
def test_lambert_target_coords_match_numpy_restatement(host, tmp_path, stag, code):
    from oracle import proj_oracle as po

    c = host.read_setup_namelist(defaults.write_namelist(str(tmp_path / "nl"), nx=181, ny=107, dx=30000.0))
    lat, lon = host.target_coords(c, code)
    wlat, wlon = po.lc_grid(181, 107, 30000.0, 38.5, -97.5, 38.5, 38.5, -97.5, stag)
    assert lat.shape == wlat.shape
    np.testing.assert_allclose(lat, wlat, rtol=0, atol=1e-11)
    np.testing.assert_allclose(lon, wlon, rtol=0, atol=1e-11)
    if stag == "M":
        # reference point sits at (known_x, known_y) = grid centre; spacing ~ dx at the true latitude
        assert abs(lat[52:54, 89:91].mean() - 38.5) < 0.01 and abs(lon[52:54, 89:91].mean() + 97.5) < 0.01
        xyz = np.stack([np.cos(np.radians(lat)) * np.cos(np.radians(lon)), np.cos(np.radians(lat)) * np.sin(np.radians(lon)),
                        np.sin(np.radians(lat))], -1)
        d = np.linalg.norm(xyz[53, 90] - xyz[53, 89]) * 6370000.0
        assert abs(d / 30000.0 - 1) < 2e-3
        assert lon.min() >= -180 and lon.max() <= 180


def test_global_latlon_target_coords(host, tmp_path):
    from oracle import proj_oracle as po

    f = tmp_path / "nl"
    f.write_text("&config target_grid_type='lat-lon' nx=361 ny=181 is_regional=.false. stand_lon=-180. /")
    c = host.read_setup_namelist(str(f))
    lat, lon = host.target_coords(c, 0)
    assert lat.shape == (180, 360)
    assert lat[0, 0] == -89.5 and lat[-1, 0] == 89.5 and lon[0, 0] == -179.5 and lon[0, -1] == 179.5
    wlat, wlon = po.latlon_global_grid(361, 181, -180.0)
    np.testing.assert_allclose(lat, wlat, atol=1e-12)
    np.testing.assert_allclose(lon, wlon, atol=1e-12)
    clat, clon = host.target_coords(c, 3)
    assert clat.shape == (181, 361) and clat[0, 0] == -90.0 and clon[0, 0] == -180.0


def test_rotang_matches_restatement_and_is_small_near_stand_lon(host, tmp_path):
    from oracle import proj_oracle as po

    c = host.read_setup_namelist(defaults.write_namelist(str(tmp_path / "nl"), nx=181, ny=107, dx=30000.0))
    lat, lon = host.target_coords(c, 0)
    cosa, sina = host.get_rotang(lat, lon)
    wc, ws = po.rotang(lat, lon)
    np.testing.assert_allclose(cosa, wc, atol=1e-13)
    np.testing.assert_allclose(sina, ws, atol=1e-13)
    assert np.abs(cosa ** 2 + sina ** 2 - 1).max() < 1e-14
    mid = lon.shape[1] // 2
    assert np.abs(sina[:, mid - 1:mid + 1]).max() < 3e-3                    # grid north ~ true north beside stand_lon
    assert (sina[:, -1] < -0.05).all() and (sina[:, 0] > 0.05).all()        # east: columns lean toward larger lon going north
