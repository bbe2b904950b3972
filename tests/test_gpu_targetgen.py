"""Target-grid coordinates, map factors and rotation angles generated on the device (SURVEY.md 8 row f3;
mprg_set_target_projected / mprg_target_map_factor / mprg_set_rotation_from_target) against the host mirror of the
reference's WPS projection code (mpassit_b200/host/target_grid.cpp, itself held to oracle/proj_oracle.py).

Device fp64 trig is within 1-2 ulp of the host libm, not bit-identical; what matters for parity is that the weight
matrices keep their STRUCTURE (mapped mask, indices): checked on BASELINE configs[0] (global lat-lon), the 12-km
miniature of configs[1] and configs[4] (1-km conservative) -- configs[1] at full size is in test_gpu_fullsize.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulps(a, b):
    """|a - b| in units of the last place of b (fp64)."""
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(b), 1e-300))


def compare_target_generation(wl, routes, min_weight_digits=1e-9):
    from mpassit_b200 import host
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    m = wl.mesh
    proj = host.projection(wl.cfg)
    staggers = [(k, s) for k, s in (("M", l.CENTER), ("U", l.EDGE1), ("V", l.EDGE2), ("CORNER", l.CORNER)) if k in wl.grids]
    out = {"ulp_lon": 0.0, "ulp_lat": 0.0, "structure_diffs": 0, "rows": 0}
    rh, rd = Regridder(device=0), Regridder(device=0)
    for rg in (rh, rd):
        rg.set_mesh(m.lonCell, m.latCell, m.lonVertex, m.latVertex, m.verticesOnCell)
        rg.set_grid_kind(l.GRID_NOPERI if wl.cfg.is_regional else l.GRID_1PERI_MONOPOLE)
    for k, s in staggers:
        lat, lon = wl.grids[k]
        rh.set_target(s, lon, lat)
        rd.set_target_projected(s, lon.shape[1], lon.shape[0], proj)
        dlon, dlat = rd.target_lonlat(s)
        out["ulp_lon"] = max(out["ulp_lon"], float(ulps(dlon, lon).max()))
        out["ulp_lat"] = max(out["ulp_lat"], float(ulps(dlat, lat).max()))
    for key in routes:
        a, b = rh.store(*key), rd.store(*key)
        ra, ca, wa = a.export_csr()
        rb, cb, wb = b.export_csr()
        out["rows"] += ra.size - 1
        same = np.array_equal(ra, rb) and np.array_equal(ca, cb)
        if not same:
            if np.array_equal(ra, rb):     # same mapped mask: count the rows whose indices differ
                bad = np.flatnonzero(ca != cb)
                out["structure_diffs"] += int(np.unique(np.searchsorted(ra, bad, side="right") - 1).size)
            else:
                out["structure_diffs"] += int((np.diff(ra) != np.diff(rb)).sum())
        else:
            # conservative weights are ratios of areas of ~1e-8 steradian cells: ulp-level coordinate differences are
            # amplified to ~1e-9 by the cancellation in the area sums (the oracle shows the same conditioning)
            assert np.abs(wa - wb).max() <= (2e-8 if key[0] == l.CONSERVE else min_weight_digits), key
        a.release(); b.release()
    # map factors and rotation angles: no index decision depends on them
    if wl.cfg.proj_code == host.PROJ_LC:
        lat, lon = wl.grids["M"]
        mf_host = host.get_map_factor(wl.cfg, lat)
        mf_dev = rd.target_map_factor(l.CENTER, 1, wl.cfg.truelat1, wl.cfg.truelat2)
        assert np.abs(mf_dev - mf_host).max() <= 1e-13
        ca, sa = rd.set_rotation_from_target()
        # (alpha comes from differences of neighbouring coordinates: ulp-level coordinate differences are amplified)
        assert np.abs(ca - wl.cosa).max() <= 1e-10 and np.abs(sa - wl.sina).max() <= 1e-10
        assert rd.has_rotation() if hasattr(rd, "has_rotation") else True
    rh.close(); rd.close()
    return out


@pytest.mark.parametrize("name", ["c1", "mid", "c5"])
def test_device_generated_targets_match_the_host_mirror(engine_lib, name):
    from mpassit_b200 import build, workload
    from mpassit_b200 import lib as l

    build.build_host()
    wl = workload.make(name)
    routes = [(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER), (l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER),
              (l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER), (l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE1),
              (l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE2)]
    res = compare_target_generation(wl, routes)
    print(name, res)
    # coordinates: a few ulp (device sin / cos / atan2 / pow are not the host libm); the lat-lon projection is pure
    # arithmetic and must be exact
    if name == "c1":
        assert res["ulp_lon"] == 0 and res["ulp_lat"] == 0
    else:
        assert res["ulp_lon"] <= 8 and res["ulp_lat"] <= 8, res
    # The matrices keep their structure: same mapped mask, same indices.  The 1-degree global target is the exception
    # that shows why the host path stays the default: its points lie EXACTLY on edges of the icosahedral mesh's
    # triangles and on the poles (exact ties), where a last-ulp difference of the Cartesian coordinates picks the
    # neighbouring element -- a few dozen of ~390,000 rows, each with the same interpolated value to 1e-10.
    # The 1-km conservative case differs in a handful of rows of 5 million (a sliver overlap of ~zero area appears or
    # disappears).  Measured on B200: c1 44 / 388,800 rows, mid 0, c2 0 (test_gpu_fullsize), c5 4 / 5,002,000.
    if name == "c1":
        assert res["structure_diffs"] <= 1e-3 * res["rows"], res
    else:
        assert res["structure_diffs"] <= 1e-5 * res["rows"], res
