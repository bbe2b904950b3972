"""The C oracle against the committed golden vectors (tests/golden/lc_small.npz).

The vectors come from an independent numpy restatement of the same ESMF rules with different
algorithms (tests/golden/make_golden.py); agreement of the two pins each against coding
errors.  Neither is ESMF: parity with the real reference remains unpinned.
Indices, masks and row structure must be identical; weights agree to 1e-12 (bilinear, nearest,
quad) / 1e-9 (conservative: two different spherical-area formulas)."""
import os

import numpy as np
import pytest

from tests import helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lc_small.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def geom(orc, g):
    lo, la = orc.mesh_rad_to_deg(g["lonCell"], g["latCell"])
    lov, lav = orc.mesh_rad_to_deg(g["lonVertex"], g["latVertex"])
    return orc.sph_deg_to_cart(lo, la), orc.sph_deg_to_cart(lov, lav), orc.sph_deg_to_cart(g["lon_M"], g["lat_M"])


def test_fixture_is_reproducible(g, tmp_path):
    """The committed file is what the committed script writes (inputs and integer results exactly)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mesh, grids = mg.build_case()
    assert np.array_equal(mesh.verticesOnCell.astype(np.int32), g["verticesOnCell"])
    assert np.array_equal(mesh.lonCell, g["lonCell"]) and np.array_equal(grids["M"][0], g["lat_M"])
    lo, la = mg.mesh_deg(mesh.lonCell, mesh.latCell)
    assert np.array_equal(mg.nearest(mg.cart(lo, la), mg.cart(grids["M"][1], grids["M"][0])), g["nearest_idx"])


def test_dual_triangles_and_nearest(orc, g, geom):
    cxyz, vxyz, dM = geom
    tri = orc.dual_triangles(g["verticesOnCell"], g["lonVertex"].size)
    assert np.array_equal(tri, g["tri"])
    assert np.array_equal(orc.nearest(cxyz, dM), g["nearest_idx"])
    assert np.array_equal(orc.nearest(cxyz, dM, brute=True), g["nearest_idx"])


def test_bilinear(orc, g, geom):
    cxyz, vxyz, dM = geom
    e, c, w = orc.bilinear(cxyz, g["tri"], g["verticesOnCell"], dM)
    assert np.array_equal(e, g["bil_elem"])            # which dual element, incl. the unmapped mask
    assert (e < 0).sum() == 70 and (e >= 0).sum() == 122
    m = e >= 0
    assert np.array_equal(c[m], g["bil_col"][m])
    assert np.abs(w[m] - g["bil_w"][m]).max() <= 1e-12


def test_quadgrid(orc, g, geom):
    cxyz, vxyz, dM = geom
    sx = dM.reshape(*g["lat_M"].shape, 3)
    for s in ("U", "V"):
        e, c, w = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(g[f"lon_{s}"], g[f"lat_{s}"]))
        assert np.array_equal(e, g[f"quad{s}_elem"])
        m = e >= 0
        assert np.array_equal(c[m], g[f"quad{s}_col"][m])
        assert np.abs(w[m] - g[f"quad{s}_w"][m]).max() <= 1e-12


def _drop_slivers(rp, c, w, eps=1e-12):
    keep = w > eps
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))[keep]
    return rows, c[keep], w[keep]


def test_conserve(orc, g, geom):
    cxyz, vxyz, dM = geom
    cor = orc.sph_deg_to_cart(g["lon_CORNER"], g["lat_CORNER"]).reshape(*g["lat_CORNER"].shape, 3)
    rp, c, w = orc.conserve(cxyz, vxyz, g["verticesOnCell"], cor)
    r1, c1, w1 = _drop_slivers(rp, c, w)
    r2, c2, w2 = _drop_slivers(g["cons_rowptr"], g["cons_col"], g["cons_w"])
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2)
    assert np.abs(w1 - w2).max() <= 1e-9
    assert w.size - w1.size <= 4 and g["cons_w"].size - w2.size <= 4  # slivers are rare


def test_node_bilinear(orc, g, geom):
    cxyz, vxyz, dM = geom
    e, c, w = orc.bilinear_node(cxyz, vxyz, g["verticesOnCell"], dM)
    assert np.array_equal(e, g["node_elem"])
    m = e >= 0
    assert np.array_equal(c[m], g["node_col"][m])
    assert np.abs(w[m] - g["node_w"][m]).max() <= 1e-12


def test_applied_fields(orc, g, geom):
    cxyz, vxyz, dM = geom
    e, c, w = orc.bilinear(cxyz, g["tri"], g["verticesOnCell"], dM)
    bil = orc.ell_to_csr(e >= 0, c, w)
    got = orc.apply(*bil, g["src_theta"], np.float32)
    # fp64 sums in a different order: at most one fp32 ulp apart after the single rounding
    np.testing.assert_allclose(got, g["dst_theta"], rtol=1.2e-7, atol=0)
    assert np.array_equal(got[:, e < 0], np.zeros((5, 70), np.float32))           # unmapped -> 0.0
    nst = orc.nearest_to_csr(orc.nearest(cxyz, dM))
    assert np.array_equal(orc.apply(*nst, g["src_xland"], np.float32), g["dst_xland"])  # bit-exact
    um, vm = orc.apply(*bil, g["src_u"], np.float64), orc.apply(*bil, g["src_v"], np.float64)
    orc.rotate_winds(um, vm, g["cosa"].reshape(-1), g["sina"].reshape(-1))
    np.testing.assert_allclose(um, g["dst_umass"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(vm, g["dst_vmass"], rtol=1e-13, atol=1e-13)
    sx = dM.reshape(*g["lat_M"].shape, 3)
    for s, f in (("U", um), ("V", vm)):
        eq, cq, wq = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(g[f"lon_{s}"], g[f"lat_{s}"]))
        got = orc.apply_planes(*orc.ell_to_csr(eq >= 0, cq, wq), f).astype(np.float32)
        np.testing.assert_allclose(got, g[f"dst_{s}"], rtol=1.2e-7, atol=1e-7)


def test_host_mirror_reproduces_fixture_grid(g, tmp_path):
    """The C++ host mirror's Lambert coordinates equal the fixture's (numpy restatement) to 1e-10 deg."""
    from mpassit_b200 import build, defaults, host

    build.build_host()
    host.load()
    cfg = host.read_setup_namelist(defaults.write_namelist(str(tmp_path / "namelist.input"), nx=17, ny=13, dx=40000.0))
    for s, code in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3)):
        lat, lon = host.target_coords(cfg, code)
        assert lat.shape == g[f"lat_{s}"].shape
        assert np.abs(lat - g[f"lat_{s}"]).max() < 1e-10 and np.abs(lon - g[f"lon_{s}"]).max() < 1e-10
