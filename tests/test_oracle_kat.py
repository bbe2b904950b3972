"""Known-answer tests that pin the CPU oracle (SURVEY.md §8c list (i)-(vii)).

The reference ships no tests or golden vectors, and ESMF is not available here, so
these analytic cases are the only pin ("parity unpinned", see oracle header)."""
import numpy as np
import pytest

from tests import helpers as H


def _xyz(lon_deg, lat_deg):
    lo, la = np.radians(lon_deg), np.radians(lat_deg)
    return np.stack([np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.sin(la)], -1)


def test_coordinate_conventions(orc):
    # model_grid.F90:450-454: degrees, wrap > 180 by -360 (exactly 180 stays)
    lon = np.array([0.0, np.pi / 2, np.pi, 1.5 * np.pi, 2 * np.pi - 1e-9])
    lat = np.array([0.0, 0.5, -0.5, 1.0, -1.0])
    lo, la = orc.mesh_rad_to_deg(lon, lat)
    assert np.allclose(lo, [0, 90, 180, -90, -1e-9 * 180 / np.pi], atol=1e-12)
    assert lo[2] == 180.0
    assert np.allclose(la, np.degrees(lat), atol=1e-13)
    xyz = orc.sph_deg_to_cart(lo, la)
    assert np.allclose(xyz, _xyz(lo, la), atol=1e-15)
    assert np.allclose(np.linalg.norm(xyz, axis=1), 1.0, atol=1e-15)


def _three_cell_mesh():
    """One Delaunay triangle: three cells around a single shared vertex."""
    ang = np.radians([90.0, 210.0, 330.0])
    r = 2.0  # degrees
    clon, clat = r * np.cos(ang), r * np.sin(ang)
    cxyz = _xyz(clon, clat)
    voc = np.array([[1, 0, 0], [1, 0, 0], [1, 0, 0]], np.int32)  # each cell touches vertex 1 only
    return cxyz, voc


def test_single_triangle_kat(orc):
    cxyz, voc = _three_cell_mesh()
    tri = orc.dual_triangles(voc, 1)
    assert tri.tolist() == [[0, 1, 2]]
    cen = cxyz.mean(0)
    cen /= np.linalg.norm(cen)
    pts = np.stack([cen, cxyz[0], _xyz(np.array(10.0), np.array(0.0))])
    elem, col, w = orc.bilinear(cxyz, tri, voc, pts, brute=True)
    assert elem.tolist() == [0, 0, -1]                      # centroid in, vertex in, far point unmapped
    assert np.allclose(w[0], 1 / 3, atol=1e-12)             # equilateral: exactly 1/3 each
    assert np.allclose(w[1], [1, 0, 0], atol=1e-12)         # at a corner: (1,0,0)
    assert col[2].tolist() == [-1, -1, -1] and np.all(w[2] == 0)
    # just outside an edge (beyond the 1e-10 parametric tolerance) -> unmapped; just inside -> mapped
    mid = cxyz[0] + cxyz[1]
    mid /= np.linalg.norm(mid)
    out = mid + 1e-6 * (mid - cen)
    inn = mid - 1e-6 * (mid - cen)
    e2, _, w2 = orc.bilinear(cxyz, tri, voc, np.stack([out / np.linalg.norm(out), inn / np.linalg.norm(inn)]), brute=True)
    assert e2.tolist() == [-1, 0]
    assert abs(w2[1, 2]) < 1e-5 and abs(w2[1, 0] - 0.5) < 1e-5


def test_boundary_vertices_have_no_dual_element(orc):
    mesh = H.small_regional(800)
    tri = orc.dual_triangles(mesh.verticesOnCell, mesh.nVertices)
    deg = np.bincount(mesh.verticesOnCell[mesh.verticesOnCell > 0] - 1, minlength=mesh.nVertices)
    assert np.array_equal(tri[:, 0] >= 0, deg == 3)
    assert (deg < 3).any()
    assert np.all(np.diff(tri[tri[:, 0] >= 0], axis=1) > 0)   # ascending corners


@pytest.mark.parametrize("mk", ["global", "regional"])
def test_accelerated_search_equals_brute_force(orc, mk):
    mesh = H.small_global(2562, 0.2) if mk == "global" else H.small_regional(2500)
    cxyz, vxyz, tri = H.oracle_geometry(orc, mesh)
    if mk == "global":
        lon, lat = H.latlon_grid(96, 48)
    else:
        lon, lat = H.latlon_grid(90, 70, -103.0, -92.0, 35.5, 41.5)   # overhangs the mesh: far-outside points
    dxyz = orc.sph_deg_to_cart(lon, lat)
    assert np.array_equal(orc.nearest(cxyz, dxyz), orc.nearest(cxyz, dxyz, brute=True))
    e1, c1, w1 = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    e2, c2, w2 = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz, brute=True)
    assert np.array_equal(e1, e2) and np.array_equal(c1, c2) and np.array_equal(w1, w2)
    if mk == "regional":
        assert (e1 < 0).any() and (e1 >= 0).any()


def test_nearest_matches_numpy_argmin_and_tie_rule(orc):
    mesh = H.small_global(642, 0.1)
    cxyz, _, _ = H.oracle_geometry(orc, mesh)
    lon, lat = H.latlon_grid(40, 20)
    dxyz = orc.sph_deg_to_cart(lon, lat)
    d2 = ((dxyz[:, None, :] - cxyz[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(orc.nearest(cxyz, dxyz, brute=True), d2.argmin(1))
    # exact tie: two mirror-image sources equidistant from the target -> smallest index
    src = _xyz(np.array([10.0, -10.0, 50.0]), np.array([0.0, 0.0, 0.0]))
    dst = _xyz(np.array([0.0]), np.array([0.0]))
    assert orc.nearest(src, dst, brute=True).tolist() == [0]
    assert orc.nearest(src[[1, 0, 2]], dst).tolist() == [0]


def test_bilinear_reproduces_constant_and_linear_fields(orc):
    mesh = H.small_global(2562, 0.15)
    cxyz, _, tri = H.oracle_geometry(orc, mesh)
    lon, lat = H.latlon_grid(90, 45)
    dxyz = orc.sph_deg_to_cart(lon, lat)
    elem, col, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    assert (elem >= 0).all()                                 # closed sphere: everything mapped
    assert np.abs(w.sum(1) - 1).max() < 1e-14                # constant field -> constant
    assert w.min() > -1e-10
    g = np.array([0.3, -0.2, 0.9])
    q = (w[:, :, None] * cxyz[col]).sum(1)                   # gnomonic image point on the flat triangle
    assert np.abs(q / np.linalg.norm(q, axis=1, keepdims=True) - dxyz).max() < 1e-14
    assert np.abs((w * (cxyz @ g)[col]).sum(1) - q @ g).max() < 1e-14
    rp, cc, ww = orc.ell_to_csr(elem >= 0, col, w)
    out = orc.apply(rp, cc, ww, np.full((mesh.nCells, 3), 7.25, np.float32))
    assert np.all(out == np.float32(7.25))


def test_smallest_element_id_wins_on_shared_edges_and_corners(orc):
    mesh = H.small_global(642, 0.0)
    cxyz, _, tri = H.oracle_geometry(orc, mesh)
    # targets exactly at cell centres: every dual triangle around the cell accepts them
    elem, col, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, cxyz[:200], brute=True)
    for c in range(200):
        around = mesh.verticesOnCell[c][mesh.verticesOnCell[c] > 0] - 1
        assert elem[c] == around.min()
        assert np.isclose(w[c][col[c] == c][0], 1.0, atol=1e-12)


def test_apply_zero_fills_and_is_linear(orc):
    rng = np.random.default_rng(0)
    nSrc, nDst = 300, 200
    lens = rng.integers(0, 6, nDst)
    rp = np.zeros(nDst + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    col = rng.integers(0, nSrc, rp[-1]).astype(np.int32)
    w = rng.random(rp[-1])
    a = rng.standard_normal((nSrc, 5))
    b = rng.standard_normal((nSrc, 5))
    oa, ob = orc.apply(rp, col, w, a, np.float64), orc.apply(rp, col, w, b, np.float64)
    assert np.all(oa[:, lens == 0] == 0)
    assert np.allclose(orc.apply(rp, col, w, 2 * a - b, np.float64), 2 * oa - ob, atol=1e-12)
    import scipy.sparse as sp

    W = sp.csr_matrix((w, col, rp), shape=(nDst, nSrc))
    assert np.allclose(oa, (W @ a).T, atol=1e-12)
    assert np.array_equal(orc.apply(rp, col, w, a.astype(np.float32), np.float32, tiled=True),
                          orc.apply(rp, col, w, a.astype(np.float32), np.float32))


def test_rotate_winds_kat(orc):
    # interp.F90:739-745: identity for alpha = 0; closed form with the SEQUENTIAL u'-then-v' rule
    n = 7
    u = np.arange(n, dtype=np.float64).reshape(1, n) + 1
    v = 2 * u
    u0, v0 = orc.rotate_winds(u.copy(), v.copy(), np.ones(n), np.zeros(n))
    assert np.array_equal(u0, u) and np.array_equal(v0, v)
    al = 0.3
    ca, sa = np.full(n, np.cos(al)), np.full(n, np.sin(al))
    u1, v1 = orc.rotate_winds(u.copy(), v.copy(), ca, sa)
    t = sa / ca
    ue = (u + v * t) / (ca + sa * t)
    assert np.allclose(u1, ue, rtol=1e-15) and np.allclose(v1, (v - ue * sa) / ca, rtol=1e-15)
    # algebraically: u' = u cos + v sin, v' = v cos - u sin  (earth -> grid rotation by alpha)
    assert np.allclose(u1, u * ca + v * sa, rtol=1e-13) and np.allclose(v1, v * ca - u * sa, rtol=1e-13)


# ---------------------------------------------------------------------------
# global targets: ESMF_GridCreate1PeriDim + MONOPOLE (model_grid.F90:684-696)
# ---------------------------------------------------------------------------
def _global_staggers(ni=24, nj=12):
    """Cell-centred global lat-lon target the way the reference lays it out (program_setup.F90:197-210):
    centres (i + .5) dlon, U at i dlon (ni + 1 columns, first and last on the seam), V at rows -90 ... +90."""
    dlon, dlat = 360.0 / ni, 180.0 / nj
    lonc = -180.0 + (np.arange(ni) + 0.5) * dlon
    latc = -90.0 + (np.arange(nj) + 0.5) * dlat
    lonu = -180.0 + np.arange(ni + 1) * dlon
    latv = -90.0 + np.arange(nj + 1) * dlat
    M = np.meshgrid(lonc, latc)
    U = np.meshgrid(lonu, latc)
    V = np.meshgrid(lonc, latv)
    return M, U, V


def test_periodic_grid_maps_the_seam_columns(orc):
    """Centre -> EDGE1 on a periodic grid: the quad between the last and the first centre column exists, so the two
    seam columns of U are interpolated across the seam (the same weights, on columns ni-1 and 0) instead of being
    left unmapped as on a GridCreateNoPeriDim grid."""
    ni, nj = 24, 12
    (lonM, latM), (lonU, latU), _ = _global_staggers(ni, nj)
    sx = orc.sph_deg_to_cart(lonM, latM).reshape(nj, ni, 3)
    dU = orc.sph_deg_to_cart(lonU, latU)
    e0, c0, w0 = orc.bilinear_quadgrid(sx, dU)                               # non-periodic
    assert (e0.reshape(nj, ni + 1)[:, [0, ni]] < 0).all() and (e0.reshape(nj, ni + 1)[:, 1:ni] >= 0).all()
    topo = orc.TOPO_PERI | orc.TOPO_SPOLE | orc.TOPO_NPOLE
    e, c, w = orc.bilinear_quadgrid(sx, dU, topo=topo)
    eb, cb, wb = orc.bilinear_quadgrid(sx, dU, topo=topo, brute=True)        # kd-tree search == brute force
    assert np.array_equal(e, eb) and np.array_equal(c, cb) and np.array_equal(w, wb)
    assert (e >= 0).all() and (e < ni * (nj - 1)).all()                      # every U point sits in a quad
    rp, cc, ww = orc.quadgrid_csr(ni, nj, e, c, w, topo)
    assert np.array_equal(np.diff(rp), np.full(dU.shape[0], 4))
    W = ww.reshape(nj, ni + 1, 4)
    Cc = cc.reshape(nj, ni + 1, 4)
    assert np.abs(W.sum(-1) - 1).max() <= 1e-12
    # interior columns agree with the non-periodic matrix (same quads, same ids order within a row)
    assert np.array_equal(c.reshape(nj, ni + 1, 4)[:, 1:ni], c0.reshape(nj, ni + 1, 4)[:, 1:ni])
    assert np.array_equal(w.reshape(nj, ni + 1, 4)[:, 1:ni], w0.reshape(nj, ni + 1, 4)[:, 1:ni])
    # seam: column 0 and column ni are the same point, mapped by the wrap quad onto columns ni-1 and 0
    assert np.array_equal(Cc[:, 0], Cc[:, ni]) and np.abs(W[:, 0] - W[:, ni]).max() <= 1e-12
    cols = Cc[:, 0] % ni
    heavy = np.take_along_axis(cols, np.argsort(-W[:, 0], axis=1)[:, :2], axis=1)
    assert all(set(r) == {ni - 1, 0} for r in heavy.tolist())
    # the interpolated position t p is parallel to p: weights reproduce Cartesian coordinates along the ray
    pos = (W[..., None] * sx.reshape(-1, 3)[Cc]).sum(2).reshape(-1, 3)
    assert np.abs(np.cross(pos, dU)).max() <= 1e-12
    # a zonally periodic field is continuous across the seam
    f = (np.cos(np.radians(latM)) * np.cos(np.radians(lonM))).reshape(1, -1)
    u = orc.apply_planes(rp, cc, ww, f).reshape(nj, ni + 1)
    assert np.abs(u[:, 0] - u[:, ni]).max() <= 1e-13
    assert np.abs(u[:, 0] - np.cos(np.radians(latU[:, 0])) * np.cos(np.radians(-180.0))).max() <= 0.01


def test_monopole_caps_average_the_end_rows(orc):
    """Centre -> EDGE2: the V rows at +-90 lie at the artificial pole node, whose value is the average of the
    neighbouring centre row (polemethod ALLAVG): ni entries of 1/ni.  Interior rows are untouched."""
    ni, nj = 24, 12
    (lonM, latM), _, (lonV, latV) = _global_staggers(ni, nj)
    sx = orc.sph_deg_to_cart(lonM, latM).reshape(nj, ni, 3)
    dV = orc.sph_deg_to_cart(lonV, latV)
    e0, c0, w0 = orc.bilinear_quadgrid(sx, dV)
    assert (e0.reshape(nj + 1, ni)[[0, nj]] < 0).all()                        # no caps: pole rows unmapped
    topo = orc.TOPO_PERI | orc.TOPO_SPOLE | orc.TOPO_NPOLE
    e, c, w = orc.bilinear_quadgrid(sx, dV, topo=topo)
    E = e.reshape(nj + 1, ni)
    nq = ni * (nj - 1)
    assert (E[1:nj] >= 0).all() and (E[1:nj] < nq).all()
    assert (E[0] >= nq).all() and (E[0] < nq + ni).all() and (E[nj] >= nq + ni).all()
    rp, cc, ww = orc.quadgrid_csr(ni, nj, e, c, w, topo)
    cnt = np.diff(rp).reshape(nj + 1, ni)
    assert (cnt[[0, nj]] == ni).all() and (cnt[1:nj] == 4).all()
    for row, base in ((0, 0), (nj, (nj - 1) * ni)):
        for i in range(ni):
            t = row * ni + i
            assert np.array_equal(cc[rp[t]:rp[t + 1]], base + np.arange(ni))
            assert np.abs(ww[rp[t]:rp[t + 1]] - 1.0 / ni).max() <= 1e-12
    f = (3.0 + np.sin(np.radians(latM)) + 0.3 * np.cos(np.radians(lonM))).reshape(1, -1)
    v = orc.apply_planes(rp, cc, ww, f).reshape(nj + 1, ni)
    assert np.abs(v[0] - f.reshape(nj, ni)[0].mean()).max() <= 1e-12
    assert np.abs(v[nj] - f.reshape(nj, ni)[nj - 1].mean()).max() <= 1e-12
    # a row block that does not hold the pole row carries no cap
    eb, _, _ = orc.bilinear_quadgrid(sx[2:6], dV.reshape(nj + 1, ni, 3)[3:6].reshape(-1, 3), topo=orc.TOPO_PERI)
    assert (eb >= 0).all()
