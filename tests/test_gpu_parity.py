"""GPU parity: CUDA engine (through the C ABI) vs the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): bit-exact for nearest-neighbour columns, masks and
indices; weights <= 1e-12 abs; fp32 bilinear fields <= 1e-5 relative (here far tighter).
"""
import numpy as np
import pytest

from mpassit_b200 import check
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rg(engine_lib):
    from mpassit_b200.regrid import Regridder

    r = Regridder(device=0)
    yield r
    r.close()


CASES = {
    "global_icos_jit": lambda: (H.small_global(2562, 0.15), H.latlon_grid(144, 72)),
    "global_c1": lambda: (H.synth.global_mesh(40962), H.latlon_grid(360, 180)),
    "regional_ragged": lambda: (H.small_regional(3000), H.latlon_grid(150, 110, -102.0, -93.0, 36.0, 41.0)),
}


def _setup(rg, orc, name):
    mesh, (lon, lat) = CASES[name]()
    rg.set_mesh(mesh.lonCell, mesh.latCell, mesh.lonVertex, mesh.latVertex, mesh.verticesOnCell)
    rg.set_target(0, lon, lat)
    cxyz, vxyz, tri = H.oracle_geometry(orc, mesh)
    dxyz = orc.sph_deg_to_cart(lon, lat)
    return mesh, lon, lat, cxyz, vxyz, tri, dxyz


@pytest.mark.parametrize("name", list(CASES))
def test_nearest_indices_bit_exact(rg, orc, name):
    from mpassit_b200 import lib as l

    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, name)
    want = orc.nearest(cxyz, dxyz, brute=(name != "global_c1"))
    r = rg.store(l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER)
    rowptr, col, w = r.export_csr()
    assert np.array_equal(rowptr, np.arange(lon.size + 1))
    assert np.array_equal(col, want)
    assert np.all(w == 1.0)
    assert r.info()["nUnmapped"] == 0
    r.release()


@pytest.mark.parametrize("name", list(CASES))
def test_bilinear_structure_and_weights(rg, orc, name):
    from mpassit_b200 import lib as l

    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, name)
    elem, col, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz, brute=(name != "global_c1"))
    mask = elem >= 0
    rp_want, col_want, w_want = orc.ell_to_csr(mask, col, w)
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    rowptr, gcol, gw = r.export_csr()
    assert np.array_equal(rowptr, rp_want)          # mapped / unmapped mask bit-exact
    assert np.array_equal(gcol, col_want)           # indices bit-exact
    assert np.abs(gw - w_want).max() <= 1e-12 if gw.size else True
    assert r.info()["nUnmapped"] == int((~mask).sum())
    if name == "regional_ragged":
        assert 0 < (~mask).sum() < mask.size        # the case really exercises unmapped rows
    r.release()


@pytest.fixture
def knobs(rg):
    """Tuning options are restored after the test (the engine is shared by the module)."""
    yield rg
    for k, v in (("accumulate", "f32"), ("pipe_split", "1"), ("apply", "pipe"), ("pipe_minb", "0")):
        rg.set_option(k, v)


@pytest.mark.parametrize("acc", ["f32", "f64"])
@pytest.mark.parametrize("nlev", [1, 4, 55, 60, 61, 62, 64, 130])
def test_apply_bilinear_levels(knobs, rg, orc, nlev, acc):
    from mpassit_b200 import lib as l

    rg.set_option("accumulate", acc)   # default is f32 accumulation; f64 = the reference's R8 arithmetic

    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, "regional_ragged")
    elem, col, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    rp, cc, ww = orc.ell_to_csr(elem >= 0, col, w)
    src = H.synth.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=nlev)
    want = orc.apply(rp, cc, ww, src, np.float32)
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    got = np.full((nlev, lon.size), np.nan, np.float32)
    rg.apply(r, [src if nlev > 1 else src.reshape(-1)], [got], nlev=[nlev])
    r.release()
    unm = elem < 0
    assert np.all(got[:, unm] == 0.0)                # zero fill of unmapped rows
    # contract (BASELINE.json north_star): <= 1e-5 relative for fp32 bilinear fields
    check.assert_field_close(got, want, f"nlev={nlev} acc={acc}")   # per element
    if acc == "f64":
        # fp64 accumulation, one rounding: at most 1 ulp from the oracle's rounding of the same sum
        np.testing.assert_allclose(got, want, rtol=2e-7, atol=0)
    else:
        # fp32 FMA accumulation of a convex combination: a few ulp
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)


@pytest.mark.parametrize("split", ["1", "0"])
def test_device_sources_at_any_element_alignment_and_mixed_stacks(knobs, rg, orc, split):
    """Device sources are used in place.  A view that starts 4 bytes into an allocation (base not 16-byte aligned)
    and stacks mixing aligned / unaligned level counts and a wind pair go through ONE column launch and must give
    exactly what each field gives on its own from an aligned copy."""
    import torch

    from mpassit_b200 import lib as l

    rg.set_option("pipe_split", split)   # aligned plain fields in their own launch, or everything in one
    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, "regional_ragged")
    n, nd = mesh.nCells, lon.size
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    levs = [60, 61, 55, 64, 60, 9, 130]
    srcs, views = [], []
    for k, nl in enumerate(levs):
        off = k % 4                                              # 0, 4, 8, 12 bytes into the allocation
        buf = torch.randn(n * nl + 4, generator=g, device="cuda")
        views.append(buf[off:off + n * nl].view(n, nl))
        srcs.append(views[-1].clone())                           # aligned copy of the same values
        assert views[-1].data_ptr() % 16 == 4 * off
    stacked = [torch.full((nl, nd), float("nan"), device="cuda") for nl in levs]
    n0 = rg.kernel_launches
    rg.apply(r, views, stacked, nlev=levs)
    rg.synchronize()
    assert rg.kernel_launches - n0 <= (2 if split == "1" else 1)   # the 3-D fields of an apply take one or two column launches
    for k, nl in enumerate(levs):
        one = torch.full((nl, nd), float("nan"), device="cuda")
        rg.apply(r, [srcs[k]], [one], nlev=[nl])
        rg.synchronize()
        assert torch.equal(one, stacked[k]), (split, nl)
    e, c, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    want = orc.apply(*orc.ell_to_csr(e >= 0, c, w), srcs[1].cpu().numpy(), np.float32)
    check.assert_field_close(stacked[1].cpu().numpy(), want, "61 levels, base + 4 bytes")
    r.release()


def test_apply_nearest_bit_exact_and_stacked(rg, orc):
    from mpassit_b200 import lib as l

    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, "global_icos_jit")
    idx = orc.nearest(cxyz, dxyz)
    rp, cc, ww = orc.nearest_to_csr(idx)
    r = rg.store(l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER)
    xland = H.synth.integer_field(mesh.nCells, 2)
    ivg = H.synth.integer_field(mesh.nCells, 20, seed=3)
    soil = np.ascontiguousarray(np.stack([H.synth.integer_field(mesh.nCells, 9, seed=s) for s in range(4)], 1))
    outs = [np.empty((1, lon.size), np.float32), np.empty((1, lon.size), np.float32), np.empty((4, lon.size), np.float32)]
    rg.apply(r, [xland, ivg, soil], outs, nlev=[1, 1, 4])
    r.release()
    assert np.array_equal(outs[0][0], xland[idx])
    assert np.array_equal(outs[1][0], ivg[idx])
    assert np.array_equal(outs[2], soil[idx].T)
    assert np.array_equal(outs[2], orc.apply(rp, cc, ww, soil, np.float32))


@pytest.mark.parametrize("sdt,ddt", [(np.float32, np.float64), (np.float64, np.float64), (np.float64, np.float32)])
def test_apply_dtype_combinations(rg, orc, sdt, ddt):
    from mpassit_b200 import lib as l

    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, "global_icos_jit")
    elem, col, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    rp, cc, ww = orc.ell_to_csr(elem >= 0, col, w)
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    for nlev in (8, 55):
        src = H.synth.smooth_field(mesh.lonCell, mesh.latCell, nlev, dtype=sdt)
        want = orc.apply(rp, cc, ww, src, ddt)
        got = np.empty((nlev, lon.size), ddt)
        rg.apply(r, [src], [got], nlev=[nlev])
        np.testing.assert_allclose(got, want, rtol=(2e-7 if ddt == np.float32 else 1e-14))
    r.release()


def test_import_csr_ragged_rows_and_empty(rg, orc):
    """Conservative-like ragged rows (0..20 entries) incl. rows longer than the register/smem fast paths."""
    rng = np.random.default_rng(5)
    nSrc, nDst = 5000, 3333
    lens = rng.integers(0, 21, nDst)
    lens[:40] = 0
    lens[100] = 700   # longer than the per-tile CSR cache
    rp = np.zeros(nDst + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    col = rng.integers(0, nSrc, rp[-1]).astype(np.int32)
    w = rng.random(rp[-1])
    r = rg.import_csr(nSrc, rp, col, w)
    for nlev in (1, 7, 64, 65):
        src = rng.standard_normal((nSrc, nlev)).astype(np.float32)
        got = np.empty((nlev, nDst), np.float32)
        rg.apply(r, [src if nlev > 1 else src.reshape(-1)], [got], nlev=[nlev])
        want = orc.apply(rp, col, w, src, np.float32)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)   # up to 700 random-sign terms per row
        assert np.all(got[:, lens == 0] == 0.0)
    r.release()


def test_device_buffers_and_memoised_store(rg, orc):
    import torch

    from mpassit_b200 import lib as l

    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, "global_icos_jit")
    r1 = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    n0 = rg.kernel_launches
    r2 = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)     # memoised: no new kernels
    assert r2.handle == r1.handle and rg.kernel_launches == n0
    src = H.synth.smooth_field(mesh.lonCell, mesh.latCell, 60)
    host = np.empty((60, lon.size), np.float32)
    rg.apply(r1, [src], [host])
    dsrc = torch.from_numpy(src).cuda()
    ddst = torch.empty((60, lon.size), dtype=torch.float32, device="cuda")
    rg.apply(r1, [dsrc], [ddst])
    rg.synchronize()
    assert np.array_equal(ddst.cpu().numpy(), host)
    assert rg.kernel_launches > n0
    r1.release(); r2.release()


def test_rotate_winds(rg, orc):
    mesh, lon, lat, *_ = _setup(rg, orc, "global_icos_jit")
    rng = np.random.default_rng(3)
    n = lon.size
    alpha = rng.uniform(-0.4, 0.4, n)
    cosa, sina = np.cos(alpha), np.sin(alpha)
    rg.set_rotation(cosa.reshape(lon.shape), sina.reshape(lon.shape))
    for dt, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        u = rng.standard_normal((5, n)).astype(dt) * 10
        v = rng.standard_normal((5, n)).astype(dt) * 10
        wu, wv = orc.rotate_winds(u.copy(), v.copy(), cosa, sina)
        rg.rotate_winds(u, v, 5)
        np.testing.assert_allclose(u, wu, rtol=tol, atol=tol)
        np.testing.assert_allclose(v, wv, rtol=tol, atol=tol)


def test_errors_are_reported_not_fatal(rg):
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import MprgError

    with pytest.raises(MprgError):
        rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CORNER if l.CORNER not in rg.shape else 7)
    with pytest.raises(MprgError):
        rg.import_csr(10, np.array([0, 2], np.int32), np.array([1, 11], np.int32), np.ones(2))


@pytest.mark.parametrize("nlev", [1, 4, 60, 61])
def test_host_sources_upload_only_the_referenced_id_range(rg, orc, nlev):
    """Source halo-sharding: a host-buffer apply copies just the cell-id range [lo, hi) that the route's
    weights reference; rows outside it are never read (here: NaN) and the byte counter says so."""
    rng = np.random.default_rng(9)
    nSrc, nDst, lo, hi = 6000, 2500, 1201, 3456
    lens = rng.integers(0, 4, nDst)
    rp = np.zeros(nDst + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    col = rng.integers(lo, hi, rp[-1]).astype(np.int32)
    col[:2] = (lo, hi - 1)
    w = rng.random(rp[-1])
    r = rg.import_csr(nSrc, rp, col, w)
    src = rng.standard_normal((nSrc, nlev)).astype(np.float32)
    want = orc.apply(rp, col, w, src, np.float32)
    src[:lo] = np.nan
    src[hi:] = np.nan
    got = np.empty((nlev, nDst), np.float32)
    b0 = rg.io_bytes()
    rg.apply(r, [src if nlev > 1 else src.reshape(-1)], [got], nlev=[nlev])
    b1 = rg.io_bytes()
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    assert b1[0] - b0[0] == (hi - lo) * nlev * 4 and b1[1] - b0[1] == nDst * nlev * 4
    r.release()


def test_wrf_post_ops_match_the_writer_formulas(rg, orc):
    """write_data.F90:1339-1432 (`wrf_mod_vars`): T-300 and PHB = 9.81 zgrid as fused epilogues of the apply,
    Z_C mid-level average and this rank's share of P_TOP as mprg_post_*, against oracle/post_oracle.py."""
    import torch

    from mpassit_b200 import lib as l
    from oracle import post_oracle as po

    mesh, lon, lat, cxyz, vxyz, tri, dxyz = _setup(rg, orc, "regional_ragged")
    elem, col, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    rp, cc, ww = orc.ell_to_csr(elem >= 0, col, w)
    nz = 9
    theta = H.synth.smooth_field(mesh.lonCell, mesh.latCell, nz, seed=1) + 20.0
    zgrid = np.ascontiguousarray(np.cumsum(np.abs(H.synth.smooth_field(mesh.lonCell, mesh.latCell, nz + 1, seed=2)), axis=1))
    pres = np.ascontiguousarray((1.0e5 * np.exp(-np.arange(nz)[None, :] / 3.0) *
                                 (1 + 0.01 * H.synth.smooth_field(mesh.lonCell, mesh.latCell, nz, seed=3) / 300)).astype(np.float32))
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    n = lon.size
    raw_t, raw_z, raw_p = (orc.apply(rp, cc, ww, a, np.float32) for a in (theta, zgrid, pres))
    t300 = np.empty((nz, n), np.float32)
    phb = np.empty((nz + 1, n), np.float32)
    rg.apply(r, [theta], [t300], nlev=[nz], epi_op=[l.EPI_ADD], epi_arg=[-300.0])
    rg.apply(r, [zgrid], [phb], nlev=[nz + 1], epi_op=[l.EPI_MUL], epi_arg=[9.81])
    np.testing.assert_allclose(t300, po.t_minus_300(raw_t), rtol=1e-6, atol=1e-4)
    assert np.all(t300[:, elem < 0] == -300.0)              # quirk: unmapped T is -300, not 0
    np.testing.assert_allclose(phb, po.phb(raw_z), rtol=1e-6, atol=1e-4)
    # Z_C and P_TOP on host and on device buffers
    zreg = np.empty((nz + 1, n), np.float32)
    preg = np.empty((nz, n), np.float32)
    rg.apply(r, [zgrid, ], [zreg], nlev=[nz + 1])
    rg.apply(r, [pres], [preg], nlev=[nz])
    zc = np.empty((nz, n), np.float32)
    rg.post_midlevels(zreg, zc, nz + 1)
    assert np.array_equal(zc, po.z_c(zreg))
    dz = torch.from_numpy(zreg).cuda()
    dzc = torch.empty((nz, n), dtype=torch.float32, device="cuda")
    rg.post_midlevels(dz, dzc, nz + 1)
    rg.synchronize()
    assert np.array_equal(dzc.cpu().numpy(), zc)
    for buf in (preg, torch.from_numpy(preg).cuda()):
        mx, mn = rg.post_ptop(buf, nz)
        assert mx == float(preg.astype(np.float64).max())
        assert min(mx, mn) == po.p_top(preg)
    # a field whose top level is everywhere below the threshold: only the maxval term survives
    small = np.full((nz, n), 5.0, np.float32)
    mx, mn = rg.post_ptop(small, nz)
    assert mx == 5.0 and mn == float("inf") and min(mx, mn) == po.p_top(small)
    r.release()


def test_new_entry_points_report_errors(rg, orc):
    """apply_into / ipc / rotation epilogues: misuse is an error code + message, never a crash."""
    import torch

    from mpassit_b200 import lib as l

    mesh, lon, lat, *_ = _setup(rg, orc, "global_icos_jit")
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    n = lon.size
    u = torch.zeros((mesh.nCells, 8), device="cuda")
    full = torch.zeros((8, n), device="cuda")
    with pytest.raises(l.MprgError) as e:   # wind pairs are intermediates, not gathered fields
        rg.apply_into(r, [u, u], [full, full], nlev=[8, 8], epi_op=[l.EPI_ROT_U, l.EPI_ROT_V])
    assert "wind pairs" in str(e.value)
    with pytest.raises(l.MprgError):        # ROT_V without its ROT_U
        rg.apply(r, [u], [full], nlev=[8], epi_op=[l.EPI_ROT_V])
    with pytest.raises(l.MprgError):        # not a device allocation
        rg.ipc_export(12345678)
    imp = rg.import_csr(mesh.nCells, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([1.0]))
    with pytest.raises(l.MprgError) as e:   # an imported CSR has no grid to place its rows in
        rg.apply_into(imp, [u], [full], nlev=[8])
    assert "full-grid" in str(e.value)
    imp.release()
    # the context is still usable afterwards
    out = torch.empty((8, n), device="cuda")
    rg.apply(r, [u], [out], nlev=[8])
    rg.synchronize()
    assert float(out.abs().max()) == 0.0
    r.release()


@pytest.mark.parametrize("order", ["morton", "random"])
def test_results_do_not_depend_on_the_cell_numbering(engine_lib, orc, order):
    """The column kernel fetches runs of consecutively numbered cells with one bulk copy; a renumbered mesh
    (Z-order / random) changes which copies are issued but must give the same output, bit for bit, and
    the same weights as the oracle."""
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    mesh = H.synth.regional_hex_mesh(spacing_m=40000.0, extent_x_m=2400e3, extent_y_m=1600e3, seed=5)
    ren = H.synth.renumber_cells(mesh, order, seed=2)
    lon, lat = H.latlon_grid(96, 64, lon0=-108.0, lon1=-87.0, lat0=32.0, lat1=45.0)
    # match renumbered cells to the originals by coordinates
    key = {(a, b): i for i, (a, b) in enumerate(zip(mesh.lonCell, mesh.latCell))}
    perm = np.array([key[(a, b)] for a, b in zip(ren.lonCell, ren.latCell)])
    src = H.synth.smooth_field(mesh.lonCell, mesh.latCell, 60, seed=8)
    outs, infos = [], []
    for m, f in ((mesh, src), (ren, np.ascontiguousarray(src[perm]))):
        rg = Regridder(device=0)
        rg.set_mesh(m.lonCell, m.latCell, m.lonVertex, m.latVertex, m.verticesOnCell)
        rg.set_target(l.CENTER, lon, lat)
        r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
        infos.append(r.info())
        out = np.empty((60, lon.size), np.float32)
        rg.apply(r, [f], [out], nlev=[60])
        outs.append(out)
        if m is ren:
            cxyz, vxyz, tri = H.oracle_geometry(orc, m)
            e, c, w = orc.bilinear(cxyz, tri, m.verticesOnCell, orc.sph_deg_to_cart(lon, lat))
            rp, cc, ww = orc.ell_to_csr(e >= 0, c, w)
            grp, gc, gw = r.export_csr()
            assert np.array_equal(grp, rp) and np.array_equal(gc, cc) and np.abs(gw - ww).max() <= 1e-12
        r.release()
        rg.close()
    # same three cells, same weights; the order of the three products inside a row follows the cell ids
    np.testing.assert_allclose(outs[1], outs[0], rtol=3e-7, atol=0)
    assert infos[0]["tile_columns"] == infos[1]["tile_columns"]
    assert infos[0]["tile_runs"] < infos[1]["tile_runs"] <= infos[1]["tile_columns"]  # row-major numbering has the longest runs


@pytest.mark.parametrize("order", ["rowmajor", "morton", "random"])
def test_host_sources_upload_only_the_referenced_id_ranges(engine_lib, order):
    """Host-buffer applies move only the id ranges a rank's weights reference (route.srcRanges): on a row slab of a
    mesh numbered along a Z-order curve that is a short list of ranges -- a fraction of the field -- where the single
    enclosing range [srcLo, srcHi) would be most of it.  Results equal the device-buffer apply bit for bit, gaps
    between the ranges (never uploaded) included in the addressing."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    mesh = H.synth.regional_hex_mesh(spacing_m=12000.0, extent_x_m=2400e3, extent_y_m=1600e3, seed=5)
    if order != "rowmajor":
        mesh = H.synth.renumber_cells(mesh, order, seed=2)
    lon, lat = H.latlon_grid(192, 128, lon0=-108.0, lon1=-87.0, lat0=32.0, lat1=45.0)
    rng = np.random.default_rng(3)
    srcs = [rng.standard_normal((mesh.nCells, nl)).astype(np.float32) for nl in (60, 61, 1)]
    moved = {}
    for nranks, rank in ((1, 0), (4, 1)):
        rg = Regridder(device=0, rank=rank, nranks=nranks)
        rg.set_mesh(mesh.lonCell, mesh.latCell, mesh.lonVertex, mesh.latVertex, mesh.verticesOnCell)
        rg.set_target(l.CENTER, lon, lat)
        r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
        n = r.info()["nDst"]
        host_out = [np.full((s.shape[1], n), np.nan, np.float32) for s in srcs]
        b0 = rg.io_bytes()[0]
        rg.apply(r, srcs, host_out, nlev=[s.shape[1] for s in srcs])
        moved[nranks] = rg.io_bytes()[0] - b0
        dev_src = [torch.from_numpy(s).cuda() for s in srcs]
        dev_out = [torch.full((s.shape[1], n), float("nan"), device="cuda") for s in srcs]
        rg.apply(r, dev_src, dev_out, nlev=[s.shape[1] for s in srcs])
        rg.synchronize()
        for h, d in zip(host_out, dev_out):
            assert np.array_equal(h, d.cpu().numpy())
        r.release()
        rg.close()
    full = sum(s.nbytes for s in srcs)
    assert moved[1] <= full
    if order in ("rowmajor", "morton"):       # a quarter of the rows needs about a quarter of the cells (+ halo, + merged gaps)
        assert moved[4] < 0.45 * full, (moved, full)
    else:                                      # a random numbering references cells all over the id space
        assert moved[4] <= full
