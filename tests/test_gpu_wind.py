"""Composed wind routes (mprg_store_wind / mprg_apply_wind): the reference's wind chain -- regrid the cell-centre
winds to the mass points (interp.F90:256-289), rotate_winds_cgrid (:291-293, :689-749), regrid to EDGE1 / EDGE2
(:295-328) -- as one matrix per staggered grid.  Checked against the product of the chain's own matrices, against the
three-step chain on the device, and against the oracle's R8 chain."""
import numpy as np
import pytest

from mpassit_b200 import check, defaults
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(engine_lib):
    from mpassit_b200 import build, host

    build.build_host()
    host.load()
    return host


@pytest.fixture(scope="module")
def lc_case(host, tmp_path_factory):
    d = tmp_path_factory.mktemp("lcw")
    cfg = host.read_setup_namelist(defaults.write_namelist(str(d / "namelist.input"), nx=71, ny=47, dx=30000.0))
    grids = {k: host.target_coords(cfg, s) for k, s in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3))}
    mesh = H.synth.regional_delaunay_mesh(11000, extent_x_m=2000e3, extent_y_m=1300e3, seed=5, lloyd_iters=3)
    cosa, sina = host.get_rotang(*grids["M"])
    return cfg, grids, mesh, cosa, sina


def _load(rg, grids, mesh, cosa, sina):
    rg.set_mesh(mesh.lonCell, mesh.latCell, mesh.lonVertex, mesh.latVertex, mesh.verticesOnCell)
    for k, s in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3)):
        lat, lon = grids[k]
        rg.set_target(s, lon, lat)
    rg.set_rotation(cosa, sina)
    rg.set_option("wind", "composed")      # the default is "chain" (measured slower on the 3-km case)


def _csr(rp, col, w, ncol):
    import scipy.sparse as sp

    return sp.csr_matrix((w, col, rp), shape=(rp.size - 1, ncol))


def _chain(rg, l, u, v, nlev, dst_dtype):
    """The three steps on the device: (u, v) -> mass points with fused rotation -> EDGE1 / EDGE2."""
    import torch

    W = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    nM = W.info()["nDst"]
    um = torch.empty((nlev, nM), dtype=dst_dtype, device="cuda")
    vm = torch.empty_like(um)
    rg.apply(W, [u, v], [um, vm], nlev=[nlev, nlev], epi_op=[l.EPI_ROT_U, l.EPI_ROT_V])
    out = []
    for stag in (l.EDGE1, l.EDGE2):
        S = rg.store(l.BILINEAR, l.SRC_GRID_CENTER, stag)
        d = torch.empty((nlev, S.info()["nDst"]), dtype=dst_dtype, device="cuda")
        rg.apply(S, [um if stag == l.EDGE1 else vm], [d], nlev=[nlev])
        out.append(d)
        S.release()
    W.release()
    rg.synchronize()
    return out


def test_composed_matrix_is_the_product_of_the_chain(engine_lib, lc_case):
    """A = S diag(rden) W, B = S diag(rden tana) W for EDGE1 (and the v' coefficients for EDGE2), entry by entry."""
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    cfg, grids, mesh, cosa, sina = lc_case
    rg = Regridder(device=0)
    _load(rg, grids, mesh, cosa, sina)
    W = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    Wm = _csr(*W.export_csr(), mesh.nCells)
    nM = Wm.shape[0]
    ca, sa = np.asarray(cosa, np.float64).ravel(), np.asarray(sina, np.float64).ravel()
    tana = sa / ca
    rca, rden = 1.0 / ca, 1.0 / (ca + sa * tana)
    import scipy.sparse as sp

    coef = {l.EDGE1: (rden, rden * tana), l.EDGE2: (-(rca * (sa * rden)), rca * (1.0 - sa * (rden * tana)))}
    for stag in (l.EDGE1, l.EDGE2):
        S = rg.store(l.BILINEAR, l.SRC_GRID_CENTER, stag)
        Sm = _csr(*S.export_csr(), nM)
        Cw = rg.store_wind(stag)
        assert Cw is not None
        rp, col, wa = Cw.export_csr()
        wb = Cw.export_w2()
        assert (np.diff(rp) <= 12).all() and np.diff(rp).max() >= 4
        for k in range(rp.size - 1):               # ascending cell ids inside every row
            assert (np.diff(col[rp[k]:rp[k + 1]]) > 0).all()
        fa, fb = coef[stag]
        wantA = (Sm @ sp.diags(fa) @ Wm).tocsr()
        wantB = (Sm @ sp.diags(fb) @ Wm).tocsr()
        gotA, gotB = _csr(rp, col, wa, mesh.nCells), _csr(rp, col, wb, mesh.nCells)
        assert abs(gotA - wantA).max() < 1e-14 and abs(gotB - wantB).max() < 1e-14
        # unmapped staggered points (outside the centre hull: first / last column of EDGE1, first / last row of EDGE2)
        # stay empty rows
        assert np.array_equal(np.diff(rp) == 0, np.diff(Sm.indptr) == 0) or (np.diff(rp) == 0).sum() >= (np.diff(Sm.indptr) == 0).sum()
        info = Cw.info()
        assert info["tiles"] > 0 and info["tile_columns"] > 0
        Cw.release(); S.release()
    W.release()
    rg.close()


@pytest.mark.parametrize("nlev,src_dt,acc", [(8, "f32", "f32"), (60, "f32", "f32"), (64, "f32", "f32"), (72, "f32", "f32"),
                                             (132, "f32", "f32"), (60, "f32", "f64"), (12, "f64", "f64"), (4, "f32", "f32")])
def test_composed_apply_against_chain_and_oracle(engine_lib, orc, lc_case, nlev, src_dt, acc):
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder
    from oracle import interp_oracle

    cfg, grids, mesh, cosa, sina = lc_case
    rg = Regridder(device=0)
    rg.set_option("accumulate", acc)
    _load(rg, grids, mesh, cosa, sina)
    S = H.synth
    tdt = torch.float32 if src_dt == "f32" else torch.float64
    u_h = (8.0 * S.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=31) - 2200.0).astype(np.float32)
    v_h = (5.0 * S.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=32) - 1400.0).astype(np.float32)
    u = torch.from_numpy(u_h).to("cuda", tdt).contiguous()
    v = torch.from_numpy(v_h).to("cuda", tdt).contiguous()
    odt = torch.float32 if src_dt == "f32" else torch.float64
    chainU, chainV = _chain(rg, l, u, v, nlev, odt)
    want = interp_oracle.wind_chain(mesh, grids, u_h.astype(np.float64), v_h.astype(np.float64), cosa, sina)
    for stag, ch, key in ((l.EDGE1, chainU, "U"), (l.EDGE2, chainV, "V")):
        Cw = rg.store_wind(stag)
        assert Cw is not None
        d = torch.full((nlev, Cw.info()["nDst"]), float("nan"), dtype=odt, device="cuda")
        rg.apply_wind(Cw, u, v, d, nlev)
        rg.synchronize()
        g = d.cpu().numpy()
        assert not np.isnan(g).any()
        check.assert_field_close(g, ch.cpu().numpy(), "composed vs chain " + key)
        check.assert_field_close(g, want[key].reshape(g.shape), "composed vs oracle " + key)
        # unmapped staggered points are exact zeros, the same ones as in the chain
        assert np.array_equal(g == 0, ch.cpu().numpy() == 0)
        if acc == "f64":       # R8 arithmetic end to end: one rounding of the result
            w = want[key].reshape(g.shape)
            tol = 2e-7 if odt == torch.float32 else 1e-12
            assert np.abs(g - w).max() <= tol * np.abs(w).max()
        Cw.release()
    rg.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_composed_rank_slabs_equal_single_rank(engine_lib, lc_case, nranks):
    """Row slabs (ranks emulated one after another) tile the single-rank result bit for bit: the composed rows of a
    slab's first / last edge row draw on the halo mass-point rows of MPRG_CENTER_HALO."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    cfg, grids, mesh, cosa, sina = lc_case
    nlev = 20
    S = H.synth
    u = torch.from_numpy(S.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=41).astype(np.float32)).cuda()
    v = torch.from_numpy(S.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=42).astype(np.float32)).cuda()

    def run(rank, n):
        rg = Regridder(device=0, rank=rank, nranks=n)
        _load(rg, grids, mesh, cosa, sina)
        out = {}
        for stag, key in ((l.EDGE1, "U"), (l.EDGE2, "V")):
            Cw = rg.store_wind(stag)
            assert Cw is not None
            d = torch.full((nlev, Cw.info()["nDst"]), float("nan"), dtype=torch.float32, device="cuda")
            rg.apply_wind(Cw, u, v, d, nlev)
            rg.synchronize()
            out[key] = d.cpu().numpy()
            Cw.release()
        rg.close()
        return out

    single = run(0, 1)
    parts = [run(k, nranks) for k in range(nranks)]
    for key in ("U", "V"):
        ni = grids[key][0].shape[1]
        got = np.concatenate([p[key].reshape(nlev, -1, ni) for p in parts], axis=1).reshape(nlev, -1)
        assert np.array_equal(got, single[key]), key


def test_store_wind_declines_where_the_chain_is_needed(engine_lib, lc_case):
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    cfg, grids, mesh, cosa, sina = lc_case
    rg = Regridder(device=0)
    rg.set_mesh(mesh.lonCell, mesh.latCell, mesh.lonVertex, mesh.latVertex, mesh.verticesOnCell)
    for k, s in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3)):
        lat, lon = grids[k]
        rg.set_target(s, lon, lat)
    assert rg.get_option("wind") == "chain" and rg.store_wind(l.EDGE1) is None      # the default: hosts keep the chain
    rg.set_option("wind", "composed")
    assert rg.store_wind(l.EDGE1) is None                       # no rotation registered
    rg.set_rotation(cosa, sina)
    rg.set_option("wind", "chain")
    assert rg.store_wind(l.EDGE1) is None
    rg.set_option("wind", "composed")
    r1 = rg.store_wind(l.EDGE1)
    assert r1 is not None
    r2 = rg.store_wind(l.EDGE1)                                 # memoised
    assert r2.handle == r1.handle
    with pytest.raises(l.MprgError):                            # a composed route is not an mprg_apply route
        rg.apply(r1, [np.zeros((mesh.nCells, 4), np.float32)], [np.zeros((4, r1.info()["nDst"]), np.float32)], nlev=[4])
    with pytest.raises(l.MprgError):
        rg.store_wind(l.CENTER)
    rg.set_rotation(cosa, -np.asarray(sina))                    # new angles: the memoised composition is dropped
    r3 = rg.store_wind(l.EDGE1)
    assert r3 is not None and not np.array_equal(r3.export_w2(), r1.export_w2())
    for r in (r1, r2, r3):
        r.release()
    rg.set_grid_kind(l.GRID_1PERI_MONOPOLE)                     # periodic target grid: seam / pole rows need the chain
    assert rg.store_wind(l.EDGE1) is None
    rg.close()


def test_interp_data_device_pass_composed_equals_chain(engine_lib, host):
    """The bench's code path on the 12-km miniature: the pass with composed wind routes against the same pass with
    option wind = chain -- U and V per element, every other field bit for bit; the composed pass runs no grid-source
    (planes) launch."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder

    wl = workload.make("mid")
    outs, kinds = {}, {}
    for mode in ("composed", "chain"):
        rg = Regridder(device=0)
        rg.set_option("wind", mode)
        workload.load_geometry(rg, wl)
        F = workload.make_fields(wl, device="cuda:0")
        workload.run_interp(rg, wl, F["dev"], l.DEVICE)
        rg.synchronize()
        rg.profile(True)
        workload.run_interp(rg, wl, F["dev"], l.DEVICE)
        rg.synchronize()
        kinds[mode] = sorted(r["kind"] for r in rg.profile_read())
        rg.profile(False)
        got = {s.name: s.dst.cpu().numpy() for g in ("diag", "hist_2d", "hist_3d", "soil") for s in F["dev"][g]
               if not s.name.startswith("uReconstruct")}
        got["HGT"], got["U"], got["V"] = (F["dev"][k].cpu().numpy() for k in ("hgt", "u_stag", "v_stag"))
        outs[mode] = got
        rg.close()
    for nm, w in outs["chain"].items():
        if nm in ("U", "V"):
            check.assert_field_close(outs["composed"][nm], w, nm)
        else:
            assert np.array_equal(outs["composed"][nm], w), nm
    assert 3 in kinds["chain"] and 4 not in kinds["chain"]          # kind 3: grid-source apply, 4: composed wind route
    assert 4 in kinds["composed"] and 3 not in kinds["composed"]
