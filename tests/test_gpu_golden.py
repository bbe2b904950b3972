"""The CUDA engine, through the C ABI, against the committed golden vectors
(tests/golden/lc_small.npz, written by the independent numpy restatement in
tests/golden/make_golden.py).  Indices / masks / row structure bit-exact; weights 1e-12
(1e-9 conservative); applied fp32 fields within the path's 1e-5 relative tolerance
(nearest-neighbour bit-exact)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lc_small.npz")
RTOL = 1e-5  # BASELINE.json north_star: "<= 1e-5 for fp32 fields"


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def _load(rg, g):
    rg.set_mesh(g["lonCell"], g["latCell"], g["lonVertex"], g["latVertex"], g["verticesOnCell"])
    for s, code in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3)):
        rg.set_target(code, g[f"lon_{s}"], g[f"lat_{s}"])
    rg.set_rotation(g["cosa"], g["sina"])


@pytest.fixture(scope="module")
def rg(engine_lib, g):
    from mpassit_b200.regrid import Regridder

    r = Regridder(device=0)
    _load(r, g)
    yield r
    r.close()


def _ell(mask, col, w):
    k = col.shape[1]
    rp = np.zeros(mask.size + 1, np.int32)
    np.cumsum(np.where(mask, k, 0), out=rp[1:])
    return rp, col[mask].reshape(-1).astype(np.int32), w[mask].reshape(-1)


def _close(got, want, rtol=RTOL):
    scale = max(float(np.abs(want).max()), 1e-30)
    assert np.abs(got.astype(np.float64) - want).max() <= rtol * scale


def test_nearest(rg, g):
    from mpassit_b200 import lib as l

    r = rg.store(l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER)
    rp, c, w = r.export_csr()
    n = g["nearest_idx"].size
    assert np.array_equal(rp, np.arange(n + 1)) and np.array_equal(c, g["nearest_idx"]) and np.array_equal(w, np.ones(n))
    out = np.empty((1, n), np.float32)
    rg.apply(r, [g["src_xland"]], [out])
    assert np.array_equal(out, g["dst_xland"])  # bit-exact
    r.release()


def test_bilinear_and_unmapped_mask(rg, g):
    from mpassit_b200 import lib as l

    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    rp, c, w = r.export_csr()
    m = g["bil_elem"] >= 0
    wrp, wc, ww = _ell(m, g["bil_col"], g["bil_w"])
    assert np.array_equal(rp, wrp) and np.array_equal(c, wc)
    assert np.abs(w - ww).max() <= 1e-12
    assert r.info()["nUnmapped"] == int((~m).sum()) == 70
    out = np.full((5, m.size), np.nan, np.float32)
    rg.apply(r, [g["src_theta"]], [out])
    _close(out, g["dst_theta"])
    assert np.array_equal(out[:, ~m], np.zeros((5, 70), np.float32))  # zeroregion=TOTAL
    r.release()


def test_stagger_routes(rg, g):
    from mpassit_b200 import lib as l

    for s, stag in (("U", l.EDGE1), ("V", l.EDGE2)):
        r = rg.store(l.BILINEAR, l.SRC_GRID_CENTER, stag)
        rp, c, w = r.export_csr()
        wrp, wc, ww = _ell(g[f"quad{s}_elem"] >= 0, g[f"quad{s}_col"], g[f"quad{s}_w"])
        assert np.array_equal(rp, wrp) and np.array_equal(c, wc)
        assert np.abs(w - ww).max() <= 1e-12
        r.release()


def test_conserve(rg, g):
    from mpassit_b200 import lib as l

    r = rg.store(l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER)
    rp, c, w = r.export_csr()

    def drop(rp, c, w):
        keep = w > 1e-12
        return np.repeat(np.arange(rp.size - 1), np.diff(rp))[keep], c[keep], w[keep]

    r1, c1, w1 = drop(rp, c, w)
    r2, c2, w2 = drop(g["cons_rowptr"], g["cons_col"], g["cons_w"])
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2)
    assert np.abs(w1 - w2).max() <= 1e-9
    out = np.empty((1, rp.size - 1), np.float32)
    rg.apply(r, [g["src_snow"]], [out])
    _close(out, g["dst_snow"])
    r.release()


def test_node_bilinear(rg, g):
    from mpassit_b200 import lib as l

    r = rg.store(l.BILINEAR, l.SRC_MESH_NODE, l.CENTER)
    rp, c, w = r.export_csr()
    wrp, wc, ww = _ell(g["node_elem"] >= 0, g["node_col"], g["node_w"])
    assert np.array_equal(rp, wrp) and np.array_equal(c, wc) and np.abs(w - ww).max() <= 1e-12
    out = np.empty((5, g["node_elem"].size), np.float32)
    rg.apply(r, [g["src_vort"]], [out])
    _close(out, g["dst_vort"])
    r.release()


def test_wind_chain(rg, g):
    """mass-point bilinear (fp64) -> rotate_winds_cgrid -> centre->edge stagger (interp.F90:259-325)."""
    import torch

    from mpassit_b200 import lib as l

    n = g["bil_elem"].size
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    um = torch.empty((5, n), dtype=torch.float64, device="cuda")
    vm = torch.empty((5, n), dtype=torch.float64, device="cuda")
    rg.apply(r, [torch.from_numpy(g["src_u"]).cuda(), torch.from_numpy(g["src_v"]).cuda()], [um, vm], nlev=[5, 5])
    rg.rotate_winds(um, vm, 5)
    rg.synchronize()
    np.testing.assert_allclose(um.cpu().numpy(), g["dst_umass"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(vm.cpu().numpy(), g["dst_vmass"], rtol=1e-12, atol=1e-12)
    for s, stag, f in (("U", l.EDGE1, um), ("V", l.EDGE2, vm)):
        rs = rg.store(l.BILINEAR, l.SRC_GRID_CENTER, stag)
        out = torch.empty((5, g[f"lat_{s}"].size), dtype=torch.float32, device="cuda")
        rg.apply(rs, [f], [out], nlev=[5])
        rg.synchronize()
        _close(out.cpu().numpy(), g[f"dst_{s}"])
        rs.release()
    r.release()


@pytest.mark.parametrize("nranks", [2, 3, 5])
def test_row_slabs_of_every_rank_tile_the_full_result(engine_lib, g, nranks):
    """Target-row slab partition (para_range, model_grid.F90:2428): rank r of n computes rows
    [j0, j1) only; the slabs of all ranks, concatenated, equal the single-rank result bit-for-bit
    (all ranks emulated one after another on one device; no collective is involved)."""
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    nj, ni = g["lat_M"].shape
    pieces = {k: [] for k in ("theta", "xland", "snow", "U", "rows")}
    for rank in range(nranks):
        r = Regridder(device=0, rank=rank, nranks=nranks)
        _load(r, g)
        j0, j1 = r.slab(l.CENTER)
        pieces["rows"].append((j0, j1))
        for name, method, src in (("theta", l.BILINEAR, g["src_theta"]), ("xland", l.NEAREST_STOD, g["src_xland"]),
                                  ("snow", l.CONSERVE, g["src_snow"])):
            rt = r.store(method, l.SRC_MESH_ELEMENT, l.CENTER)
            nlev = 1 if src.ndim == 1 else src.shape[1]
            out = np.empty((nlev, (j1 - j0) * ni), np.float32)
            r.apply(rt, [src], [out])
            pieces[name].append(out.reshape(nlev, j1 - j0, ni))
            rt.release()
        r.close()
    rows = pieces["rows"]
    assert rows[0][0] == 0 and rows[-1][1] == nj and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    theta = np.concatenate(pieces["theta"], 1).reshape(5, -1)
    _close(theta, g["dst_theta"])
    assert np.array_equal(np.concatenate(pieces["xland"], 1).reshape(1, -1), g["dst_xland"])
    _close(np.concatenate(pieces["snow"], 1).reshape(1, -1), g["dst_snow"])


@pytest.mark.parametrize("split", ["1", "0"])
@pytest.mark.parametrize("nlev,dt", [(5, "f32"), (60, "f32"), (61, "f32"), (130, "f32"), (60, "f64")])
def test_fused_wind_rotation_equals_apply_then_rotate(rg, g, nlev, dt, split):
    """MPRG_EPI_ROT_U / ROT_V (rotation fused into the store of the apply kernel) is bit-identical to
    mprg_apply followed by mprg_rotate_winds_on, for aligned, unaligned and multi-chunk level counts, with the
    pair in its own launch or sharing one with the other fields."""
    import torch

    from mpassit_b200 import lib as l

    rg.set_option("pipe_split", split)

    tdt = torch.float32 if dt == "f32" else torch.float64
    n = g["bil_elem"].size
    nC = g["lonCell"].size
    gen = torch.Generator(device="cuda")
    gen.manual_seed(nlev)
    u = (10.0 * torch.randn((nC, nlev), generator=gen, device="cuda")).to(torch.float32)
    v = (10.0 * torch.randn((nC, nlev), generator=gen, device="cuda")).to(torch.float32)
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    a_u, a_v, b_u, b_v = (torch.full((nlev, n), float("nan"), dtype=tdt, device="cuda") for _ in range(4))
    rg.apply(r, [u, v], [a_u, a_v], nlev=[nlev, nlev])
    rg.rotate_winds(a_u, a_v, nlev)
    rg.apply(r, [u, v], [b_u, b_v], nlev=[nlev, nlev], epi_op=[l.EPI_ROT_U, l.EPI_ROT_V])
    rg.synchronize()
    assert torch.equal(a_u, b_u) and torch.equal(a_v, b_v)
    assert not torch.isnan(b_u).any() and (b_u[:, torch.from_numpy(g["bil_elem"] < 0).cuda()] == 0).all()
    # a third, unrelated field stacked behind the pair is untouched by the rotation
    c_t, c_u, c_v = (torch.empty((nlev, n), dtype=tdt, device="cuda") for _ in range(3))
    t_src = (u * 0.5 + 3.0).contiguous()
    plain = torch.empty((nlev, n), dtype=tdt, device="cuda")
    rg.apply(r, [t_src], [plain], nlev=[nlev])
    rg.apply(r, [u, v, t_src], [c_u, c_v, c_t], nlev=[nlev] * 3, epi_op=[l.EPI_ROT_U, l.EPI_ROT_V, l.EPI_NONE])
    rg.synchronize()
    assert torch.equal(c_t, plain) and torch.equal(c_u, a_u) and torch.equal(c_v, a_v)
    with pytest.raises(l.MprgError):
        rg.apply(r, [u], [c_u], nlev=[nlev], epi_op=[l.EPI_ROT_U])
    r.release()
    rg.set_option("pipe_split", "1")


def test_more_ranks_than_rows_gives_empty_slabs(engine_lib, g):
    """para_range hands no rows to the last ranks when nranks > nj: every call must accept the empty
    slab (zero destination points) and the non-empty ranks must still tile the result."""
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    nj, ni = g["lat_M"].shape
    nranks = nj + 3
    pieces = []
    for rank in (0, nj - 1, nj, nranks - 1):
        r = Regridder(device=0, rank=rank, nranks=nranks)
        _load(r, g)
        j0, j1 = r.slab(l.CENTER)
        for method, src_loc, stag in ((l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER), (l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER),
                                      (l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER), (l.BILINEAR, l.SRC_MESH_NODE, l.CENTER),
                                      (l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER_HALO), (l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE1),
                                      (l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE2)):
            rt = r.store(method, src_loc, stag)
            info = rt.info()
            s0, s1 = r.slab(stag)
            assert info["nDst"] == (s1 - s0) * r.shape[stag if stag != l.CENTER_HALO else l.CENTER][1]
            if method == l.BILINEAR and src_loc == l.SRC_MESH_ELEMENT and stag == l.CENTER:
                out = np.full((5, max(info["nDst"], 1)), np.nan, np.float32)[:, : info["nDst"]]
                out = np.ascontiguousarray(out)
                r.apply(rt, [g["src_theta"]], [out])
                if info["nDst"]:
                    pieces.append((j0, out.reshape(5, j1 - j0, ni)))
            rt.release()
        r.close()
    assert len(pieces) == 2 and pieces[0][0] == 0 and pieces[1][0] == nj - 1
    want = g["dst_theta"].reshape(5, nj, ni)
    for j0, p in pieces:
        scale = np.abs(want).max()
        assert np.abs(p - want[:, j0:j0 + p.shape[1]]).max() <= RTOL * scale


@pytest.mark.parametrize("nranks", [1, 3])
def test_apply_into_full_fields_fuses_the_gather(engine_lib, g, nranks):
    """mprg_apply_into: every rank stores its rows straight into the FULL field (here one device buffer
    shared by ranks emulated one after another; on real multi-GPU runs the writing rank's buffer mapped with
    CUDA IPC).  After all ranks ran, the full fields equal slab apply + gather, bit for bit, for the column
    kernel (aligned / unaligned levels), the flat kernel (2-D, short columns) and the stagger (planes) kernel."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    nj, ni = g["lat_M"].shape
    nC = g["lonCell"].size
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    srcs = {nl: torch.randn((nC, nl), generator=gen, device="cuda") for nl in (1, 4, 60, 61)}
    um = torch.randn((6, nj * ni), generator=gen, device="cuda")   # mass-point wind on the full CENTER grid

    def one_rank_reference():
        r = Regridder(device=0)
        _load(r, g)
        out = {}
        rt = r.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
        for nl, s in srcs.items():
            d = torch.empty((nl, nj * ni), device="cuda")
            r.apply(rt, [s], [d], nlev=[nl])
            out[nl] = d
        rs = r.store(l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE1)
        d = torch.empty((6, g["lat_U"].size), device="cuda")
        r.apply(rs, [um], [d], nlev=[6])
        out["U"] = d
        r.synchronize()
        r.close()
        return out

    want = one_rank_reference()
    full = {nl: torch.full((nl, nj * ni), float("nan"), device="cuda") for nl in srcs}
    fullU = torch.full((6, g["lat_U"].size), float("nan"), device="cuda")
    for rank in range(nranks):
        r = Regridder(device=0, rank=rank, nranks=nranks)
        _load(r, g)
        rt = r.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
        r.apply_into(rt, [srcs[60], srcs[61], srcs[1], srcs[4]], [full[60], full[61], full[1], full[4]], nlev=[60, 61, 1, 4])
        # the stagger reads this rank's CENTER_HALO rows of the mass-point field
        h0, h1 = r.slab(l.CENTER_HALO)
        halo = um.reshape(6, nj, ni)[:, h0:h1].contiguous()
        rs = r.store(l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE1)
        r.apply_into(rs, [halo], [fullU], nlev=[6])
        r.synchronize()
        r.close()
    for nl in srcs:
        assert torch.equal(full[nl], want[nl]), nl
    assert torch.equal(fullU, want["U"])
