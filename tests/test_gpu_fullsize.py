"""BASELINE.json configs at FULL size.

configs[1] (c2: 3-km regional mesh, 2.4 M cells -> Lambert 1801 x 1061, the configuration the metric is quoted on):
  * the whole interp_data pass of the bench, every output field against the CPU oracle, element by element
    (test_c2_full_pass_every_field_against_the_oracle; the oracle needs ~10 s for it on the GPU box's host cores);
  * size-independent properties -- partition of unity, reproduction of constants, linearity, nearest-neighbour
    output drawn bit-exactly from the source, conservative weights bounded by 1, zero fill, rank slabs.
configs[2] (c3: 914 level-columns in ONE stacked apply), configs[3] (c4: 6.5 M-cell variable-resolution global mesh
-> 0.03 degree), configs[4] (c5: 1-km conservative, weights rebuilt per run): against the oracle on row samples / in full."""
import numpy as np
import pytest

from mpassit_b200 import check

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(engine_lib):
    import torch

    from mpassit_b200 import build, workload
    from mpassit_b200.regrid import Regridder

    build.build_host()
    wl = workload.make("c2")
    rg = Regridder(device=0)
    workload.load_geometry(rg, wl)
    yield wl, rg, torch
    rg.close()


def test_c2_full_pass_every_field_against_the_oracle(c2, orc):
    """The bench's own pass (workload.run_interp, device buffers, stock var-lists, fused wind rotation, stagger) at
    the full 1801 x 1061 size: 44 output fields, 2.1e9 values, each compared with the oracle's interp_data --
    nearest-neighbour classes bit-exact, everything else per element within 1e-5 relative."""
    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from oracle import interp_oracle

    wl, rg, torch = c2
    F = workload.make_fields(wl, device="cuda:0")
    workload.run_interp(rg, wl, F["dev"], l.DEVICE)
    rg.synchronize()
    fields = {g: [(s.name, s.src.cpu().numpy()) for s in F["dev"][g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
    fields["ter"] = F["dev"]["ter"].cpu().numpy()
    want = interp_oracle.interp_data(wl.mesh, wl.grids, fields, wl.cosa, wl.sina, lc=True)
    got = {s.name: s.dst for g in ("diag", "hist_2d", "hist_3d", "soil") for s in F["dev"][g]}
    got["HGT"], got["U"], got["V"] = (F["dev"][k] for k in ("hgt", "u_stag", "v_stag"))
    exact = {"xland", "tslb", "smois", "sh2o"}
    res = {}
    for nm in list(want):
        w = want.pop(nm)
        if nm.startswith("uReconstruct"):
            continue
        g = got[nm].cpu().numpy().reshape(w.shape)
        if nm in exact:
            check.assert_field_exact(g, w, nm)
        res[nm] = check.assert_field_close(g, w, nm)
    assert len(res) >= 44
    summ = check.summarize(res)
    assert summ["ok_fields"] == summ["fields"] and summ["exact_fields"] >= 4 and summ["max_rel"] <= 1e-5
    # exact zeros of the oracle (moisture / snow fields are ~70 % zeros) are exact zeros here
    assert all(r["zeros_kept"] for r in res.values())
    del F, fields, got


def test_c2_device_generated_target_keeps_the_matrix_structure(c2):
    """SURVEY.md 8 f3 at the full 1801 x 1061 size: coordinates of all four staggers generated on the device agree
    with the host mirror to a few ulp, and the bilinear / nearest / conservative / stagger matrices built from them
    have the same mapped mask and the same indices as those built from the host's coordinates."""
    from mpassit_b200 import lib as l
    from tests.test_gpu_targetgen import compare_target_generation

    wl, rg, torch = c2
    routes = [(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER), (l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER),
              (l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER), (l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE1),
              (l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE2)]
    res = compare_target_generation(wl, routes)
    print("c2", res)
    assert res["ulp_lon"] <= 8 and res["ulp_lat"] <= 8 and res["structure_diffs"] <= 1e-5 * res["rows"], res


def test_bilinear_weights_partition_of_unity_and_constant_field(c2):
    from mpassit_b200 import lib as l

    wl, rg, torch = c2
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    info = r.info()
    assert info["nDst"] == 1800 * 1060 and info["nnz"] == 3 * (info["nDst"] - info["nUnmapped"])
    rp, col, w = r.export_csr()
    rows = np.diff(rp)
    assert set(np.unique(rows)) <= {0, 3}
    ws = w.reshape(-1, 3)
    assert np.abs(ws.sum(1) - 1).max() <= 1e-12 and ws.min() >= -1e-10 and ws.max() <= 1 + 1e-10
    assert col.min() >= 0 and col.max() < wl.mesh.nCells
    # constant in -> the same constant out on every mapped point (fp32 accumulation of a convex combination)
    n = wl.mesh.nCells
    src = torch.full((n, 60), 273.15, dtype=torch.float32, device="cuda")
    dst = torch.empty((60, info["nDst"]), dtype=torch.float32, device="cuda")
    rg.apply(r, [src], [dst], nlev=[60])
    rg.synchronize()
    mapped = torch.from_numpy(rows == 3).cuda()
    assert torch.all(torch.abs(dst[:, mapped] - 273.15) <= 273.15 * 2.4e-7)
    assert torch.all(dst[:, ~mapped] == 0)
    r.release()


def test_linearity_and_unaligned_levels(c2):
    from mpassit_b200 import lib as l

    wl, rg, torch = c2
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    n, nd = wl.mesh.nCells, 1800 * 1060
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    for nlev in (60, 61):
        f = torch.randn((n, nlev), generator=g, device="cuda")
        h = torch.randn((n, nlev), generator=g, device="cuda")
        comb = (2.0 * f - 0.5 * h).contiguous()
        of, oh, oc = (torch.empty((nlev, nd), dtype=torch.float32, device="cuda") for _ in range(3))
        rg.apply(r, [f, h, comb], [of, oh, oc], nlev=[nlev] * 3)
        rg.synchronize()
        err = (oc - (2.0 * of - 0.5 * oh)).abs().max().item()
        assert err <= 1e-5 * max(oc.abs().max().item(), 1.0), (nlev, err)
    r.release()


def test_nearest_output_is_drawn_bit_exactly_from_the_source(c2):
    from mpassit_b200 import lib as l

    wl, rg, torch = c2
    r = rg.store(l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER)
    rp, col, w = r.export_csr()
    assert np.array_equal(rp, np.arange(rp.size)) and np.all(w == 1.0)
    n = wl.mesh.nCells
    src = torch.arange(n * 4, dtype=torch.float32, device="cuda").reshape(n, 4) % 8191.0
    dst = torch.empty((4, col.size), dtype=torch.float32, device="cuda")
    rg.apply(r, [src], [dst], nlev=[4])
    rg.synchronize()
    want = src[torch.from_numpy(col.astype(np.int64)).cuda()].T.contiguous()
    assert torch.equal(dst, want)
    # idempotence of the search: the nearest cell of a cell centre is that cell
    r.release()


def test_conservative_weights_are_area_fractions(c2):
    from mpassit_b200 import lib as l

    wl, rg, torch = c2
    r = rg.store(l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER)
    rp, col, w = r.export_csr()
    assert w.min() > 0 and w.max() <= 1 + 1e-9
    frac = np.add.reduceat(w, rp[:-1][np.diff(rp) > 0])
    # the hex mesh covers the whole Lambert target: every destination cell is fully covered
    assert np.all(np.diff(rp) > 0) and np.abs(frac - 1).max() <= 1e-9
    # rows are sorted by ascending source id (ESMF's factorIndexList order)
    inner = np.ones(col.size, bool)
    inner[rp[1:-1]] = False
    assert np.all(np.diff(col)[inner[1:]] > 0)
    r.release()


def test_rank_slab_equals_rows_of_the_single_rank_result(c2):
    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder

    wl, rg, torch = c2
    n = wl.mesh.nCells
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    src = torch.randn((n, 60), generator=g, device="cuda")
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    full = torch.empty((60, 1800 * 1060), dtype=torch.float32, device="cuda")
    rg.apply(r, [src], [full], nlev=[60])
    rg.synchronize()
    r.release()
    r8 = Regridder(device=0, rank=5, nranks=8)
    workload.load_geometry(r8, wl)
    j0, j1 = r8.slab(l.CENTER)
    rt = r8.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    slab = torch.empty((60, (j1 - j0) * 1800), dtype=torch.float32, device="cuda")
    r8.apply(rt, [src], [slab], nlev=[60])
    r8.synchronize()
    assert torch.equal(slab.reshape(60, j1 - j0, 1800), full.reshape(60, 1060, 1800)[:, j0:j1])
    rt.release()
    r8.close()


def test_large_target_needs_64_bit_offsets(engine_lib):
    """BASELINE.json configs[3] in spirit: a fine global lat-lon target (6000 x 3000 = 18 M points) from a coarse
    global mesh; one 61-level field is 4.4 GB of output, so every destination offset beyond 2^32 bytes is
    exercised.  Checked through reproduction of a constant and of nearest-neighbour indices at the far end."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder
    from tests import helpers as H

    mesh = H.small_global(2562, jitter=0.1, seed=4)
    ni, nj = 6000, 3000
    lon = (-180.0 + (np.arange(ni) + 0.5) * 360.0 / ni)[None, :].repeat(nj, 0)
    lat = (-90.0 + (np.arange(nj) + 0.5) * 180.0 / nj)[:, None].repeat(ni, 1)
    rg = Regridder(device=0)
    rg.set_mesh(mesh.lonCell, mesh.latCell, mesh.lonVertex, mesh.latVertex, mesh.verticesOnCell)
    rg.set_target(l.CENTER, np.ascontiguousarray(lon), np.ascontiguousarray(lat))
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    info = r.info()
    assert info["nDst"] == ni * nj and info["nUnmapped"] == 0 and info["nnz"] == 3 * ni * nj
    nlev = 61
    src = torch.full((mesh.nCells, nlev), 250.0, device="cuda") + torch.arange(nlev, device="cuda")[None, :]
    dst = torch.empty((nlev, ni * nj), device="cuda")
    assert dst.numel() * 4 > 2 ** 32
    rg.apply(r, [src], [dst], nlev=[nlev])
    rg.synchronize()
    want = 250.0 + torch.arange(nlev, device="cuda", dtype=torch.float32)
    err = (dst - want[:, None]).abs().amax(dim=1)
    assert float((err / want).max()) <= 2.4e-7
    r.release()
    del dst
    rn = rg.store(l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER)
    ids = torch.arange(mesh.nCells, device="cuda", dtype=torch.float32).reshape(-1, 1).repeat(1, 4).contiguous()
    out = torch.empty((4, ni * nj), device="cuda")
    rg.apply(rn, [ids], [out], nlev=[4])
    rg.synchronize()
    rp, col, w = rn.export_csr()
    assert torch.equal(out[3], torch.from_numpy(col.astype(np.float32)).cuda())
    rn.release()
    rg.close()


def test_c3_914_level_columns_in_one_stacked_apply(c2, orc):
    """BASELINE.json configs[2]: every histlist_3d + histlist_soil field of the 3-km case -- 11 x 60 + 2 x 61 +
    2 x 60 + 3 x 4 = 914 level-columns, 1.74e9 point-levels -- through ONE mprg_apply on the bilinear route.
    Every 53rd target row (20 rows, 36,000 points x 914 columns) is compared with the oracle per element; over the
    whole grid the stack must equal the same fields applied one call at a time, bit for bit."""
    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from tests import helpers as H

    wl, rg, torch = c2
    names, srcs, dsts = workload.make_stacked(wl, "cuda:0")
    levs = [int(t.shape[1]) for t in srcs]
    assert sum(levs) == 914
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    n0 = rg.kernel_launches
    rg.apply(r, srcs, dsts, nlev=levs)
    rg.synchronize()
    assert 1 <= rg.kernel_launches - n0 <= 3          # one stacked apply = a handful of launches, not one per field
    lat, lon = wl.grids["M"]
    nj, ni = lat.shape
    rows = np.arange(7, nj, 53)
    pick = torch.from_numpy((rows[:, None] * ni + np.arange(ni)[None, :]).reshape(-1)).cuda()
    cxyz, _, tri = H.oracle_geometry(orc, wl.mesh)
    e, c, w = orc.bilinear(cxyz, tri, wl.mesh.verticesOnCell, orc.sph_deg_to_cart(lon[rows], lat[rows]))
    csr = orc.ell_to_csr(e >= 0, c, w)
    for nm, sr, ds in zip(names, srcs, dsts):
        want = orc.apply(*csr, sr.cpu().numpy(), np.float32)
        check.assert_field_close(ds[:, pick].cpu().numpy(), want, nm)
    # the same fields one call at a time
    for k in (0, 1, 3, len(names) - 1):
        one = torch.empty_like(dsts[k])
        rg.apply(r, [srcs[k]], [one], nlev=[levs[k]])
        rg.synchronize()
        assert torch.equal(one, dsts[k]), names[k]
    r.release()


def test_c4_variable_resolution_global_6p5M_cells_to_0p03_degree(engine_lib, orc):
    """BASELINE.json configs[3] at its stated size: 15-3 km variable-resolution global mesh, 6,496,362 cells
    (geodesic frequency 806, graded 5:1 towards CONUS) -> the 0.03-degree global lat-lon target, 12000 x 6000 =
    72 M points; one bilinear 55-level field (15.8 GB of output) and one nearest-neighbour integer field.
    Against the oracle on every 47th target row (128 rows, 1.5 M points): indices bit-exact, nearest output
    bit-exact, bilinear per element within 1e-5; over the whole target: nothing unmapped, 3 weights per point,
    convex-combination bounds, every nearest value drawn from the source."""
    import torch

    from mpassit_b200 import build, workload
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder
    from tests import helpers as H

    build.build_host()
    wl = workload.make("c4")
    mesh = wl.mesh
    assert mesh.nCells == 6496362
    lat, lon = wl.grids["M"]
    nj, ni = lat.shape
    assert (ni, nj) == (12000, 6000)
    nlev = wl.nz
    rg = Regridder(device=0)
    workload.load_geometry(rg, wl)
    rows = np.arange(23, nj, 47)
    pick = torch.from_numpy((rows[:, None] * ni + np.arange(ni)[None, :]).reshape(-1)).cuda()
    cxyz, _, tri = H.oracle_geometry(orc, mesh)
    dxyz = orc.sph_deg_to_cart(lon[rows], lat[rows])

    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    src = (280.0 + 20.0 * torch.randn((mesh.nCells, nlev), generator=g, device="cuda")).contiguous()
    r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
    info = r.info()
    assert info["nDst"] == ni * nj and info["nUnmapped"] == 0 and info["nnz"] == 3 * ni * nj
    dst = torch.empty((nlev, ni * nj), device="cuda")
    rg.apply(r, [src], [dst], nlev=[nlev])
    rg.synchronize()
    got = dst[:, pick].cpu().numpy()
    lo_, hi_ = src.amin(dim=0), src.amax(dim=0)   # a convex combination stays inside the source range, everywhere
    assert torch.all(dst.amin(dim=1) >= lo_ - 1e-3) and torch.all(dst.amax(dim=1) <= hi_ + 1e-3)
    del dst
    e, c, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    assert (e >= 0).all()
    # indices and weights of the sampled rows: bit-exact structure, 1e-12 weights
    rp, cc, ww = r.export_csr()
    sel = (rows[:, None] * ni + np.arange(ni)[None, :]).reshape(-1).astype(np.int64)
    assert np.array_equal(cc.reshape(-1, 3)[sel], c) and np.abs(ww.reshape(-1, 3)[sel] - w).max() <= 1e-12
    del rp, cc, ww
    want = orc.apply(*orc.ell_to_csr(e >= 0, c, w), src.cpu().numpy(), np.float32)
    check.assert_field_close(got, want, "c4 bilinear")
    r.release()

    rn = rg.store(l.NEAREST_STOD, l.SRC_MESH_ELEMENT, l.CENTER)
    ids = torch.arange(mesh.nCells, device="cuda", dtype=torch.float32).reshape(-1, 1).contiguous()  # exact below 2^24
    veg = (torch.arange(mesh.nCells, device="cuda") * 2654435761 % 20 + 1).to(torch.float32).reshape(-1, 1).contiguous()
    oi, ov = torch.empty((1, ni * nj), device="cuda"), torch.empty((1, ni * nj), device="cuda")
    rg.apply(rn, [ids, veg], [oi, ov], nlev=[1, 1])
    rg.synchronize()
    near = orc.nearest(cxyz, dxyz)
    assert np.array_equal(oi[0, pick].cpu().numpy().astype(np.int64), near.astype(np.int64))
    assert torch.equal(ov[0], veg[:, 0][oi[0].long()])          # every one of the 72 M values is drawn from the source
    rn.release()
    rg.close()


def test_c5_conservative_1km_weights_rebuilt_every_run(engine_lib, orc):
    """BASELINE.json configs[4] at full size: 1-km regional mesh (1.4 M cells) -> 1-km Lambert 1001 x 1001
    (1 M destination cells), first-order conservative `snow` / `snowh`, the weights rebuilt on every run.
    The oracle finishes this one in seconds, so parity is direct: which sources each destination cell
    overlaps (bit-exact, ascending ids), the overlap fractions (1e-12), and both fields (1e-5) -- and the
    rebuilt routes are identical run after run."""
    import torch

    from mpassit_b200 import build, workload
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder
    from tests import helpers as H

    build.build_host()
    wl = workload.make("c5")
    mesh = wl.mesh
    rg = Regridder(device=0)
    workload.load_geometry(rg, wl)
    cxyz, vxyz, _ = H.oracle_geometry(orc, mesh)
    clat, clon = wl.grids["CORNER"]
    cor = orc.sph_deg_to_cart(clon, clat).reshape(clat.shape[0], clat.shape[1], 3)
    rp, cc, ww = orc.conserve(cxyz, vxyz, mesh.verticesOnCell, cor)
    assert rp.size - 1 == 1000 * 1000
    snow = H.synth.patchy_field(mesh.lonCell, mesh.latCell)
    snowh = (0.01 * snow).astype(np.float32)
    want = [orc.apply(rp, cc, ww, f, np.float32) for f in (snow, snowh)]
    ds, dh = torch.from_numpy(snow).cuda(), torch.from_numpy(snowh).cuda()
    for run in range(3):
        rg.clear_routes()                      # a new run: nothing memoised
        r = rg.store(l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER)
        o1, o2 = torch.full((1, 1000 * 1000), float("nan"), device="cuda"), torch.full((1, 1000 * 1000), float("nan"), device="cuda")
        rg.apply(r, [ds, dh], [o1, o2], nlev=[1, 1])
        rg.synchronize()
        grp, gc, gw = r.export_csr()
        assert np.array_equal(grp, rp) and np.array_equal(gc, cc), run
        assert np.abs(gw - ww).max() <= 1e-12, run
        for got, w in zip((o1, o2), want):
            g = got.cpu().numpy()
            check.assert_field_close(g, w, f"c5 run {run}")
    # the mesh overhangs the target on every side: each destination cell is fully covered
    # (1-km cells are 1e-8 of the unit sphere: area differences cancel to ~4e-9 relative, in the oracle alike)
    frac = np.add.reduceat(gw, grp[:-1])
    assert np.abs(frac - 1).max() <= 2e-8
    rg.close()
