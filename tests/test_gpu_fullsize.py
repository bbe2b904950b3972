"""BASELINE.json configs[1] at FULL size (3-km regional mesh, 2.4 M cells -> Lambert 1801 x 1061):
too large for the CPU oracle inside a test, so parity is checked through size-independent properties
of the path -- partition of unity, exact reproduction of constants, linearity, nearest-neighbour
output drawn bit-exactly from the source, conservative weights bounded by 1, zero fill, and
rank-slab results equal to the same rows of the single-rank result."""
import numpy as np
import pytest

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
