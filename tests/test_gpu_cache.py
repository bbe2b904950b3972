"""Cross-run weight cache (include/mpassit_rg.h: mprg_set_weight_cache): a second run of the same
case loads every matrix instead of generating it, results are byte-identical with and without the cache, and
anything that changes a matrix (geometry, slab, topology) changes its key."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pass(wl, cache_dir, rank=0, nranks=1, mutate=None):
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder

    rg = Regridder(device=0, rank=rank, nranks=nranks)
    rg.set_weight_cache(cache_dir)
    if mutate:
        mutate(wl)
    workload.load_geometry(rg, wl)
    F = workload.make_fields(wl, device="cuda:0", rg=rg)
    workload.run_interp(rg, wl, F["dev"], l.DEVICE)
    rg.synchronize()
    out = {s.name: s.dst.cpu().numpy().copy() for g in ("diag", "hist_2d", "hist_3d", "soil") for s in F["dev"][g]
           if not s.name.startswith("uReconstruct")}
    for k in ("hgt", "u_stag", "v_stag"):
        out[k] = F["dev"][k].cpu().numpy().copy()
    csr = {}
    for tag, key in (("bil", (l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)), ("cons", (l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER)),
                     ("u", (l.BILINEAR, l.SRC_GRID_CENTER, l.EDGE1))):
        r = rg.store(*key)
        csr[tag] = r.export_csr()
        r.release()
    stats = rg.weight_cache_stats()
    rg.close()
    torch.cuda.synchronize()
    return out, csr, stats


def test_second_run_loads_every_matrix_and_gives_the_same_bytes(engine_lib, tmp_path):
    from mpassit_b200 import build, workload

    build.build_host()
    os.makedirs(tmp_path / "run")
    wl = workload.make("mini", rundir=str(tmp_path / "run"))
    cache = str(tmp_path / "wcache")
    plain, csr0, st0 = _pass(wl, None)
    assert st0 == (0, 0)
    first, csr1, st1 = _pass(wl, cache)
    assert st1[0] == 0 and st1[1] >= 5            # bilinear, nearest, conserve, stagger u / v: generated and written
    files = sorted(glob.glob(os.path.join(cache, "mprg_*.w")))
    assert len(files) == st1[1]
    second, csr2, st2 = _pass(wl, cache)
    assert st2 == (st1[1], 0)                      # every matrix loaded, none generated
    for nm in plain:
        assert np.array_equal(plain[nm], first[nm]) and np.array_equal(plain[nm], second[nm]), nm
    for tag in csr0:
        for a, b, c in zip(csr0[tag], csr1[tag], csr2[tag]):
            assert np.array_equal(a, b) and np.array_equal(a, c), tag
    # a damaged file is ignored and rewritten
    with open(files[0], "r+b") as fh:
        fh.truncate(os.path.getsize(files[0]) // 2)
    third, _, st3 = _pass(wl, cache)
    assert st3 == (st1[1] - 1, 1)
    assert all(np.array_equal(plain[nm], third[nm]) for nm in plain)


def test_keys_follow_geometry_slab_and_topology(engine_lib, tmp_path):
    from mpassit_b200 import build, workload

    build.build_host()
    os.makedirs(tmp_path / "run")
    os.makedirs(tmp_path / "run2")
    wl = workload.make("mini", rundir=str(tmp_path / "run"))
    cache = str(tmp_path / "wcache")
    _, _, st = _pass(wl, cache)
    n0 = st[1]

    def nudge(w):
        w.mesh.latCell[w.mesh.nCells // 2] += 1e-9            # one cell centre moves by 6 mm

    _, _, st_moved = _pass(wl, cache, mutate=nudge)
    assert st_moved[0] < n0 and st_moved[1] >= 1               # mesh-source routes miss; the grid-to-grid ones still hit
    wl2 = workload.make("mini", rundir=str(tmp_path / "run2"))
    _, _, st_rank = _pass(wl2, cache, rank=1, nranks=2)        # another slab: other rows, other keys
    assert st_rank[0] == 0 and st_rank[1] >= 5
    _, _, st_rank_again = _pass(wl2, cache, rank=1, nranks=2)
    assert st_rank_again == (st_rank[1], 0)
