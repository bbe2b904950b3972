"""N > 1 on the CPU: two gloo ranks run the target-row-slab decomposition (para_range,
model_grid.F90:2428; regDecomp=(/1,npets/), model_grid.F90:693) with the oracle standing in for
the kernels, then gather slabs to rank 0 with the run placement the engine's mprg_gather uses.
Checks that (a) every destination row is owned exactly once, (b) a slab computed alone equals
the same rows of the single-rank result bit-for-bit (rows of W are independent: no data-path
collective is needed), (c) the gathered [lev][nj][ni] field equals the single-rank field."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from mpassit_b200 import host
    from oracle import oracle as orc
    from tests import helpers as H

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc.set_num_threads(1)
        g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lc_small.npz")))
        lo, la = orc.mesh_rad_to_deg(g["lonCell"], g["latCell"])
        cxyz = orc.sph_deg_to_cart(lo, la)
        nlev = g["src_theta"].shape[1]
        results = {}
        for stag in ("M", "V"):  # CENTER has nj rows, EDGE2 nj + 1: each stagger partitions its own row count
            lat, lon = g[f"lat_{stag}"], g[f"lon_{stag}"]
            nj, ni = lat.shape
            j0, j1 = host.slab_rows(nj, world, rank)
            d = orc.sph_deg_to_cart(lon[j0:j1], lat[j0:j1])
            e, c, w = orc.bilinear(cxyz, g["tri"], g["verticesOnCell"], d)
            slab = orc.apply(*orc.ell_to_csr(e >= 0, c, w), g["src_theta"], np.float32)  # [nlev][(j1-j0)*ni]
            runs = host.gather_runs(ni, nj, nlev, world, rank)
            if rank == 0:
                full = np.full(nlev * nj * ni, np.nan, np.float32)
                for so, fo, n in runs:
                    full[fo:fo + n] = slab.reshape(-1)[so:so + n]
                for p in range(1, world):
                    for so, fo, n in host.gather_runs(ni, nj, nlev, world, p):
                        buf = torch.empty(n, dtype=torch.float32)
                        dist.recv(buf, src=p)
                        full[fo:fo + n] = buf.numpy()
                results[stag] = (full.reshape(nlev, nj * ni), (j0, j1))
            else:
                flat = torch.from_numpy(slab.reshape(-1).copy())
                for so, fo, n in runs:
                    dist.send(flat[so:so + n].contiguous(), dst=0)
        dist.barrier()
        if rank == 0:
            q.put(results)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_and_gather(orc, world):
    import torch.multiprocessing as mp

    from mpassit_b200 import build, host

    build.build_host()
    host.load()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lc_small.npz")))
    # single-rank result of the same field
    lo, la = orc.mesh_rad_to_deg(g["lonCell"], g["latCell"])
    cxyz = orc.sph_deg_to_cart(lo, la)
    for stag in ("M", "V"):
        e, c, w = orc.bilinear(cxyz, g["tri"], g["verticesOnCell"], orc.sph_deg_to_cart(g[f"lon_{stag}"], g[f"lat_{stag}"]))
        want = orc.apply(*orc.ell_to_csr(e >= 0, c, w), g["src_theta"], np.float32)
        full, _ = res[stag]
        assert not np.isnan(full).any()          # every row owned by some rank
        assert np.array_equal(full, want)        # bit-for-bit: destination rows are independent
    np.testing.assert_allclose(res["M"][0], g["dst_theta"], rtol=1.2e-7)


def test_gather_runs_cover_the_field_exactly_once(engine_lib):
    from mpassit_b200 import build, host

    build.build_host()
    host.load()
    for ni, nj, nlev, world in ((1800, 1060, 3, 8), (1800, 1061, 2, 8), (5, 3, 4, 8), (7, 1, 1, 2)):
        cover = np.zeros(nlev * nj * ni, np.int32)
        total = 0
        for r in range(world):
            for so, fo, n in host.gather_runs(ni, nj, nlev, world, r):
                cover[fo:fo + n] += 1
                total += n
        assert (cover == 1).all() and total == cover.size


def _file_worker(rank, world, port, rundir, q):
    import ctypes as C

    import torch.distributed as dist

    from mpassit_b200 import host

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        host.load()
        cb = host.torch_comm()
        # the reductions mpassit_run asks of its host for P_TOP (write_data.F90:1364-1373)
        v = (C.c_double * 2)(10.0 + rank, -3.0 * rank)
        cb(None, host.COMM_MAX, v, 2)
        mx = list(v)
        v = (C.c_double * 1)(5.0 - rank)
        cb(None, host.COMM_MIN, v, 1)
        # header-only run: rank 0 creates the file, the others wait at the barrier and open the same layout
        st = host.run(os.path.join(rundir, "namelist.files"), rundir, device=-1, rank=rank, nranks=world, comm=cb)
        dist.barrier()
        q.put((rank, mx, v[0], st.output_version, os.path.getsize(os.path.join(rundir, "mpassit_out.nc"))))
    finally:
        dist.destroy_process_group()


def test_file_driver_comm_callback_and_shared_output_file(engine_lib, tmp_path):
    import torch.multiprocessing as mp

    from mpassit_b200 import build, host, workload
    from tests import mpas_files

    build.build_host()
    host.load()
    wl = workload.make("mini", rundir=str(tmp_path))
    F = {g: [(nm, workload._field_values_torch(wl, g, nm, wl.levels_of(g, nm), k, "cpu").numpy())
             for k, (nm, _) in enumerate(wl.lists[g])] for g in ("diag", "hist_2d", "hist_3d", "soil")}
    mpas_files.write_case(wl, str(tmp_path), F, np.zeros(wl.mesh.lonCell.size, np.float32))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_file_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mx, mn, ver, size in res:
        assert mx == [11.0, 0.0] and mn == 4.0 and ver == 2
        assert size == res[0][4] > 0
    out, g, va, dims, order = mpas_files.read_output(str(tmp_path / "mpassit_out.nc"))
    assert dims["west_east"] == wl.cfg.i_target and "PB" in order
