"""Two real GPUs (skipped on a one-GPU box): every rank regrids its own target-row slab, the slabs
are gathered to rank 0 with mprg_gather (grouped ncclSend/ncclRecv over NVLink: the path's only
collective, replacing ESMF_FieldGather, write_data.F90:1006-1453) and must equal the single-rank
result bit-for-bit, for a CENTER field, the staggered winds (EDGE1 / EDGE2) and a nearest field."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = workload.make("mini")
        rg = Regridder(device=rank, rank=rank, nranks=world)
        workload.load_geometry(rg, wl)
        ids = [rg.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        rg.comm_init(ids[0])
        F = workload.make_fields(wl, device=f"cuda:{rank}", rg=rg)
        workload.run_interp(rg, wl, F["dev"], l.DEVICE)
        names = {"theta": ("hist_3d", l.CENTER), "tslb": ("soil", l.CENTER)}
        got = {}
        for nm, (grp, stag) in names.items():
            s = next(x for x in F["dev"][grp] if x.name == nm)
            full = torch.full((s.nlev, wl.n_mass), float("nan"), device="cuda") if rank == 0 else None
            rg.gather(stag, s.nlev, s.dst, 0, full)
            got[nm] = full
        for nm, stag, key in (("U", l.EDGE1, "u_stag"), ("V", l.EDGE2, "v_stag")):
            n = wl.grids[nm][0].size
            full = torch.full((wl.nz, n), float("nan"), device="cuda") if rank == 0 else None
            rg.gather(stag, wl.nz, F["dev"][key], 0, full)
            got[nm] = full
        rg.synchronize()
        dist.barrier()
        # ---- the same pass with the gather fused into the store: rank 0 owns full-grid fields, rank 1 maps
        #      them with CUDA IPC and its apply kernels write their rows straight into rank 0's memory
        full = workload.full_outputs(wl, "cuda") if rank == 0 else None
        if rank == 0:
            for v in full.values():
                for t in (v if isinstance(v, list) else [v]):
                    t.fill_(float("nan"))
            torch.cuda.synchronize()
        box = [workload.export_full(rg, full) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        dst = full if rank == 0 else workload.open_full(rg, box[0])
        workload.run_interp(rg, wl, workload.with_destinations(F["dev"], dst), l.DEVICE, dst_full=True)
        rg.synchronize()
        dist.barrier()
        if rank != 0:
            rg.ipc_close_all()
        if rank == 0:
            # the same pass on one rank, same device, same synthetic inputs (seeded per field)
            r1 = Regridder(device=0)
            workload.load_geometry(r1, wl)
            F1 = workload.make_fields(wl, device="cuda:0")
            workload.run_interp(r1, wl, F1["dev"], l.DEVICE)
            r1.synchronize()
            want = {nm: next(x for x in F1["dev"][grp] if x.name == nm).dst for nm, (grp, _) in names.items()}
            want["U"], want["V"] = F1["dev"]["u_stag"], F1["dev"]["v_stag"]
            res = {nm: bool(torch.equal(got[nm], want[nm])) and not bool(torch.isnan(got[nm]).any()) for nm in want}
            # fused gather: every field of the pass
            for grp in ("diag", "hist_2d", "hist_3d", "soil"):
                for s1, t in zip(F1["dev"][grp], full[grp]):
                    if s1.name.startswith("uReconstruct"):
                        continue
                    res["fused:" + s1.name] = bool(torch.equal(t, s1.dst))
            for k in ("hgt", "u_stag", "v_stag"):
                res["fused:" + k] = bool(torch.equal(full[k], F1["dev"][k]))
            q.put(res)
            r1.close()
        dist.barrier()
        rg.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_slabs_gather_equals_single_rank(engine_lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from mpassit_b200 import build

    build.build_host()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res and all(res.values()), res


def _file_worker(rank, world, port, rundir, q):
    import torch
    import torch.distributed as dist

    from mpassit_b200 import host

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        host.load()
        st = host.run(os.path.join(rundir, "namelist.files"), rundir, device=rank, rank=rank, nranks=world,
                      comm=host.torch_comm())
        dist.barrier()
        q.put((rank, st.bytes_out, st.p_top, st.total_ms))
    finally:
        dist.destroy_process_group()


def test_two_gpu_file_run_equals_single_rank_file(engine_lib, tmp_path):
    """mpassit_run on two GPUs, one process each: both ranks map the same input files, regrid their row slab and
    pwrite it into the one output file (rank 0 creates it); P_TOP's max / min go through the host's communicator.
    The file equals the single-rank file byte for byte."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from mpassit_b200 import build, host, workload
    from tests import mpas_files
    from tests.test_gpu_files import _reference_pass

    build.build_host()
    host.load()
    wl = workload.make("mid", rundir=str(tmp_path))
    got, src, ter = _reference_pass(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), src, ter)
    st1 = host.run(nl, str(tmp_path), device=0)
    single = open(paths["out"], "rb").read()
    os.remove(paths["out"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_file_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][2] == res[1][2] == st1.p_top
    assert res[0][1] + res[1][1] == st1.bytes_out
    assert open(paths["out"], "rb").read() == single
    print("file run, mid workload: 1 GPU %.0f ms, 2 GPUs %.0f / %.0f ms" % (st1.total_ms, res[0][3], res[1][3]))
