"""Where an end-to-end (host-buffer) pass spends its time on N ranks of one box (run under torchrun; profiling aid).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      profiles/e2e_multi.py [--bind 0|1] [--config c2]

Per rank, all ranks at the same time:
  1. raw page-locked copies of the rank's share (H2D alone, D2H alone, both at once) -> what the host's PCIe / memory
     system delivers to N GPUs together, before and after binding the process to its GPU's NUMA node
     (mprg_host_bind_to_device);
  2. one interp_data pass with host buffers and weights rebuilt (what bench.py's e2e times), with the wall time of every
     store / apply call (MPASSIT_TRACE) and the time the last upload, kernel and download finished.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mpassit_b200 import lib as L  # noqa: E402
from mpassit_b200 import workload  # noqa: E402
from mpassit_b200.regrid import Regridder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--bind", type=int, default=1)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def log0(*s):
        if rank == 0:
            print(*s, flush=True)

    cpus0 = sorted(os.sched_getaffinity(0))
    nbytes = 1 << 30

    def raw(tag):
        h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_in.fill_(1)
        d_in, d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        res = {}
        for name in ("h2d", "d2h", "both"):
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                if name in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if name in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = allmax((time.perf_counter() - t0) / 3)
            res[name] = (2 if name == "both" else 1) * nbytes * world / dt / 1e9
        log0(f"[raw {tag}] {world} ranks x 1 GiB: H2D {res['h2d']:.1f} GB/s, D2H {res['d2h']:.1f} GB/s, both {res['both']:.1f} GB/s aggregate")
        del h_in, h_out, d_in, d_out

    raw("unbound")
    node = -1
    if a.bind:
        node = L.host_bind_to_device(lr)
    print(f"[rank {rank}] gpu {lr} numa node {node}; cpus before {cpus0[0]}..{cpus0[-1]} ({len(cpus0)}), after "
          f"{min(os.sched_getaffinity(0))}..{max(os.sched_getaffinity(0))} ({len(os.sched_getaffinity(0))})", flush=True)
    if a.bind:
        raw("numa-bound")

    wl = workload.make(a.config)
    rg = Regridder(device=lr, rank=rank, nranks=world)
    workload.load_geometry(rg, wl)
    F = workload.make_fields(wl, device=f"cuda:{lr}", pinned_host=True, rg=rg)
    step = workload.prepare_interp(rg, wl, F["host"], L.HOST)
    for k in range(2 + a.iters):
        rg.clear_routes()
        barrier()
        trace = k == 2 + a.iters - 1 and rank in (0, world - 1)
        if trace:
            rg.set_option("trace", "1")
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if trace:
            rg.set_option("trace", "0")
        mx = allmax(dt)
        if k >= 2:
            print(f"[rank {rank}] pass {k}: {1e3 * dt:.1f} ms (max over ranks {1e3 * mx:.1f})", flush=True)
    io = rg.io_bytes()
    print(f"[rank {rank}] bytes per pass: h2d {io[0] / (2 + a.iters) / 1e9:.2f} GB d2h {io[1] / (2 + a.iters) / 1e9:.2f} GB", flush=True)
    # stores alone (weights rebuilt, nothing else running)
    rg.clear_routes()
    barrier()
    t0 = time.perf_counter()
    ms = {}
    for tag, key in (("bilinear", (L.BILINEAR, L.SRC_MESH_ELEMENT, L.CENTER)), ("halo", (L.BILINEAR, L.SRC_MESH_ELEMENT, L.CENTER_HALO)),
                     ("stag_u", (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE1)), ("stag_v", (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE2)),
                     ("conserve", (L.CONSERVE, L.SRC_MESH_ELEMENT, L.CENTER)), ("nearest", (L.NEAREST_STOD, L.SRC_MESH_ELEMENT, L.CENTER))):
        t1 = time.perf_counter()
        r = rg.store(*key)
        ms[tag] = round(1e3 * (time.perf_counter() - t1), 2)
        r.release()
    print(f"[rank {rank}] stores alone (host wall ms): {ms} total {1e3 * (time.perf_counter() - t0):.1f}", flush=True)
    barrier()
    rg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
