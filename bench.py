#!/usr/bin/env python
"""bench.py -- interpolated target-point-levels/s of one MPASSIT interp_data pass.

  python bench.py --gpus N --steps K --warmup W            # this engine (N = 1 default)
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference's ESMF path

A *step* is one pass of the hot path (interp_data, /root/reference/interp.F90:92-465) over
one synthetic MPAS output time: every field of diaglist + histlist_2d/3d/soil on the
BASELINE.json configs[1] workload (3-km regional mesh -> Lambert 1801 x 1061, 60 levels).
  value : device-resident pass (sources, weights and outputs in HBM), CUDA-event timed
  e2e   : the same pass through the C ABI with HOST buffers, weights rebuilt every step
          (what one `mpassit` run does), H2D + D2H inside the timed region
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "interpolated target-point-levels/s"
UNIT = "point-levels/s"
KIND_NAMES = {0: "k_apply_pipe<aligned>", 1: "k_apply_pipe<unaligned>", 2: "k_apply_flat", 3: "k_apply_planes"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled every ~5 ms by an NVML thread while the timed region runs
    (the region lasts well under a second, too short for `nvidia-smi -lms`)."""

    def __init__(self, gpu_index: int = 0):
        self.idx = gpu_index
        self.samples = []       # (t, sm_mhz, power_w, reasons bitmask)
        self.t0 = self.t1 = None  # timed region (time.perf_counter)
        self._stop = False
        self._thr = None
        self.err = None

    def start(self):
        import threading

        try:
            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it holds plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()]
            phys = ids[self.idx] if self.idx < len(ids) else self.idx
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = f"nvml unavailable: {e}"
            return

        def loop():
            while not self._stop:
                try:
                    self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                         nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                         int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.005)

        self._nv = nv
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self) -> dict:
        if self._thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not started"]}
        self._stop = True
        self._thr.join(timeout=2)
        nv = self._nv
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        # samples inside the timed region; a region shorter than a few NVML calls falls back to the
        # samples of the warm-up passes right before it (same kernels, same load)
        inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= self.t1]
        use = inside if len(inside) >= 3 else self.samples
        mask = 0
        for _, _, _, r in use:
            mask |= r
        pw = [p for _, _, p, _ in use]
        thr = 0.5 * (min(pw) + max(pw))     # under load = samples in the upper half of the power range seen
        load = [c for _, c, p, _ in use if p >= thr] or [c for _, c, _, _ in use]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in names if mask & bit], "samples": len(use),
                "samples_in_timed_region": len(inside), "power_w_max": max(pw)}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pass(wl, frac_rows: float, steps: int, warmup: int, threads: int | None = None):
    """The reference's CPU path restated by oracle/ (ESMF cannot be built here): weight
    generation (RegridStore) + application (Regrid) for a contiguous block of target rows,
    every field class of the workload.  Returns (units per step, [seconds per step], detail)."""
    from oracle import oracle as orc
    from mpassit_b200 import synth

    if threads:
        orc.set_num_threads(threads)
    m = wl.mesh
    lo, la = orc.mesh_rad_to_deg(m.lonCell, m.latCell)
    cxyz = orc.sph_deg_to_cart(lo, la)
    lov, lav = orc.mesh_rad_to_deg(m.lonVertex, m.latVertex)
    vxyz = orc.sph_deg_to_cart(lov, lav)
    latM, lonM = wl.grids["M"]
    nj, ni = latM.shape
    nrows = max(2, int(round(nj * frac_rows)))
    j0 = (nj - nrows) // 2
    rows = slice(j0, j0 + nrows)
    rng = np.random.default_rng(1)
    # source fields (values do not affect timing; one array per distinct level count is reused)
    src = {n: synth.smooth_field(m.lonCell, m.latCell, n, seed=n) for n in {1, wl.nz, wl.nz + 1, wl.nsoil}}
    lists = wl.lists
    wrf = bool(wl.cfg.wrf_mod_vars)
    detail = {}

    def one_pass():
        t0 = time.perf_counter()
        tri = orc.dual_triangles(m.verticesOnCell, m.nVertices)
        dM = orc.sph_deg_to_cart(lonM[rows], latM[rows])
        e, c, w = orc.bilinear(cxyz, tri, m.verticesOnCell, dM)
        bil = orc.ell_to_csr(e >= 0, c, w)
        nst = orc.nearest_to_csr(orc.nearest(cxyz, dM))
        clat, clon = wl.grids["CORNER"]
        cor = orc.sph_deg_to_cart(clon[j0:j0 + nrows + 1], clat[j0:j0 + nrows + 1]).reshape(nrows + 1, ni + 1, 3)
        cons = orc.conserve(cxyz, vxyz, m.verticesOnCell, cor)
        sx = dM.reshape(nrows, ni, 3)
        ulat, ulon = wl.grids["U"]
        vlat, vlon = wl.grids["V"]
        eu, cu, wu = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(ulon[rows], ulat[rows]))
        ev, cv, wv = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(vlon[j0:j0 + nrows + 1], vlat[j0:j0 + nrows + 1]))
        ucsr, vcsr = orc.ell_to_csr(eu >= 0, cu, wu), orc.ell_to_csr(ev >= 0, cv, wv)
        t1 = time.perf_counter()
        units = 0
        for nm, _ in lists["diag"]:
            n = wl.levels_of("diag", nm)
            orc.apply(*bil, src[n], np.float32, tiled=True); units += n * dM.shape[0]
        for nm, _ in lists["hist_2d"]:
            csr = cons if nm in ("snow", "snowh") else nst if nm in ("xland", "ivgtyp", "isltyp", "landmask") else bil
            orc.apply(*csr, src[1], np.float32, tiled=True); units += dM.shape[0]
        orc.apply(*bil, src[1], np.float32, tiled=True); units += dM.shape[0]     # HGT
        winds = {}
        for nm, _ in lists["hist_3d"]:
            n = wl.levels_of("hist_3d", nm)
            if wrf and nm in ("uReconstructZonal", "uReconstructMeridional"):
                winds[nm] = orc.apply(*bil, src[n], np.float64); units += n * dM.shape[0]
            else:
                orc.apply(*bil, src[n], np.float32, tiled=True); units += n * dM.shape[0]
        if len(winds) == 2 and wl.cosa is not None:
            orc.rotate_winds(winds["uReconstructZonal"], winds["uReconstructMeridional"],
                             wl.cosa[rows].reshape(-1), wl.sina[rows].reshape(-1))
        for nm, csr in (("uReconstructZonal", ucsr), ("uReconstructMeridional", vcsr)):
            if nm in winds:
                orc.apply_planes(*csr, winds[nm]); units += wl.nz * (csr[0].size - 1)
        for nm, _ in lists["soil"]:
            orc.apply(*nst, src[wl.nsoil], np.float32, tiled=True); units += wl.nsoil * dM.shape[0]
        t2 = time.perf_counter()
        detail.update(weights_s=t1 - t0, apply_s=t2 - t1, rows=nrows, of_rows=nj)
        return units, t2 - t0

    times, units = [], 0
    for k in range(warmup + steps):
        units, dt = one_pass()
        if k >= warmup:
            times.append(dt)
    return units, times, detail


def workload_name(wl) -> str:
    if wl.name == "c1":
        return "c1: 120-km global MPAS (40962 cells, 55 levels) -> 1 deg lat-lon, histlist_2d/3d/soil, one interp_data pass"
    return (f"{wl.name}: 3-km regional MPAS ({wl.mesh.nCells} cells, {wl.nz} levels) -> Lambert {wl.cfg.nx}x{wl.cfg.ny} "
            f"dx={wl.cfg.dxkm:.0f} m, diaglist+histlist_2d/3d/soil, one interp_data pass")


# --------------------------------------------------------------------------- main
def measure_file_run(local_rank: int) -> dict:
    """`mpassit <namelist>` with NetCDF-classic files on both sides (host/run.cpp) on the 12-km
    miniature of the workload (0.6 GB in, 0.6 GB out): reported beside the metric, not part of it.  The 3-km case at
    full size (10 GB each way) is profiles/file_bench.py c2."""
    import shutil
    import tempfile

    fdir = tempfile.mkdtemp(prefix="mpassit_bench_files_")
    try:
        from mpassit_b200 import host, mpas_files, workload

        host.load()
        fwl = workload.make("mid", rundir=fdir)
        FF = workload.make_fields(fwl, device=f"cuda:{local_rank}")["dev"]
        fsrc = {g: [(s.name, s.src.cpu().numpy()) for s in FF[g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
        fnl, _ = mpas_files.write_case(fwl, fdir, fsrc, FF["ter"].cpu().numpy())
        del FF, fsrc
        runs = [host.run(fnl, fdir, device=local_rank) for _ in range(4)]
        st = min(runs[1:], key=lambda r: r.total_ms)   # the first run creates the CUDA context
        out = {"workload": workload_name(fwl), "ms_total": round(st.total_ms, 1),
               "ms": {"setup": round(st.setup_ms, 1), "read": round(st.read_ms, 1), "interp": round(st.interp_ms, 1),
                      "write": round(st.write_ms, 1)},
               "bytes_in": st.bytes_in, "bytes_out": st.bytes_out, "value": fwl.units_per_pass() / (st.total_ms * 1e-3),
               "unit": UNIT, "stat": "best of 3 after one warm-up run", "output": f"CDF-{st.output_version}",
               "note": "files in the page cache; big-endian sources swapped in HBM; each rank pwrites its slab"}
        return out
    except Exception as ex:  # the metric does not depend on it
        return {"error": f"{type(ex).__name__}: {ex}"}
    finally:
        shutil.rmtree(fdir, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("MPASSIT_BENCH_CONFIG", "c2"))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the file-to-file run of the 12-km case (file_run)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, ...)
    # is sent to stderr instead, at the file-descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from mpassit_b200 import build, workload
    from mpassit_b200 import lib as L

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        build.build_all()
        wl = workload.make(args.config)
        from oracle import oracle as orc

        orc.build()
        # One step = weights + apply for a block of target rows, every field class.  The whole grid costs
        # ~5 s of CPU per step on this workload; when K + W such steps would not end within a few minutes
        # the block is shrunk (a smaller block is dominated by the per-run fixed costs -- search trees over
        # 2.4 M cells -- and understates the CPU, so the full grid is kept whenever it fits).
        frac = float(os.environ.get("MPASSIT_BENCH_CPU_FRAC", "1.0"))
        budget_s = float(os.environ.get("MPASSIT_BENCH_CPU_BUDGET_S", "150"))
        _, t_probe, _ = cpu_reference_pass(wl, frac, 1, 0)
        nsteps = args.steps + max(args.warmup - 1, 0)
        if t_probe[0] * nsteps > budget_s:
            frac = max(0.02, frac * budget_s / (t_probe[0] * nsteps))
        units, times, det = cpu_reference_pass(wl, frac, args.steps, max(args.warmup - 1, 0))
        t = sum(times) / len(times)
        v = units / t
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(wl),
                       "reference_arm": f"CPU restatement of the ESMF path (not ESMF: unbuildable here), weights + apply, "
                                        f"{det['rows']} of {det['of_rows']} target rows per step"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                             "sample": f"{det['rows']}/{det['of_rows']} target rows x all fields; weights "
                                       f"{det['weights_s']:.2f}s + apply {det['apply_s']:.2f}s per step"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return 0

    # ---------------------------------------------------------------- this engine
    import torch

    if not torch.cuda.is_available():
        log("bench.py: no CUDA device; the engine has no CPU fallback")
        return 2
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build.build_all()
    if dist:
        dist.barrier()
    from mpassit_b200.regrid import Regridder

    # before the 3-km workload takes its 19 GB of device and 17 GB of pinned host memory
    file_run = measure_file_run(local_rank) if (rank == 0 and world == 1 and not args.no_files) else None

    t_setup = time.perf_counter()
    wl = workload.make(args.config)
    rg = Regridder(device=local_rank, rank=rank, nranks=world)
    # a dedicated (non-default) stream: engine kernels and the timing events share it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rg.use_torch_stream()
    workload.load_geometry(rg, wl)
    if world > 1:
        ids = [rg.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        rg.comm_init(ids[0])
    want_e2e = not args.no_e2e
    F = workload.make_fields(wl, device=f"cuda:{local_rank}", pinned_host=want_e2e, rg=rg)
    log(f"[rank {rank}] setup {time.perf_counter() - t_setup:.1f}s: {wl.mesh.nCells} cells -> "
        f"{wl.grids['M'][0].shape[::-1]} mass points, {wl.units_per_pass():.3e} units/pass")

    # weights once (memoised for the device-resident steps); keep handles so they stay resident
    t0 = time.perf_counter()
    held = []
    store_ms = {}
    for tag, (m, s, d) in {"bilinear": (L.BILINEAR, L.SRC_MESH_ELEMENT, L.CENTER),
                           "nearest": (L.NEAREST_STOD, L.SRC_MESH_ELEMENT, L.CENTER),
                           "conserve": (L.CONSERVE, L.SRC_MESH_ELEMENT, L.CENTER),
                           "stagger_u": (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE1),
                           "stagger_v": (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE2)}.items():
        held.append(rg.store(m, s, d))
        store_ms[tag] = rg.last_ms
    rg.synchronize()
    store_wall = time.perf_counter() - t0
    info = held[0].info()

    # The path's only collective -- output slabs -> the writing rank (ESMF_FieldGather in write_to_file,
    # write_data.F90:1006-1453) -- is not part of interp_data: it is timed on its own below and
    # reported under "gather", all fields of the pass in one NCCL group.
    gather_items = []
    if world > 1:
        def full(nlev, n):
            return torch.empty((nlev, n), dtype=torch.float32, device="cuda") if rank == 0 else None

        for g in ("diag", "hist_2d", "hist_3d", "soil"):
            for s in F["dev"][g]:
                if g == "hist_3d" and s.name in ("uReconstructZonal", "uReconstructMeridional"):
                    continue
                gather_items.append((L.CENTER, s.nlev, s.dst, full(s.nlev, wl.n_mass)))
        gather_items.append((L.CENTER, 1, F["dev"]["hgt"], full(1, wl.n_mass)))
        gather_items.append((L.EDGE1, wl.nz, F["dev"]["u_stag"], full(wl.nz, wl.grids["U"][0].size)))
        gather_items.append((L.EDGE2, wl.nz, F["dev"]["v_stag"], full(wl.nz, wl.grids["V"][0].size)))

    device_step = workload.prepare_interp(rg, wl, F["dev"], L.DEVICE)   # argument block marshalled once: one C call per pass

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    def all_ok(flag: bool) -> bool:   # every rank takes the same branch, whatever failed where
        if not dist:
            return flag
        t = torch.tensor([1 if flag else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    rg.profile(True)
    n0 = rg.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.t0 = time.perf_counter()
    ev0.record()
    t_issue = time.perf_counter()
    for _ in range(args.steps):
        device_step()
    t_issue = time.perf_counter() - t_issue   # host time to enqueue the passes (no sync inside)
    ev1.record()
    barrier()
    sampler.t1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    launches = rg.kernel_launches - n0
    prof = rg.profile_read()
    rg.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    units = wl.units_per_pass()
    value = units / (ms_step * 1e-3)

    # The same pass replayed from a CUDA graph (mprg_capture_*): one launch call per pass instead of nine kernel
    # launches.  Reported beside `value`, which stays the eager pass whose launches the roofline events time.
    graph_replay = None
    gerr, graph = "", None
    try:
        rg.capture_begin()
        device_step()
        graph = rg.capture_end()
        for _ in range(3):
            rg.graph_launch(graph)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        gerr, graph = str(e)[:200], None
    if all_ok(graph is not None):
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            rg.graph_launch(graph)
        q1.record()
        barrier()
        qms = q0.elapsed_time(q1) / args.steps
        if dist:
            tt = torch.tensor([qms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            qms = float(tt.item())
        graph_replay = {"ms_per_step": qms, "value": units / (qms * 1e-3), "unit": UNIT}
    else:
        graph_replay = {"error": gerr or "capture failed on another rank"}
    if graph is not None:
        rg.graph_release(graph)

    gather = None
    if world > 1:
        rg.gather_many(gather_items, 0)      # warm-up (NCCL channel setup)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 5
        g0.record()
        for _ in range(nrep):
            rg.gather_many(gather_items, 0)
        g1.record()
        barrier()
        tt = torch.tensor([g0.elapsed_time(g1) / nrep], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gms = float(tt.item())
        total = sum(int(it[1]) * (wl.n_mass if it[0] == L.CENTER else wl.grids["U" if it[0] == L.EDGE1 else "V"][0].size) * 4
                    for it in gather_items)
        to_root = total * (world - 1) / world
        # The same pass with the gather FUSED into the store: rank 0 owns full-grid output fields, the other
        # ranks map them with CUDA IPC and their apply kernels write their rows straight into rank 0's memory
        # over NVLink (mprg_apply_into) -- compute and collective in one set of kernels, no slab round trip.
        fused = None
        err, full, Ff = "", None, None
        try:
            full = workload.full_outputs(wl, "cuda") if rank == 0 else None
            box = [workload.export_full(rg, full) if rank == 0 else None]
        except Exception as e:  # noqa: BLE001
            err, box = str(e)[:200], [None]
        dist.broadcast_object_list(box, src=0)
        try:
            if box[0] is None:
                raise RuntimeError(err or "the writing rank could not export its buffers")
            dstf = full if rank == 0 else workload.open_full(rg, box[0])
            Ff = workload.prepare_interp(rg, wl, workload.with_destinations(F["dev"], dstf), L.DEVICE, dst_full=True)
            for _ in range(3):
                Ff()
            torch.cuda.synchronize()
            ok_here = True
        except Exception as e:  # noqa: BLE001
            err, ok_here = str(e)[:200], False
        if all_ok(ok_here):
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.steps):
                Ff()
            f1.record()
            barrier()
            tt = torch.tensor([f0.elapsed_time(f1) / args.steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fms = float(tt.item())
            ok = True
            if rank == 0:   # spot check against the slab path + NCCL gather
                ref = gather_items[2][3]
                ok = bool(torch.equal(full["diag"][2], ref)) if ref is not None and len(full["diag"]) > 2 else True
            fused = {"ms_per_pass": fms, "value": units / (fms * 1e-3), "unit": UNIT, "matches_nccl_gather": ok,
                     "GBps_into_root": to_root / (fms * 1e-3) / 1e9,
                     "note": "interp_data with every rank storing its rows directly into rank 0's full fields "
                             "(CUDA IPC peer stores over NVLink): the result is complete on the writing rank when the pass ends"}
        else:
            fused = {"error": err or "CUDA IPC mapping failed on another rank"}
        try:
            if rank != 0:
                rg.ipc_close_all()
        except Exception:  # noqa: BLE001
            pass
        del full, Ff
        gather = {"ms_per_pass": gms, "bytes_into_root": to_root, "GBps_into_root": to_root / (gms * 1e-3) / 1e9, "fused": fused,
                  "nvlink_peak_GBps": 770.0, "peak_source": "B200_PROFILING.md measured peer copy, per direction",
                  "fields": len(gather_items), "note": "slabs of every output field -> rank 0, one NCCL group; not part of "
                  "interp_data (the reference gathers in write_to_file), so not inside `value`"}

    # roofline of the dominant kernel, from per-launch CUDA events inside the timed steps.  A launch
    # class = (kernel, units per launch): the same kernel also runs a few small launches per step
    # (2 wind fields into fp64, 3 four-level soil fields) that are reported separately below.
    by_kind, by_class = {}, {}
    for r in prof:
        for table, key in ((by_kind, r["kind"]), (by_class, (r["kind"], int(r["units"])))):
            k = table.setdefault(key, dict(ms=0.0, bytes=0.0, units=0.0, n=0))
            k["ms"] += r["ms"]; k["bytes"] += r["alg_bytes"]; k["units"] += r["units"]; k["n"] += 1
    dom = max(by_class, key=lambda k: by_class[k]["ms"]) if by_class else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = None
    if dom is not None:
        d = by_class[dom]
        ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        # DRAM bytes of the same launch class from the committed `ncu --set full` capture (None if the
        # capture was taken on another launch shape, e.g. a different --gpus)
        traffic = None
        try:
            for c in json.load(open(os.path.join(ROOT, "profiles", "r01", "ncu_traffic.json")))["launch_classes"]:
                if c["kernel"] == KIND_NAMES[dom[0]] and int(c["units_per_launch"]) == int(d["units"] / d["n"]):
                    traffic = c["dram_bytes_read"] + c["dram_bytes_write"]
        except Exception:  # noqa: BLE001
            pass
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": "ncu --set full, profiles/r01/apply_v6_ncu_summary.md" if traffic else None,
                    "kernel": KIND_NAMES[dom[0]], "launches": d["n"], "avg_launch_ms": d["ms"] / d["n"],
                    "alg_bytes_per_launch": d["bytes"] / d["n"], "units_per_launch": d["units"] / d["n"],
                    "share_of_step": d["ms"] / ms_total,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "kernels": {KIND_NAMES[k]: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["n"] / args.steps,
                                                "GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9}
                                for k, v in by_kind.items()}}

    # end to end: host buffers through the C ABI, weights rebuilt each step
    e2e = None
    if want_e2e:
        io0 = rg.io_bytes()
        for r in held:
            r.release()
        held = []
        ts = []
        e2e_warm = 2   # first passes size the staging ring and fault in the pool
        host_step = workload.prepare_interp(rg, wl, F["host"], L.HOST)
        for k in range(e2e_warm + args.e2e_steps):
            rg.clear_routes()
            barrier()
            t0 = time.perf_counter()
            host_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if k >= e2e_warm:
                ts.append(dt)
        t = statistics.median(ts)   # shared host: other tenants' PCIe traffic makes single passes jitter
        # bytes the engine actually copied per pass (counted at the cudaMemcpyAsync calls), summed over ranks
        io1 = rg.io_bytes()
        h2d, d2h = [(b - a) // (e2e_warm + args.e2e_steps) for a, b in zip(io0, io1)]
        if dist:
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
            bb = torch.tensor([h2d, d2h], device="cuda", dtype=torch.int64)
            dist.all_reduce(bb, op=dist.ReduceOp.SUM)
            h2d, d2h = int(bb[0].item()), int(bb[1].item())
        nominal = workload.io_bytes(wl, F["host"])
        e2e = {"value": units / t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * t, "ms_per_step_all": [round(1e3 * x, 2) for x in ts], "stat": "median",
               "includes": "weight generation + H2D + apply + D2H, pinned host buffers, all ranks",
               "source_bytes_per_rank_if_replicated": nominal[0],
               "note": "sources are halo-sharded: each rank uploads only the cell-id range its slab's weights reference"}
        # parity spot check of the two paths (device-resident vs host-buffer) on one field
        a = F["dev"]["hist_3d"][2].dst.cpu()
        b = F["host"]["hist_3d"][2].dst
        if not torch.equal(a, b):
            log("WARNING: device-resident and host-buffer results differ")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc

        orc.build()
        frac = float(os.environ.get("MPASSIT_BENCH_CPU_FRAC", "1.0"))
        u, times, det = cpu_reference_pass(wl, frac, 1, 0)
        cpu = {"value": u / times[0], "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
               "sample": f"{det['rows']}/{det['of_rows']} target rows x all fields (weights {det['weights_s']:.2f}s + "
                         f"apply {det['apply_s']:.2f}s); CPU restatement of the ESMF path, not ESMF"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if os.environ.get("MPASSIT_GPU_ACC", "") not in ("f64", "fp64") else "f64 accumulate, f32 in/out",
            "data": "synthetic",
            "config": {"workload": workload_name(wl),
                       "units_per_step": units, "target_points": wl.n_mass, "l2_policy": "inputs (>9 GB) exceed L2; no flush",
                       "parallelism": f"target row slabs x{world}", "weights": "resident (memoised) in `value`; rebuilt per step in `e2e`",
                       "cell_numbering": wl.mesh.meta.get("cell_order", "rowmajor (as generated)"),
                       "tma_copies_per_tile": round(info["tile_runs"] / max(info["tiles"], 1), 2),
                       "columns_per_tile": round(info["tile_columns"] / max(info["tiles"], 1), 2)},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gather": gather, "graph_replay": graph_replay, "file_run": file_run,
            "host_issue_ms_per_step": round(t_issue / args.steps * 1e3, 4),
            "store_ms": store_ms, "store_wall_s": store_wall,
            "route_bilinear": info,
        }
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    rg.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
