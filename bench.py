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
KIND_NAMES = {0: "k_apply_pipe", 1: "k_apply_cols", 2: "k_apply_flat", 3: "k_apply_planes", 4: "k_apply_pipe (composed wind route)"}


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
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 # the remaining NVML reason bits, so that a clock below max is never reported without its cause
                 ("hw_power_brake_slowdown", 0x80), ("sync_boost", 0x10), ("applications_clocks_setting", 0x2),
                 ("display_clock_setting", 0x100))
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
                "reasons": [n for n, bit in names if mask & bit], "reasons_mask": hex(mask & ~0x1), "samples": len(use),
                "samples_in_timed_region": len(inside), "power_w_max": max(pw)}


# --------------------------------------------------------------------------- CPU reference arm
def host_threads() -> int:
    """Host cores this process may use (affinity mask), whatever OMP_NUM_THREADS says: torchrun exports
    OMP_NUM_THREADS=1 to its workers, which throttled the CPU arm to one core at N >= 2 in round 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_pass(wl, sources, frac_rows: float, steps: int, warmup: int, threads: int, keep: bool = False):
    """The reference's CPU path restated by oracle/ (ESMF cannot be built here): one interp_data pass = weight
    generation (RegridStore) + application (Regrid) of every field class, nothing hoisted out of the step
    (geometry conversion, dual triangles and search trees are rebuilt per step, as one `mpassit` run does).
    Returns (units per step, [seconds per step], detail, outputs of the last step if keep)."""
    from oracle import interp_oracle
    from oracle import oracle as orc

    orc.set_num_threads(threads)
    nj = wl.grids["M"][0].shape[0]
    rows = None
    if frac_rows < 1.0:
        nrows = max(2, int(round(nj * frac_rows)))
        j0 = (nj - nrows) // 2
        rows = (j0, j0 + nrows)
    times, units, det, out = [], 0, {}, None
    for k in range(warmup + steps):
        det = {}
        t0 = time.perf_counter()
        out = interp_oracle.interp_data(wl.mesh, wl.grids, sources, wl.cosa, wl.sina, wrf_mod_vars=bool(wl.cfg.wrf_mod_vars),
                                        lc=wl.cosa is not None, periodic=not wl.cfg.is_regional, rows=rows, keep=keep,
                                        tiled=not keep, timing=det)
        dt = time.perf_counter() - t0
        units = det["units"]
        if k >= warmup:
            times.append(dt)
    return units, times, det, out


def workload_name(wl) -> str:
    if wl.name == "c1":
        return "c1: 120-km global MPAS (40962 cells, 55 levels) -> 1 deg lat-lon, histlist_2d/3d/soil, one interp_data pass"
    return (f"{wl.name}: 3-km regional MPAS ({wl.mesh.nCells} cells, {wl.nz} levels) -> Lambert {wl.cfg.nx}x{wl.cfg.ny} "
            f"dx={wl.cfg.dxkm:.0f} m, diaglist+histlist_2d/3d/soil, one interp_data pass")


def config_block(wl, units: int, world: int) -> dict:
    """`config` of the JSON line: identical in both arms (the driver compares them)."""
    return {"workload": workload_name(wl), "units_per_step": int(units), "target_points": int(wl.n_mass),
            "cell_numbering": wl.mesh.meta.get("cell_order", "rowmajor (as generated)"),
            "l2_policy": "inputs (> 9 GB) exceed L2; no flush", "parallelism": f"target row slabs x{world}"}


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
    ap.add_argument("--no-numbering", action="store_true", help="skip the Z-order / random cell-numbering passes")
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

    # ---------------------------------------------------------------- reference arm (CPU)
    # The reference's own CPU implementation of the path is ESMF (unbuildable here: no Fortran, no ESMF); this arm
    # times the oracle port on all host cores.  It never maps libmpassit_rg.so / libmpassit_host.so: the workload
    # comes from oracle/ref_workload.py (numpy projection restatement + the synthetic mesh generator).
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle as orc
        from oracle import ref_workload

        orc.build()
        threads = int(os.environ.get("MPASSIT_BENCH_CPU_THREADS", "0")) or host_threads()
        wl = ref_workload.make(args.config)
        sources = ref_workload.synthetic_sources(wl)
        # One step = weights + apply for the whole target grid, every field class (~4-5 s of CPU on this workload).
        # Only when K + W such steps would not end within a few minutes is the row block shrunk (a smaller block is
        # dominated by the per-run fixed costs -- search trees over 2.4 M cells -- and understates the CPU).
        frac = float(os.environ.get("MPASSIT_BENCH_CPU_FRAC", "1.0"))
        budget_s = float(os.environ.get("MPASSIT_BENCH_CPU_BUDGET_S", "240"))
        _, t_probe, _, _ = cpu_reference_pass(wl, sources, frac, 1, 0, threads)
        nsteps = args.steps + max(args.warmup - 1, 0)
        if t_probe[0] * nsteps > budget_s:
            frac = max(0.02, frac * budget_s / (t_probe[0] * nsteps))
        units, times, det, _ = cpu_reference_pass(wl, sources, frac, args.steps, max(args.warmup - 1, 0), threads)
        t = sum(times) / len(times)
        v = units / t
        sample = (f"{det['rows']}/{det['of_rows']} target rows x all fields; weights {det['weights_s']:.2f}s + apply "
                  f"{det['apply_s']:.2f}s per step; CPU restatement of the ESMF path (oracle/), not ESMF; gcc -O3, OpenMP")
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(wl, wl.units_per_pass(), args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return 0

    from mpassit_b200 import build, workload
    from mpassit_b200 import lib as L

    # ---------------------------------------------------------------- this engine
    import dataclasses

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
    from mpassit_b200 import synth
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
    units = wl.units_per_pass()
    log(f"[rank {rank}] setup {time.perf_counter() - t_setup:.1f}s: {wl.mesh.nCells} cells -> "
        f"{wl.grids['M'][0].shape[::-1]} mass points, {units:.3e} units/pass")

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not dist:
            return x
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def all_ok(flag: bool) -> bool:   # every rank takes the same branch, whatever failed where
        if not dist:
            return flag
        t = torch.tensor([1 if flag else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def build_routes(r):
        """weights once (memoised for the device-resident steps); the handles keep them resident"""
        held, ms = [], {}
        for tag, (m, s_, d) in {"bilinear": (L.BILINEAR, L.SRC_MESH_ELEMENT, L.CENTER),
                                "nearest": (L.NEAREST_STOD, L.SRC_MESH_ELEMENT, L.CENTER),
                                "conserve": (L.CONSERVE, L.SRC_MESH_ELEMENT, L.CENTER),
                                "stagger_u": (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE1),
                                "stagger_v": (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE2)}.items():
            held.append(r.store(m, s_, d))
            ms[tag] = r.last_ms
        # the wind chain as one matrix per staggered grid (None where the grid is not composable: c1)
        for tag, st in (("wind_u", L.EDGE1), ("wind_v", L.EDGE2)):
            h = r.store_wind(st)
            if h is not None:
                held.append(h)
                ms[tag] = r.last_ms
        r.synchronize()
        return held, ms

    def rooflines(prof, ms_total, steps):
        """Per-launch CUDA events of the timed steps -> the dominant launch class and the per-kernel table.  A launch
        class = (kernel, units per launch)."""
        by_kind, by_class = {}, {}
        for r_ in prof:
            for table, key in ((by_kind, r_["kind"]), (by_class, (r_["kind"], int(r_["units"])))):
                k = table.setdefault(key, dict(ms=0.0, bytes=0.0, units=0.0, n=0))
                k["ms"] += r_["ms"]; k["bytes"] += r_["alg_bytes"]; k["units"] += r_["units"]; k["n"] += 1
        if not by_class:
            return None
        dom = max(by_class, key=lambda k: by_class[k]["ms"])
        d = by_class[dom]
        ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        all_bytes = sum(v["bytes"] for v in by_kind.values())
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": KIND_NAMES[dom[0]], "launches": d["n"], "avg_launch_ms": d["ms"] / d["n"],
                "alg_bytes_per_launch": d["bytes"] / d["n"], "units_per_launch": d["units"] / d["n"],
                "share_of_step": d["ms"] / ms_total, "peak_source": peak_source,
                "whole_pass": {"alg_bytes_per_step": all_bytes / steps, "GBps": all_bytes / (ms_total * 1e-3) / 1e9,
                               "frac": all_bytes / (ms_total * 1e-3) / 1e9 / peak,
                               "note": "algorithmic bytes of every launch of the pass / device time of the pass"},
                "kernels": {KIND_NAMES[k]: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["n"] / steps,
                                            "GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                            "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak}
                            for k, v in by_kind.items()}}

    def timed_pass(r, step, steps, warm, sample_clocks=None):
        """warm untimed passes, then `steps` passes between CUDA events (device time, max over ranks)."""
        for _ in range(warm):
            step()
        barrier()
        r.profile(True)
        n0 = r.kernel_launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sample_clocks is not None:
            sample_clocks.t0 = time.perf_counter()
        ev0.record()
        t_issue = time.perf_counter()
        for _ in range(steps):
            step()
        t_issue = time.perf_counter() - t_issue   # host time to enqueue the passes (no sync inside)
        ev1.record()
        barrier()
        if sample_clocks is not None:
            sample_clocks.t1 = time.perf_counter()
        ms_total = ev0.elapsed_time(ev1)
        launches = r.kernel_launches - n0
        prof = r.profile_read()
        r.profile(False)
        ms_total = max_over_ranks(ms_total)
        return dict(ms_step=ms_total / steps, ms_total=ms_total, launches=launches, t_issue=t_issue,
                    roofline=rooflines(prof, ms_total, steps))

    t0 = time.perf_counter()
    held, store_ms = build_routes(rg)
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
            for s_ in F["dev"][g]:
                if g == "hist_3d" and s_.name in ("uReconstructZonal", "uReconstructMeridional"):
                    continue
                gather_items.append((L.CENTER, s_.nlev, s_.dst, full(s_.nlev, wl.n_mass)))
        gather_items.append((L.CENTER, 1, F["dev"]["hgt"], full(1, wl.n_mass)))
        gather_items.append((L.EDGE1, wl.nz, F["dev"]["u_stag"], full(wl.nz, wl.grids["U"][0].size)))
        gather_items.append((L.EDGE2, wl.nz, F["dev"]["v_stag"], full(wl.nz, wl.grids["V"][0].size)))

    device_step = workload.prepare_interp(rg, wl, F["dev"], L.DEVICE)   # argument block marshalled once: one C call per pass

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main = timed_pass(rg, device_step, args.steps, max(args.warmup, 3), sampler)
    clocks = sampler.stop() if rank == 0 else None
    ms_step, launches, roofline = main["ms_step"], main["launches"], main["roofline"]
    value = units / (ms_step * 1e-3)
    if roofline is not None:
        # DRAM bytes of the dominant launch from the committed `ncu --set full` capture of this command
        # (None if the capture was taken on another launch shape, e.g. a different --gpus)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")))
            for c in tj["launch_classes"]:
                if c["kernel"] == roofline["kernel"] and int(c["units_per_launch"]) == int(roofline["units_per_launch"]):
                    roofline["traffic"] = c["dram_bytes_read"] + c["dram_bytes_write"]
                    roofline["traffic_source"] = tj.get("source", "ncu --set full, profiles/r02")
        except Exception:  # noqa: BLE001
            pass

    # The same pass at the reference's own arithmetic: fp32 fields accumulate and rotate in fp64 (R8), one rounding
    # on store (mprg_set_option "accumulate" = "f64"); the mass-point winds stay fp64 between the applies.
    f64acc = None
    try:
        rg.set_option("accumulate", "f64")
        step64 = workload.prepare_interp(rg, wl, F["dev"], L.DEVICE)
        r64 = timed_pass(rg, step64, args.steps, 3)
        f64acc = {"value": units / (r64["ms_step"] * 1e-3), "unit": UNIT, "ms_per_step": r64["ms_step"], "roofline": r64["roofline"],
                  "dtype": "f64 accumulate + rotation (the reference's R8 arithmetic), f32 in/out"}
        del step64
    except Exception as e:  # noqa: BLE001
        f64acc = {"error": str(e)[:200]}
    finally:
        rg.set_option("accumulate", "f32")
    for _ in range(2):
        device_step()          # the outputs the parity leg compares are those of the default arithmetic
    torch.cuda.synchronize()

    # The same cells and fields with the mesh numbered along a Z-order curve and with a random numbering: the column
    # kernel fetches runs of consecutively numbered cells with one TMA copy each, so its speed depends on the numbering
    # (and a random numbering puts every 240-byte column in its own DRAM page, which no kernel can stream).
    numbering = {}
    if world == 1 and args.config != "c1" and not args.no_numbering:
        for order in ("morton", "random"):
            try:
                wl2 = dataclasses.replace(wl, mesh=synth.renumber_cells(wl.mesh, order))
                rg2 = Regridder(device=local_rank)
                rg2.use_torch_stream()
                workload.load_geometry(rg2, wl2)
                held2, _ = build_routes(rg2)
                info2 = held2[0].info()
                step2 = workload.prepare_interp(rg2, wl2, F["dev"], L.DEVICE)
                r2 = timed_pass(rg2, step2, args.steps, 3)
                numbering[order] = {"value": units / (r2["ms_step"] * 1e-3), "ms_per_step": r2["ms_step"],
                                    "frac": r2["roofline"]["frac"] if r2["roofline"] else None,
                                    "whole_pass_frac": r2["roofline"]["whole_pass"]["frac"] if r2["roofline"] else None,
                                    "runs_per_tile": round(info2["tile_runs"] / max(info2["tiles"], 1), 2)}
                for h in held2:
                    h.release()
                del step2
                rg2.close()
            except Exception as e:  # noqa: BLE001
                numbering[order] = {"error": str(e)[:200]}
        for _ in range(2):
            device_step()
        torch.cuda.synchronize()

    # The same pass replayed from a CUDA graph (mprg_capture_*): one launch call per pass instead of several kernel
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
        graph_n0 = rg.kernel_launches
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            rg.graph_launch(graph)
        q1.record()
        barrier()
        qms = max_over_ranks(q0.elapsed_time(q1) / args.steps)
        graph_replay = {"ms_per_step": qms, "value": units / (qms * 1e-3), "unit": UNIT,
                        "gpu_launches": rg.kernel_launches - graph_n0}
    else:
        graph_replay = {"error": gerr or "capture failed on another rank"}
    if graph is not None:
        rg.graph_release(graph)

    gather = None
    value_incl_gather = None
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
        gms = max_over_ranks(g0.elapsed_time(g1) / nrep)
        total = sum(int(it[1]) * (wl.n_mass if it[0] == L.CENTER else wl.grids["U" if it[0] == L.EDGE1 else "V"][0].size) * 4
                    for it in gather_items)
        to_root = total * (world - 1) / world
        # multi-GPU time = first kernel launch -> completion of the gather on the root
        barrier()
        g0.record()
        for _ in range(nrep):
            device_step()
            rg.gather_many(gather_items, 0)
        g1.record()
        barrier()
        pgms = max_over_ranks(g0.elapsed_time(g1) / nrep)
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
            fms = max_over_ranks(f0.elapsed_time(f1) / args.steps)
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
                  "pass_plus_nccl_gather_ms": pgms,
                  "nvlink_peak_GBps": 770.0, "peak_source": "B200_PROFILING.md measured peer copy, per direction",
                  "fields": len(gather_items), "note": "slabs of every output field -> rank 0, one NCCL group; not part of "
                  "interp_data (the reference gathers in write_to_file), so not inside `value`.  A gather to ONE rank is "
                  "bounded by that rank's NVLink ingress (bytes_into_root / 770 GB/s) whatever N is: it cannot scale, which "
                  "is why the product path hands each rank's slab to the host / the file directly (DESIGN.md 4)"}
        best = min(pgms, fused["ms_per_pass"]) if fused and "ms_per_pass" in fused else pgms
        value_incl_gather = {"value": units / (best * 1e-3), "unit": UNIT, "ms_per_step": best,
                             "how": "fused peer stores" if fused and fused.get("ms_per_pass") == best else "pass + NCCL gather",
                             "note": "SURVEY.md 8(d) multi-GPU time: first kernel launch -> results complete on the writing rank"}

    # end to end: host buffers through the C ABI, weights rebuilt each step
    e2e = None
    if want_e2e:
        io0 = rg.io_bytes()
        for r_ in held:
            r_.release()
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
        h2d, d2h = [(b_ - a_) // (e2e_warm + args.e2e_steps) for a_, b_ in zip(io0, io1)]
        t = max_over_ranks(t)
        if dist:
            bb = torch.tensor([h2d, d2h], device="cuda", dtype=torch.int64)
            dist.all_reduce(bb, op=dist.ReduceOp.SUM)
            h2d, d2h = int(bb[0].item()), int(bb[1].item())
        nominal = workload.io_bytes(wl, F["host"])
        e2e = {"value": units / t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * t, "ms_per_step_all": [round(1e3 * x, 2) for x in ts], "stat": "median",
               "includes": "weight generation + H2D + apply + D2H, pinned host buffers, all ranks",
               "source_bytes_per_rank_if_replicated": nominal[0],
               "note": "sources are halo-sharded: each rank uploads only the cell-id range its slab's weights reference"}
        # the two paths (device-resident vs host-buffer) give the same bits
        a = F["dev"]["hist_3d"][2].dst.cpu()
        b = F["host"]["hist_3d"][2].dst
        e2e["equals_device_resident_pass"] = bool(torch.equal(a, b))
        if not e2e["equals_device_resident_pass"]:
            log("WARNING: device-resident and host-buffer results differ")

    # CPU baseline + parity: the oracle's interp_data over the SAME sources at the full size, timed, and its outputs
    # compared with the GPU pass field by field, element by element.
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from mpassit_b200 import check
        from oracle import oracle as orc

        orc.build()
        threads = host_threads()
        fields = {g: [(s_.name, s_.src.cpu().numpy()) for s_ in F["dev"][g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
        fields["ter"] = F["dev"]["ter"].cpu().numpy()
        frac = float(os.environ.get("MPASSIT_BENCH_CPU_FRAC", "1.0"))
        u, times, det, want = cpu_reference_pass(wl, fields, frac, 1, 0, threads, keep=(frac >= 1.0))
        cpu = {"value": u / times[0], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{det['rows']}/{det['of_rows']} target rows x all fields (weights {det['weights_s']:.2f}s + "
                         f"apply {det['apply_s']:.2f}s, outputs kept for the parity check); CPU restatement of the ESMF "
                         f"path, not ESMF"}
        if want is not None:
            got = {s_.name: s_.dst for g in ("diag", "hist_2d", "hist_3d", "soil") for s_ in F["dev"][g]}
            got["HGT"], got["U"], got["V"] = (F["dev"][k] for k in ("hgt", "u_stag", "v_stag"))
            exact_names = {"xland", "ivgtyp", "isltyp", "landmask"} | {nm for nm, _ in wl.lists["soil"]}
            res, not_exact = {}, []
            for nm in list(want):
                w_ = want.pop(nm)
                if nm.startswith("uReconstruct"):
                    continue
                g_ = got[nm].cpu().numpy().reshape(w_.shape)
                res[nm] = check.field_errors(g_, w_)
                if nm in exact_names and not res[nm]["exact"]:
                    not_exact.append(nm)
            parity = check.summarize(res)
            parity["nearest_fields_not_bit_exact"] = not_exact
            parity["against"] = "oracle/interp_oracle.py (CPU restatement, fp64) on the same sources, full 1801x1061 grid"
            if parity["failed"] or not_exact:
                log(f"WARNING: parity failures: {parity['failed']} {not_exact}")

    # `value`: the device-resident pass the way a time loop runs it -- recorded once, replayed per output time with one
    # launch call (mprg_graph_launch) -- unless the eager pass (whose per-launch events give the roofline) was faster;
    # both are reported, `engine.value_path` says which one `value` is.
    eager = {"ms_per_step": ms_step, "value": value, "unit": UNIT, "gpu_launches": launches,
             "note": "the same pass issued launch by launch, with the per-launch CUDA events of the roofline measurement"}
    how = "eager launches"
    if graph_replay and "ms_per_step" in graph_replay and graph_replay["ms_per_step"] <= ms_step:   # (both are K timed steps of the same pass)
        ms_step, value, launches, how = graph_replay["ms_per_step"], graph_replay["value"], graph_replay["gpu_launches"], "CUDA-graph replay of the pass"
    composed_wind = "wind_u" in store_ms and "wind_v" in store_ms
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_block(wl, units, world),
            "engine": {"weights": "resident (memoised) in `value`; rebuilt per step in `e2e`", "value_path": how,
                       "wind_path": ("composed: stagger x rotation x bilinear as one matrix per staggered grid; the mass-point winds "
                                     "UMASS / VMASS are never materialised" if composed_wind else "chain: regrid, rotate, regrid"),
                       "units_note": "units_per_step counts the reference's regrid outputs, the mass-point winds included (both arms)",
                       "value_materialised_outputs_only": (units - 2 * wl.nz * wl.n_mass) / (ms_step * 1e-3) if composed_wind else value,
                       "tma_copies_per_tile": round(info["tile_runs"] / max(info["tiles"], 1), 2),
                       "columns_per_tile": round(info["tile_columns"] / max(info["tiles"], 1), 2)},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e,
            "value_f64acc": f64acc["value"] if f64acc and "value" in f64acc else None,
            "roofline_f64acc": f64acc["roofline"] if f64acc and "roofline" in f64acc else None, "f64acc": f64acc,
            "value_morton": numbering.get("morton", {}).get("value"), "frac_morton": numbering.get("morton", {}).get("frac"),
            "value_random": numbering.get("random", {}).get("value"), "frac_random": numbering.get("random", {}).get("frac"),
            "numbering": numbering or None,
            "value_incl_gather": value_incl_gather,
            "gather": gather, "graph_replay": graph_replay, "eager": eager, "file_run": file_run,
            "host_issue_ms_per_step": round(main["t_issue"] / args.steps * 1e3, 4),
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
