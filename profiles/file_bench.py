"""Stage times of `mpassit <namelist>` with files on both sides (mpassit_b200/host/run.cpp) on a synthetic workload.

    python profiles/file_bench.py [mini|mid|c2] [--runs N] [--dir DIR]

Writes the MPAS init / diag / history files of the workload (NetCDF classic, CDF-2, fp32) under DIR, runs the file
driver N times and prints, per run, the wall time of each stage as mpassit_run reports it: setup (namelist, target
coordinates, mesh, BVHs), read (header parsing + mapping: no data is copied), interp (upload of the mapped variables,
byte swap in HBM, weights, applies), write (post-ops, swap, download, pwrite).  The first run pays the page-cache
misses of the freshly written inputs and CUDA context creation; later runs are the steady state of a time loop.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="mid")
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--dir", default=None)
    a = ap.parse_args()

    from mpassit_b200 import build, host, workload
    from mpassit_b200 import mpas_files

    build.build_all()
    host.load()
    rundir = a.dir or tempfile.mkdtemp(prefix=f"mpassit_files_{a.workload}_")
    wl = workload.make(a.workload, rundir=rundir)
    t = time.time()
    F = workload.make_fields(wl, device="cuda:0")["dev"]
    src = {g: [(s.name, s.src.cpu().numpy()) for s in F[g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
    ter = F["ter"].cpu().numpy()
    del F
    import torch

    torch.cuda.empty_cache()
    nl, paths = mpas_files.write_case(wl, rundir, src, ter)
    del src
    sizes = {k: os.path.getsize(p) for k, p in paths.items() if k != "out"}
    print(f"workload {a.workload}: {wl.mesh.lonCell.size} cells, nz {wl.nz}, target {wl.cfg.i_target} x {wl.cfg.j_target}; "
          f"input files {sum(sizes.values()) / 1e9:.2f} GB written in {time.time() - t:.1f} s", flush=True)
    rows = []
    for r in range(a.runs):
        if os.path.exists(paths["out"]):
            os.remove(paths["out"])
        t = time.time()
        st = host.run(nl, rundir, device=0)
        wall = (time.time() - t) * 1e3
        row = dict(run=r, wall_ms=round(wall, 1), setup_ms=round(st.setup_ms, 1),
                   setup_split=[round(x, 1) for x in (st.init_ms, st.target_ms, st.gridfile_ms, st.mesh_ms)], read_ms=round(st.read_ms, 1), alloc_ms=round(st.alloc_ms, 1),
                   interp_ms=round(st.interp_ms, 1), write_ms=round(st.write_ms, 1),
                   write_split=[round(st.download_ms, 1), round(st.writer_wait_ms, 1)], total_ms=round(st.total_ms, 1),
                   in_GB=round(st.bytes_in / 1e9, 3), out_GB=round(st.bytes_out / 1e9, 3),
                   interp_GBps_in=round(st.bytes_in / 1e6 / max(st.interp_ms, 1e-9), 1),
                   write_GBps_out=round(st.bytes_out / 1e6 / max(st.write_ms, 1e-9), 1),
                   point_levels_per_s=round(st.bytes_out / 4 / max(st.total_ms, 1e-9) * 1e3, 0), vars=st.n_vars_written)
        rows.append(row)
        print(json.dumps(row), flush=True)
    best = min(rows, key=lambda r: r["total_ms"])
    print("best:", json.dumps(best))
    print(f"output file {os.path.getsize(paths['out']) / 1e9:.2f} GB, CDF-{rows[-1] and st.output_version}")


if __name__ == "__main__":
    main()
