"""Steady-state timing of weight generation (mprg_store) on a workload: every route is rebuilt
`--iters` times after one warm-up build; prints device ms (CUDA events) and wall ms per route."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mpassit_b200 import lib as L  # noqa: E402
from mpassit_b200 import workload  # noqa: E402
from mpassit_b200.regrid import Regridder  # noqa: E402

ROUTES = {"bilinear": (L.BILINEAR, L.SRC_MESH_ELEMENT, L.CENTER), "nearest": (L.NEAREST_STOD, L.SRC_MESH_ELEMENT, L.CENTER),
          "conserve": (L.CONSERVE, L.SRC_MESH_ELEMENT, L.CENTER), "stagger_u": (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE1),
          "stagger_v": (L.BILINEAR, L.SRC_GRID_CENTER, L.EDGE2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    wl = workload.make(a.config)
    rg = Regridder(0)
    t0 = time.perf_counter()
    workload.load_geometry(rg, wl)
    print(f"set_mesh + set_target x4 + rotation: {1e3 * (time.perf_counter() - t0):.1f} ms wall")
    for it in range(a.iters + 1):
        rg.clear_routes()
        torch.cuda.synchronize()
        line = []
        for name, key in ROUTES.items():
            t0 = time.perf_counter()
            r = rg.store(*key)
            rg.synchronize()
            wall = 1e3 * (time.perf_counter() - t0)
            line.append(f"{name} {rg.last_ms:7.2f}/{wall:7.2f}")
            r.release()
        print(("warm-up " if it == 0 else f"iter {it}  ") + "  ".join(line) + "   (device ms / wall ms)")
    # BASELINE.json configs[4]'s own step: conservative weights rebuilt + snow/snowh applied, every iteration
    n, nd = wl.mesh.nCells, wl.n_mass
    snow = torch.rand((n, 1), device="cuda")
    snowh = torch.rand((n, 1), device="cuda")
    o1, o2 = torch.empty((1, nd), device="cuda"), torch.empty((1, nd), device="cuda")
    ts = []
    for it in range(a.iters + 2):
        rg.clear_routes()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = rg.store(*ROUTES["conserve"])
        rg.apply(r, [snow, snowh], [o1, o2], nlev=[1, 1])
        rg.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
        r.release()
    best = min(ts[2:])
    print(f"conservative weights + apply(snow, snowh), rebuilt per iteration: {best:.2f} ms wall "
          f"({nd / best * 1e3:.3e} destination cells/s; all: {[round(t, 2) for t in ts]})")
    rg.close()


if __name__ == "__main__":
    main()
