"""Kernel-level benchmark of the batched apply on the C2 geometry (profiling aid).

  python profiles/kbench.py [--config c2] [--variants f64:3,f64:2,f32:4] [--iters 5] [--only-3d]

Builds the workload once, then for every variant (accumulate type : launch grouping [b<CTAs/SM>])
runs the bilinear-route stacked apply over the 3-D fields and prints per-kernel GB/s
from the engine's per-launch CUDA events.  Also used as the short command under ncu.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mpassit_b200 import lib as L  # noqa: E402
from mpassit_b200 import workload  # noqa: E402
from mpassit_b200.regrid import Regridder  # noqa: E402

KIND = {0: "pipe", 1: "cols_fallback", 2: "flat", 3: "planes"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--variants", default="f32:split,f32:one,f64:split")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--fields", type=int, default=12, help="number of stacked nz-level fields")
    ap.add_argument("--nlev", type=int, default=0, help="override level count (default: workload nz)")
    ap.add_argument("--stack", default="", help="explicit stack, e.g. 60x12,61x2 (levels x fields)")
    ap.add_argument("--rot", action="store_true", help="the first two fields are a wind pair with fused rotation")
    ap.add_argument("--order", default="rowmajor", choices=["rowmajor", "morton", "random"], help="cell numbering of the synthetic mesh")
    ap.add_argument("--method", default="bilinear", choices=["bilinear", "nearest"], help="route the stack is applied with")
    args = ap.parse_args()
    t0 = time.time()
    wl = workload.make(args.config, cell_order=args.order)
    rg = Regridder(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rg.use_torch_stream()
    workload.load_geometry(rg, wl)
    route = rg.store(L.BILINEAR if args.method == "bilinear" else L.NEAREST_STOD, L.SRC_MESH_ELEMENT, L.CENTER)
    epi = None
    info = route.info()
    nlev = args.nlev or wl.nz
    n = wl.mesh.nCells
    levs = [nlev] * args.fields
    if args.stack:
        levs = [int(a.split("x")[0]) for a in args.stack.split(",") for _ in range(int(a.split("x")[1]))]
    srcs = [torch.randn((n, L), device="cuda", dtype=torch.float32) for L in levs]
    dsts = [torch.empty((L, wl.n_mass), device="cuda", dtype=torch.float32) for L in levs]
    if args.rot:
        epi = [L.EPI_ROT_U, L.EPI_ROT_V] + [L.EPI_NONE] * (len(levs) - 2)
    import ctypes
    em, um = ctypes.c_int32(), ctypes.c_int32()
    rg.L.mprg_debug_route_tiles(ctypes.c_void_p(route.handle), ctypes.byref(em), ctypes.byref(um))
    print(f"setup {time.time() - t0:.1f}s  order {args.order} route {info} tile entries max {em.value} uniq max {um.value}", file=sys.stderr)
    peak = 6450.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for var in args.variants.split(","):
        # accumulate : launch grouping [b4|b5]   e.g. f32:split  f32:one  f64:split  f32:splitb4  f32:direct
        acc, mode = var.split(":")
        rg.set_option("accumulate", acc)
        mb = ""
        if len(mode) > 2 and mode[-2] == "b" and mode[-1].isdigit():
            mode, mb = mode[:-2], mode[-1]
        rg.set_option("pipe_minb", mb or "0")
        if mode == "direct":
            rg.set_option("apply", "direct")
        else:
            rg.set_option("apply", "pipe")
            rg.set_option("pipe_split", "1" if mode == "split" else "0")
        for _ in range(2):
            rg.apply(route, srcs, dsts, nlev=levs, epi_op=epi)
        rg.profile(True)
        for _ in range(args.iters):
            rg.apply(route, srcs, dsts, nlev=levs, epi_op=epi)
        recs = rg.profile_read()
        rg.profile(False)
        out = {}
        for r in recs:
            k = out.setdefault(KIND[r["kind"]], [0.0, 0.0, 0])
            k[0] += r["ms"]; k[1] += r["alg_bytes"]; k[2] += 1
        for k, (ms, by, cnt) in out.items():
            gbs = by / (ms * 1e-3) / 1e9
            print(f"{args.order:8s} {var:10s} {k:12s} {args.stack or f'{nlev}x{args.fields}':>14s}: {ms / cnt:8.3f} ms/launch  {gbs:8.1f} GB/s  "
                  f"{100 * gbs / peak:5.1f}% of {peak:.0f}")
    rg.close()


if __name__ == "__main__":
    main()
