"""A/B of two builds of libmpassit_rg.so on the same stacked apply, in one process and one GPU session
(profiling aid).  Binds only the entry points both builds share.

  python profiles/ab_lib.py [--libs new=mpassit_b200/libmpassit_rg.so,r01=profiles/_ab/libmpassit_rg_r01.so]
                            [--stack 60x12] [--rot] [--order rowmajor] [--iters 8] [--only NAME]

The round-1 library is rebuilt from history with:  git show cf973de:mpassit_b200/csrc/<file> ...; nvcc (see
profiles/r02/README.md).
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mpassit_b200 import workload  # noqa: E402


def bind(path):
    L = C.CDLL(os.path.join(ROOT, path))
    vp, i32 = C.c_void_p, C.c_int32
    L.mprg_last_error.restype = C.c_char_p
    L.mprg_last_error.argtypes = [vp]
    L.mprg_init.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.mprg_set_stream.argtypes = [vp, vp]
    L.mprg_set_mesh.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.mprg_set_target.argtypes = [vp, C.c_int, i32, i32, vp, vp]
    L.mprg_set_rotation.argtypes = [vp, vp, vp]
    L.mprg_store.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.mprg_apply_ex.argtypes = [vp, vp, i32, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp]
    L.mprg_profile_enable.argtypes = [vp, C.c_int]
    L.mprg_profile_reset.argtypes = [vp]
    L.mprg_profile_read.argtypes = [vp, i32, vp, vp, vp, vp]
    L.mprg_synchronize.argtypes = [vp]
    L.mprg_finalize.argtypes = [vp]
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default="new=mpassit_b200/libmpassit_rg.so,r01=profiles/_ab/libmpassit_rg_r01.so")
    ap.add_argument("--stack", default="60x12")
    ap.add_argument("--rot", action="store_true")
    ap.add_argument("--order", default="rowmajor")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--only", default="")
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--planes", action="store_true", help="A/B the grid-source (centre -> EDGE1) apply instead of the column apply")
    args = ap.parse_args()
    wl = workload.make("c2", cell_order=args.order)
    m = wl.mesh
    levs = [int(a.split("x")[0]) for a in args.stack.split(",") for _ in range(int(a.split("x")[1]))]
    n = m.nCells
    if args.planes:
        srcs = [torch.randn((L_, wl.n_mass), device="cuda") for L_ in levs]
        dsts = [torch.empty((L_, wl.grids["U"][0].size), device="cuda") for L_ in levs]
    else:
        srcs = [torch.randn((n, L_), device="cuda") for L_ in levs]
        dsts = [torch.empty((L_, wl.n_mass), device="cuda") for L_ in levs]
    torch.cuda.synchronize()
    peak = 6450.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    libs = [kv.split("=") for kv in args.libs.split(",")]
    ctxs = {}
    for name, path in libs:
        if args.only and name != args.only:
            continue
        L = bind(path)
        ctx = C.c_void_p()
        assert L.mprg_init(0, 0, 1, C.byref(ctx)) == 0

        def ck(rc, L=L, ctx=ctx):
            if rc:
                raise RuntimeError(L.mprg_last_error(ctx).decode())

        voc = np.ascontiguousarray(m.verticesOnCell, np.int32)
        ck(L.mprg_set_mesh(ctx, m.nCells, m.nVertices, voc.shape[1], m.lonCell.ctypes.data, m.latCell.ctypes.data,
                           m.lonVertex.ctypes.data, m.latVertex.ctypes.data, voc.ctypes.data))
        lat, lon = wl.grids["M"]
        lon, lat = np.ascontiguousarray(lon), np.ascontiguousarray(lat)
        ck(L.mprg_set_target(ctx, 0, lon.shape[1], lon.shape[0], lon.ctypes.data, lat.ctypes.data))
        ca, sa = np.ascontiguousarray(wl.cosa), np.ascontiguousarray(wl.sina)
        ck(L.mprg_set_rotation(ctx, ca.ctypes.data, sa.ctypes.data))
        rh = C.c_void_p()
        if args.planes:
            ulat, ulon = wl.grids["U"]
            ulon, ulat = np.ascontiguousarray(ulon), np.ascontiguousarray(ulat)
            ck(L.mprg_set_target(ctx, 1, ulon.shape[1], ulon.shape[0], ulon.ctypes.data, ulat.ctypes.data))
            ck(L.mprg_store(ctx, 0, 2, 1, C.byref(rh)))      # BILINEAR, SRC_GRID_CENTER, EDGE1
        else:
            ck(L.mprg_store(ctx, 0, 0, 0, C.byref(rh)))
        ctxs[name] = (L, ctx, rh, ck)
    k = len(levs)
    sp = (C.c_void_p * k)(*[t.data_ptr() for t in srcs])
    dp = (C.c_void_p * k)(*[t.data_ptr() for t in dsts])
    nl = (C.c_int32 * k)(*levs)
    eo = (C.c_int32 * k)(*([3, 4] + [0] * (k - 2) if args.rot else [0] * k))
    ea = (C.c_double * k)(*[0.0] * k)
    ref = None
    for rnd in range(args.rounds):
        for name, (L, ctx, rh, ck) in ctxs.items():
            for _ in range(2):
                ck(L.mprg_apply_ex(ctx, rh, k, sp, nl, 0, 1, dp, 0, 1, eo, ea))
            ck(L.mprg_synchronize(ctx))
            L.mprg_profile_enable(ctx, 1)
            L.mprg_profile_reset(ctx)
            for _ in range(args.iters):
                ck(L.mprg_apply_ex(ctx, rh, k, sp, nl, 0, 1, dp, 0, 1, eo, ea))
            cap = 64 * args.iters
            kind, ms = (C.c_int32 * cap)(), (C.c_double * cap)()
            ab, un = (C.c_double * cap)(), (C.c_double * cap)()
            nrec = L.mprg_profile_read(ctx, cap, kind, ms, ab, un)
            L.mprg_profile_enable(ctx, 0)
            tms = sum(ms[i] for i in range(nrec)) / args.iters
            tb = sum(ab[i] for i in range(nrec)) / args.iters
            gbs = tb / (tms * 1e-3) / 1e9
            out = dsts[0].clone()
            same = ""
            if ref is None:
                ref = out
            else:
                same = "  bit-identical to the first build" if torch.equal(ref, out) else "  DIFFERS from the first build"
            print(f"round {rnd} {name:6s} {args.stack:>14s}{' rot' if args.rot else ''}: {tms:8.3f} ms/apply ({nrec // args.iters} launches)  "
                  f"{gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of {peak:.0f}{same}", flush=True)
    for name, (L, ctx, rh, ck) in ctxs.items():
        L.mprg_finalize(ctx)


if __name__ == "__main__":
    main()
