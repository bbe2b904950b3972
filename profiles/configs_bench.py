"""One number per BASELINE.json config (the table of BASELINE.md section 5); bench.py itself measures configs[1].

  python profiles/configs_bench.py [--configs c1,c3,c4,c5] [--iters 10] [--cpu]

c1  whole interp_data pass, device resident (+ the CPU oracle on 1 thread: "1 MPI rank on CPU")
c3  every histlist_3d + histlist_soil field in ONE stacked apply on the bilinear route (914 level-columns)
c4  6.5 M-cell variable-resolution mesh -> 0.03 degree: weights (bilinear + nearest) and apply (55 levels + one 2-D field)
c5  1-km conservative: weights + apply of snow / snowh, weights rebuilt every iteration
Prints one JSON object per config.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mpassit_b200 import lib as L  # noqa: E402
from mpassit_b200 import workload  # noqa: E402
from mpassit_b200.regrid import Regridder  # noqa: E402

PEAK = 6450.0
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def prof(rg, fn, iters):
    for _ in range(2):
        fn()
    rg.profile(True)
    for _ in range(iters):
        fn()
    recs = rg.profile_read()
    rg.profile(False)
    by = sum(r["alg_bytes"] for r in recs) / iters
    ms = sum(r["ms"] for r in recs) / iters
    return ms, by


def new_rg():
    rg = Regridder(0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    rg.use_torch_stream()
    return rg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c3,c5,c4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle")
    a = ap.parse_args()
    for name in a.configs.split(","):
        out = {"config": name}
        t0 = time.time()
        wl = workload.make(name)
        rg = new_rg()
        workload.load_geometry(rg, wl)
        out["setup_s"] = round(time.time() - t0, 1)
        if name == "c1":
            F = workload.make_fields(wl, device="cuda:0")
            step = workload.prepare_interp(rg, wl, F["dev"], L.DEVICE)
            step()
            ms = timed(step, a.iters)
            kms, by = prof(rg, step, a.iters)
            units = wl.units_per_pass()
            out.update(phase="interp_data pass (weights memoised)", units=units, ms=ms, value=units / (ms * 1e-3),
                       kernels_ms=kms, GBps=by / (kms * 1e-3) / 1e9, frac=by / (kms * 1e-3) / 1e9 / PEAK,
                       note="64,800-point target: the pass is a handful of launches of ~10 us each, launch-bound, not bandwidth-bound")
            if a.cpu:
                from bench import cpu_reference_pass
                from oracle import oracle as orc
                orc.build()
                fields = {g: [(s.name, s.src.cpu().numpy()) for s in F["dev"][g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
                fields["ter"] = F["dev"]["ter"].cpu().numpy()
                for thr in (1, 0):
                    from bench import host_threads
                    t = thr or host_threads()
                    u, times, det, _ = cpu_reference_pass(wl, fields, 1.0, 2, 1, t)
                    out[f"cpu_{t}_threads"] = {"value": u / min(times), "weights_s": det["weights_s"], "apply_s": det["apply_s"]}
        elif name == "c3":
            names, srcs, dsts = workload.make_stacked(wl, "cuda:0")
            levs = [int(t.shape[1]) for t in srcs]
            r = rg.store(L.BILINEAR, L.SRC_MESH_ELEMENT, L.CENTER)
            fn = lambda: rg.apply(r, srcs, dsts, nlev=levs)  # noqa: E731
            ms = timed(fn, a.iters)
            kms, by = prof(rg, fn, a.iters)
            units = sum(levs) * wl.n_mass
            out.update(phase="one stacked apply, K = %d level-columns" % sum(levs), units=units, ms=ms, value=units / (ms * 1e-3),
                       kernels_ms=kms, GBps=by / (kms * 1e-3) / 1e9, frac=by / (kms * 1e-3) / 1e9 / PEAK)
        elif name == "c4":
            nlev = wl.nz
            n, nd = wl.mesh.nCells, wl.n_mass
            g = torch.Generator(device="cuda")
            g.manual_seed(4)
            src = (280.0 + 20.0 * torch.randn((n, nlev), generator=g, device="cuda")).contiguous()
            veg = (torch.arange(n, device="cuda") * 2654435761 % 20 + 1).to(torch.float32).reshape(-1, 1).contiguous()
            dst = torch.empty((nlev, nd), device="cuda")
            ov = torch.empty((1, nd), device="cuda")
            st = {}
            for tag, m in (("bilinear", L.BILINEAR), ("nearest", L.NEAREST_STOD)):
                ts = []
                for _ in range(3):
                    rg.clear_routes()
                    t1 = time.perf_counter()
                    rr = rg.store(m, L.SRC_MESH_ELEMENT, L.CENTER)
                    rg.synchronize()
                    ts.append(1e3 * (time.perf_counter() - t1))
                    rr.release()
                st[tag] = min(ts)
            rb = rg.store(L.BILINEAR, L.SRC_MESH_ELEMENT, L.CENTER)
            rn = rg.store(L.NEAREST_STOD, L.SRC_MESH_ELEMENT, L.CENTER)

            def fn():
                rg.apply(rb, [src], [dst], nlev=[nlev])
                rg.apply(rn, [veg], [ov], nlev=[1])
            ms = timed(fn, a.iters)
            kms, by = prof(rg, fn, a.iters)
            units = (nlev + 1) * nd
            out.update(phase="weights (bilinear + nearest, 72 M target points) / apply (55-level field + integer field)", units=units,
                       store_ms=st, store_points_per_s={k: nd / (v * 1e-3) for k, v in st.items()}, ms=ms, value=units / (ms * 1e-3),
                       kernels_ms=kms, GBps=by / (kms * 1e-3) / 1e9, frac=by / (kms * 1e-3) / 1e9 / PEAK)
        elif name == "c5":
            m = wl.mesh
            from mpassit_b200 import synth
            snow = torch.from_numpy(synth.patchy_field(m.lonCell, m.latCell)).cuda()
            snowh = (0.01 * snow).contiguous()
            o1, o2 = torch.empty((1, wl.n_mass), device="cuda"), torch.empty((1, wl.n_mass), device="cuda")

            def fn():
                rg.clear_routes()
                r = rg.store(L.CONSERVE, L.SRC_MESH_ELEMENT, L.CENTER)
                rg.apply(r, [snow, snowh], [o1, o2], nlev=[1, 1])
                r.release()
            ts = []
            for _ in range(3 + a.iters):
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                fn()
                rg.synchronize()
                ts.append(1e3 * (time.perf_counter() - t1))
            ms = float(np.median(ts[3:]))
            units = 2 * wl.n_mass
            out.update(phase="weights + apply, rebuilt every step (snow, snowh)", units=units, ms=ms, value=units / (ms * 1e-3),
                       store_points_per_s=wl.n_mass / (ms * 1e-3))
            if a.cpu:
                from oracle import interp_oracle
                from oracle import oracle as orc
                orc.build()
                t1 = time.perf_counter()
                cxyz, vxyz, _ = interp_oracle.geometry(m)
                clat, clon = wl.grids["CORNER"]
                cor = orc.sph_deg_to_cart(clon, clat).reshape(clat.shape[0], clat.shape[1], 3)
                rp, cc, ww = orc.conserve(cxyz, vxyz, m.verticesOnCell, cor)
                for f in (snow, snowh):
                    orc.apply(rp, cc, ww, f.cpu().numpy(), np.float32)
                dt = time.perf_counter() - t1
                out["cpu"] = {"threads": orc.num_threads(), "s": dt, "value": units / dt}
        rg.close()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
