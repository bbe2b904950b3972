"""Where the end-to-end (host-buffer) pass spends its time on the C2 workload:
raw PCIe copies of one 3-D field, the main stacked apply with host buffers (weights memoised),
weight generation, and whole interp_data passes with weights memoised / rebuilt."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    wl = workload.make(a.config)
    rg = Regridder(0)
    workload.load_geometry(rg, wl)
    F = workload.make_fields(wl, device="cuda:0", pinned_host=True, rg=rg)
    h2d, d2h = workload.io_bytes(wl, F["host"])
    # ---- raw copies
    src = F["host"]["hist_3d"][2].src
    dst = F["host"]["hist_3d"][2].dst
    dsrc = torch.empty_like(src, device="cuda")
    ddst = torch.empty_like(dst, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (("H2D", lambda: dsrc.copy_(src, non_blocking=True)),
                     ("D2H", lambda: dst.copy_(ddst, non_blocking=True))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        nb = src.numel() * 4 if name == "H2D" else dst.numel() * 4
        print(f"raw {name}: {nb / 1e6:.0f} MB in {1e3 * dt:.2f} ms = {nb / dt / 1e9:.1f} GB/s")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            dsrc.copy_(src, non_blocking=True)
        with torch.cuda.stream(s2):
            dst.copy_(ddst, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"raw H2D+D2H concurrent: {1e3 * dt:.2f} ms per pair = {(src.numel() + dst.numel()) * 4 / dt / 1e9:.1f} GB/s total")
    print(f"pass moves H2D {h2d / 1e9:.2f} GB, D2H {d2h / 1e9:.2f} GB")

    # ---- whole passes
    def one(mem, rebuild):
        if rebuild:
            rg.clear_routes()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        workload.run_interp(rg, wl, F["host" if mem == L.HOST else "dev"], mem)
        rg.synchronize()
        return 1e3 * (time.perf_counter() - t0)

    for label, mem, rebuild in (("device buffers, weights memoised", L.DEVICE, False),
                                ("host buffers, weights memoised", L.HOST, False),
                                ("host buffers, weights rebuilt", L.HOST, True),
                                ("device buffers, weights rebuilt", L.DEVICE, True)):
        one(mem, rebuild)
        ts = [one(mem, rebuild) for _ in range(a.iters)]
        print(f"interp_data, {label}: " + " ".join(f"{t:.1f}" for t in ts) + " ms")
    rg.close()


if __name__ == "__main__":
    main()
