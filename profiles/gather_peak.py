"""What HBM delivers for the access pattern of a mesh numbered WITHOUT locality (profiling aid).

The column kernel reads ~64 source columns of 240 bytes per 32-target tile.  With a spatially coherent numbering the
columns of a tile are neighbours in memory (a few long runs per tile); with a random numbering every column sits in
its own DRAM page.  This script measures the plain device ceiling for both patterns with a library gather
(torch.index_select over rows of 60 floats, every row read once and written once), so that the kernel's figure on a
random-numbered mesh can be read against what the memory system can do for that pattern, not against the streaming
copy peak of MEASURED_PEAKS.json.

  python profiles/gather_peak.py [--rows 2447080] [--nlev 60]
"""
import argparse
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2447080)
    ap.add_argument("--nlev", type=int, default=60)
    ap.add_argument("--fields", type=int, default=12)
    args = ap.parse_args()
    peak = 6450.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n, L, F = args.rows, args.nlev, args.fields
    src = torch.randn((F, n, L), device="cuda")
    dst = torch.empty_like(src)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    ident = torch.arange(n, device="cuda")
    rand = torch.randperm(n, device="cuda", generator=g)
    # locality at the scale of a tile: blocks of 20 consecutive rows in random order (what a row-major / Z-order
    # numbering gives the kernel: runs of ~20 columns)
    blocks = torch.randperm(n // 20, device="cuda", generator=g)
    runs20 = (blocks[:, None] * 20 + torch.arange(20, device="cuda")[None, :]).reshape(-1)
    byts = 2.0 * F * n * L * 4
    out = {}
    for name, idx in (("copy (streaming)", None), ("gather, identity order", ident), ("gather, runs of 20 rows", runs20),
                      ("gather, random rows", rand)):
        if idx is None:
            ms = timed(lambda: dst.copy_(src))
            b = byts
        else:
            m = idx.numel()

            def f(idx=idx):
                for k in range(F):
                    torch.index_select(src[k], 0, idx, out=dst[k, :idx.numel()])
            ms = timed(f)
            b = 2.0 * F * m * L * 4
        gbs = b / (ms * 1e-3) / 1e9
        out[name] = gbs
        print(f"{name:28s} rows of {L * 4} B: {ms:7.3f} ms  {gbs:7.1f} GB/s  {100 * gbs / peak:5.1f}% of {peak:.0f}")


if __name__ == "__main__":
    main()
