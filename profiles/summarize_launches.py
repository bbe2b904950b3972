"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py:
per-kernel totals over the whole run and the launch-by-launch breakdown of the last step.
(ncu times are cold-cache and serialised: compare SHARES with the CUDA-event numbers, not absolutes.)

    python profiles/summarize_launches.py gpurun_out/launches_bench.csv > profiles/r01/launches_bench.md
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.split("(")[0]
    name = re.sub(r"<.*", "", name) if name.startswith("cub::") or "Device" in name else name
    return name.replace("mprg::", "")


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        seq.append((short(row["Kernel Name"]), ms, row["Grid Size"]))
    pipes = [i for i, s in enumerate(seq) if s[0].startswith("k_apply_pipe")]
    print(f"# ncu launch list: {len(seq)} launches of engine + CUB kernels\n")
    agg = collections.OrderedDict()
    for n, ms, _ in seq:
        a = agg.setdefault(n.split("<")[0], [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print("## whole run (setup: BVH builds + 5 weight generations; then warm-up and timed interp_data passes)\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t:.3f} | {100 * t / tot:.1f} % |")
    if len(pipes) >= 4:
        # the pass starts with the column launch of the bilinear apply (the one with the largest duration);
        # its period = distance between the last two such launches
        big = max(seq[i][1] for i in pipes)
        heads = [i for i in pipes if seq[i][1] > 0.5 * big]
        per = heads[-1] - heads[-2]
        start = len(seq) - per
        step = seq[start:]
        st = sum(ms for _, ms, _ in step)
        print(f"\n## last interp_data pass ({per} launches, {st:.3f} ms under ncu)\n")
        print("| # | kernel | grid | ms | share of step |\n|---|---|---|---|---|")
        for k, (n, ms, g) in enumerate(step):
            print(f"| {k} | {n[:70]} | {g} | {ms:.4f} | {100 * ms / st:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1])
