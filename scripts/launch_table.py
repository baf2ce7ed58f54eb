"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import collections
import csv
import re
import sys


def table(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        name, v, unit = row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        v = v / 1e3 if unit == "us" else v / 1e6 if unit == "ns" else v * 1e3 if unit == "s" else v
        m = re.search(r"gemm_kernel<(\d+), (\d+), (\d+)(?:, (\d+))?>", name)
        if m:
            k = f"gemm_tcgen05<A_MN={m.group(1)},B_MN={m.group(2)},EPI={m.group(3)},CL={m.group(4)}>"
        else:
            m = re.search(r"(\w+_kernel)(<\d+>)?", name)
            k = (m.group(1) + (m.group(2) or "")) if m else name[:60]
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    out = [f"{'ms':>9} {'%':>6} {'n':>5} {'avg us':>9}  kernel"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{t:9.3f} {100 * t / tot:6.1f} {n:5d} {t / n * 1e3:9.1f}  {k}")
    out.append(f"{tot:9.3f}  total device time of the profiled step")
    return "\n".join(out)


if __name__ == "__main__":
    print(table(sys.argv[1]))
