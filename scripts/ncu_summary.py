"""Summarise an .ncu-rep (captured with `ncu --set full --import-source on`) into the handful of numbers the
roofline discussion needs + the top stall reasons per source line. Usage: ncu_summary.py <file.ncu-rep> [out.md]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor_op_gmma.sum", "lts__t_sector_hit_rate.pct"]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = run([rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    out = [f"# ncu summary: {rep}", ""]
    for v in vals:
        name = v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append(f"## {name[:140]}")
        out.append("| metric | value | unit |\n|---|---|---|")
        for i, h in enumerate(hdr):
            if any(h == k or h.startswith(k) for k in KEYS) or "tensor" in h and "pct" in h or h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("ratio"):
                out.append(f"| {h} | {v[i]} | {units[i]} |")
    src = run([rep, "--page", "source", "--csv"])
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        sh = srows[0]
        try:
            col_s = next(i for i, h in enumerate(sh) if h.strip() == "Source")
            col_samp = next(i for i, h in enumerate(sh) if "Warp Stall Sampling (All" in h or h.strip() == "# Samples")
        except StopIteration:
            col_s = col_samp = None
        if col_s is not None:
            agg = defaultdict(float)
            for r in srows[1:]:
                try:
                    agg[r[col_s].strip()[:110]] += float(r[col_samp] or 0)
                except (ValueError, IndexError):
                    pass
            tot = sum(agg.values()) or 1
            out.append("\n### hottest source lines (stall samples)\n| % | line |\n|---|---|")
            for line, s in sorted(agg.items(), key=lambda kv: -kv[1])[:25]:
                out.append(f"| {100 * s / tot:.1f} | `{line}` |")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
