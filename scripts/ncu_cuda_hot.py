"""Warp-stall samples aggregated per CUDA source line (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg = defaultdict(float); cur = None; hdr = None; fname = ""
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0] in ("Line No", "Address") or (len(r) > 1 and r[1] == "Source"):
        hdr = r; continue
    if hdr is None: continue
    if hdr[0] == "Line No" and r[0].isdigit() and len(r) >= 2 and "Warp Stall Sampling (All Samples)" in hdr:
        try: agg[(fname, int(r[0]), r[1].strip()[:100])] += float(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0)
        except ValueError: pass
T = sum(agg.values()) or 1
for (f, ln, src), s in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
    print(f"{100*s/T:5.1f}%  {f}:{ln:<5d} {src}")
