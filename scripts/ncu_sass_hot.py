"""Top SASS instructions by warp-stall samples + stall-reason totals from an .ncu-rep source page."""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
samp = ci["Warp Stall Sampling (All Samples)"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = defaultdict(float); lines = []
for r in rows[2:]:
    if len(r) <= samp: continue
    try: s = float(r[samp] or 0)
    except ValueError: continue
    lines.append((s, r[ci["Source"]].strip(), {h: float(r[ci[h]] or 0) for h in stalls}))
    for h in stalls: tot[h] += float(r[ci[h]] or 0)
T = sum(l[0] for l in lines) or 1
print(f"total samples {T:.0f}")
print("stall reason totals:", ", ".join(f"{h[6:]}={100*v/T:.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
for idx, (s, src, st) in enumerate(lines):
    pass
order = sorted(range(len(lines)), key=lambda i: -lines[i][0])[:topn]
for i in order:
    s, src, st = lines[i]
    top = max(st.items(), key=lambda kv: kv[1])
    print(f"{100*s/T:5.1f}%  #{i:5d}  {src[:95]:95s}  {top[0][6:]}")
