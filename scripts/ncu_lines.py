#!/usr/bin/env python
"""Top SASS lines by samples of one launch in an .ncu-rep, with their main stall reasons.
Usage: ncu_lines.py rep [launch_idx] [topN] [lo hi]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{launch + 1}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
si, ii, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = {x: hdr.index(x) for x in hdr if x.startswith("stall_") and "Not Issued" not in x}
tots = sum(int(r[smp]) for r in data)
sel = [(k, r) for k, r in enumerate(data) if lo <= k <= hi]
if len(sys.argv) > 4:
  for k, r in sel:
    st = sorted(((int(r[c]), x[6:]) for x, c in stalls.items()), reverse=True)[:2]
    print(f"{k:5d} {int(r[smp])/tots*100:5.2f}% {int(r[ii])/1e6:8.1f}M  {r[si].strip()[:70]:70s} " + " ".join(f"{n}:{c}" for c, n in st if c))
else:
  for k, r in sorted(sel, key=lambda kr: -int(kr[1][smp]))[:top]:
    st = sorted(((int(r[c]), x[6:]) for x, c in stalls.items()), reverse=True)[:3]
    print(f"{k:5d} {int(r[smp])/tots*100:5.2f}% {int(r[ii])/1e6:8.1f}M  {r[si].strip()[:70]:70s} " + " ".join(f"{n}:{c}" for c, n in st if c))
if len(sys.argv) > 4:
  import collections
  c = collections.Counter()
  for k, r in sel:
    for x, ci in stalls.items():
      c[x[6:]] += int(r[ci])
  tot = sum(c.values())
  print("region stall mix:", ", ".join(f"{k} {v/tot*100:.1f}%" for k, v in c.most_common(10)), f"(region = {tot/sum(int(r[ci]) for r in data for ci in stalls.values())*100:.1f}% of all)")
