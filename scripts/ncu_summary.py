#!/usr/bin/env python
"""Summarise an .ncu-rep: headline raw metrics per launch + SASS opcode mix / stall reasons /
phase split (regions between BAR.SYNC) of one launch.  Usage: ncu_summary.py rep [launch_idx]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
for w in want:
  if w in hdr:
    i = hdr.index(w)
    print(f"{w:70s} {rows[1][i]:12s}", [r[i][:40] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{launch + 1}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h]) and r[rows[h].index('# Samples')].isdigit()]
si, ii, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ii]) for r in data)
tots = sum(int(r[smp]) for r in data)
print(f"\nlaunch {launch}: {tot/1e9:.2f} G warp-instr, {len(data)} SASS lines")
by, bys = collections.Counter(), collections.Counter()
for r in data:
  toks = r[si].strip().split()
  op = toks[1] if toks[0].startswith("@") else toks[0]
  op = op.split(".")[0]
  by[op] += int(r[ii])
  bys[op] += int(r[smp])
print("opcode mix:", ", ".join(f"{op} {c/tot*100:.1f}%" for op, c in by.most_common(16)))
stalls = {x: hdr.index(x) for x in hdr if x.startswith("stall_") and "Not Issued" not in x}
sc = {x: sum(int(r[c]) for r in data) for x, c in stalls.items()}
print("stalls:", ", ".join(f"{x[6:]} {c/tots*100:.1f}%" for x, c in sorted(sc.items(), key=lambda kv: -kv[1])[:8]))
regions, cur = [], []
for k, r in enumerate(data):
  cur.append((k, r))
  if "BAR.SYNC" in r[si]:
    regions.append(cur)
    cur = []
regions.append(cur)
for reg in regions:
  s = sum(int(r[smp]) for _, r in reg)
  ins = sum(int(r[ii]) for _, r in reg)
  print(f"  sass[{reg[0][0]:4d}..{reg[-1][0]:4d}] samples {s/tots*100:5.1f}%  instr {ins/1e9:6.2f} G")
