#!/usr/bin/env python
"""Static view of a kernel's SASS loops: for every backward branch, the opcode mix of the
loop body.  Usage: sass_loops.py lib.so kernel_substring [min_len]"""
import collections, re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
  name = f.split("\n", 1)[0]
  if pat not in name:
    continue
  ins = []
  for line in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
      ins.append((int(m.group(1), 16), m.group(2).strip()))
  print(name, len(ins), "instructions")
  addr2i = {a: i for i, (a, _) in enumerate(ins)}
  for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
    if m:
      tgt = int(m.group(1), 16)
      if tgt <= a and tgt in addr2i and i - addr2i[tgt] >= minlen:
        body = ins[addr2i[tgt]:i + 1]
        c = collections.Counter()
        for _, b in body:
          toks = b.split()
          op = toks[1] if toks[0].startswith("@") else toks[0]
          c[op.split(".")[0]] += 1
        print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr:", ", ".join(f"{k} {v}" for k, v in c.most_common(20)))
