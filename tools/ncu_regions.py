#!/usr/bin/env python
"""Warp-instructions per 32 rays by source region, from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`.
A SASS instruction of inlined code is listed once per level of its inline stack; here it is counted once, under the
INNERMOST... no: under every (file, line) it is listed at, and the regions below are ranges of trace_f32.cuh lines
(the outermost inline level of the per-ray code), so each instruction lands in exactly one region.
usage: ncu_regions.py src.csv nrays 'file:name:lo-hi,...' (rules in priority order)"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
nrays = float(sys.argv[2])
# rules in priority order: file:name:lo-hi
regions = [(r.split(":")[0], r.split(":")[1], int(r.split(":")[2].split("-")[0]), int(r.split(":")[2].split("-")[1])) for r in sys.argv[3].split(",")]
cur = None; hdr = None; line = None
seen = {}       # address -> (inst, opcode)
inreg = {}      # address -> region
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; iinst = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < 10: continue
    if r[2] == "-":
        try: line = int(r[0])
        except ValueError: line = None
        continue
    if not r[2].startswith("0x"): continue
    addr = r[2]
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3])
    try: inst = int(r[iinst] or 0)
    except ValueError: continue
    seen[addr] = (inst, m.group(2) if m else "?")
    if line is not None:
        for pri, (f, name, lo, hi) in enumerate(regions):
            if f == cur and lo <= line <= hi:
                if addr not in inreg or pri < inreg[addr][0]: inreg[addr] = (pri, name)
                break
witer = nrays / 32
tot = sum(v[0] for v in seen.values())
print(f"total {tot/witer:.1f} warp-instructions per 32 rays ({len(seen)} SASS instructions)")
by = collections.defaultdict(lambda: [0, collections.Counter()])
for a, (inst, op) in seen.items():
    k = inreg.get(a, (99, "(other)"))[1]
    by[k][0] += inst; by[k][1][op] += inst
for k, (n, ops) in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print(f"{n/witer:7.1f}  {k:14s} " + ", ".join(f"{o} {c/witer:.0f}" for o, c in ops.most_common(12)))
