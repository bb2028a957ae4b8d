#!/usr/bin/env python
"""SASS of one source region with executed counts per 32 rays: ncu_region_sass.py src.csv nrays file lo hi"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
nrays = float(sys.argv[2]); fname = sys.argv[3]; lo = int(sys.argv[4]); hi = int(sys.argv[5])
cur = None; hdr = None; line = None; out = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; iinst = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < 10: continue
    if r[2] == "-":
        try: line = int(r[0])
        except ValueError: line = None
        continue
    if r[2].startswith("0x") and cur == fname and line is not None and lo <= line <= hi:
        out[r[2]] = (line, r[3].strip(), int(r[iinst] or 0))
for a in sorted(out):
    l, s, n = out[a]
    print(f"{n/(nrays/32):6.2f}  L{l:<4d} {s}")
