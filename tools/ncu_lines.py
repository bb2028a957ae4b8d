#!/usr/bin/env python
"""Per-source-line instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
nrays = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9
hdr = None; cur = None
lines = collections.OrderedDict(); ops = collections.Counter(); total = 0
seen_sass = False
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    iln, isrc, iaddr, isass = 0, 1, 2, 3
    iinst = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
    try: inst = int(r[iinst] or 0); samp = int(r[isamp] or 0)
    except ValueError: continue
    if r[iaddr] == "-":   # a CUDA source line (aggregate of its SASS)
        key = (cur.split("/")[-1] if cur else "?", r[iln])
        if key not in lines: lines[key] = [0, 0, r[isrc].strip()]
        lines[key][0] += inst; lines[key][1] += samp
    elif r[iaddr].startswith("0x"):
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[isass])
        if m: ops[m.group(2)] += inst
        total += inst
witer = nrays / 32
print(f"total SASS warp-instructions {total:.4g} = {total/witer:.0f} per warp-iteration")
print("opcodes per warp-iteration:", ", ".join(f"{k} {v/witer:.0f}" for k, v in ops.most_common(30)))
tot_lines = sum(v[0] for v in lines.values())
print(f"\nhot source lines (warp-instr per warp-iteration, samples%) of {tot_lines/witer:.0f}:")
tot_s = sum(v[1] for v in lines.values()) or 1
for (f, ln), (inst, samp, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print(f"{inst/witer:7.1f} {100*samp/tot_s:5.1f}%  {f}:{ln}  {src[:110]}")
