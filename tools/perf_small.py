#!/usr/bin/env python
"""How much of the fused kernel's time is table traffic: the same setups with tables small enough to live in L1
(246 x 300 solar model, 200 x 200 reflectivity) against the full-size ones. Development aid."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from solaraxionraytracing_b200 import raytracer as rt, tables
precs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2]
for label, (nR, nE, nA, nEn) in (("small tables", (246, 300, 200, 200)), ("full tables ", (1968, 1500, 1000, 1000))):
    for name, args, kind, ncoat in (("cast_llnl   ", ("CAST", "InGrid2018", "vacuum", "LLNL"), "abc", 4),
                                    ("babyiaxo_xmm", ("BabyIAXO", "InGridIAXO", "vacuum", "XMM"), "primakoff", 1)):
        fs = rt.initFullSetup(*args, emission=tables.synthetic_emission(nR, nE, kind),
                              reflectivity=tables.synthetic_reflectivity(ncoat, nA, nEn))
        with rt.RayTracer(fs) as tr:
            for prec in precs:
                tr.set_precision(prec)
                n = 10**9
                tr.trace_mc(n // 10, 1); tr.synchronize()
                best = 1e9
                for _ in range(3):
                    tr.reset_image(); tr.synchronize()
                    t = time.perf_counter(); tr.trace_mc(n, 299792458); tr.synchronize(); best = min(best, time.perf_counter() - t)
                print(f"{label} {name} prec {prec}: {n/best:.4e} rays/s ({best*1e3:.2f} ms)", flush=True)
