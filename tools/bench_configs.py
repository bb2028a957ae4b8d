#!/usr/bin/env python
"""Throughput of the other BASELINE configurations (bench.py measures the headline one). Development aid."""
import os, sys, time
if os.environ.get('SART_LIB_VARIANT'): os.environ['SART_LIB'] = os.environ['SART_LIB_VARIANT']
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from solaraxionraytracing_b200 import raytracer as rt, tables

def timeit(tr, n, reps=3):
    tr.trace_mc(n // 10, 1); tr.synchronize()
    best = 1e9
    for _ in range(reps):
        tr.reset_image(); tr.synchronize()
        t = time.perf_counter(); tr.trace_mc(n, 299792458); tr.synchronize(); best = min(best, time.perf_counter() - t)
    return best

def main():
    prec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    em_abc = tables.synthetic_emission(1968, 1500, "abc"); em_prim = tables.synthetic_emission(1968, 1500, "primakoff")
    only = [a for a in sys.argv[2:]]
    cfgs = [("1/3 CAST+LLNL (window+Ar chain)", ("CAST", "InGrid2018", "vacuum", "LLNL"), em_abc, 4, None, 1e9),
            ("2 CAST+XMM (all rays opaque, see DESIGN)", ("CAST", "InGrid2018", "vacuum", "XMM"), em_prim, 1, None, 1e8),
            ("4 BabyIAXO gas, 64 masses", ("BabyIAXO", "InGridIAXO", "gas", "XMM"), em_prim, 1, np.linspace(0.004, 0.012, 64), 1e8),
            ("5 BabyIAXO+XMM vacuum", ("BabyIAXO", "InGridIAXO", "vacuum", "XMM"), em_prim, 1, None, 1e9)]
    for name, setup, em, ncoat, masses, n in cfgs:
        if only and name.split()[0] not in only:
            continue
        fs = rt.initFullSetup(*setup, emission=em, reflectivity=tables.synthetic_reflectivity(ncoat, 1000, 1000))
        with rt.RayTracer(fs) as tr:
            tr.set_precision(prec)
            if masses is not None: tr.set_axion_masses(masses)
            dt = timeit(tr, int(n))
            c = tr.read_image().counters[0]
            m = 1 if masses is None else len(masses)
            print(f"config {name}: prec={prec} {n:.0e} rays x {m} masses in {dt*1e3:.1f} ms = {n/dt:.3e} rays/s"
                  f" ({n*m/dt:.3e} ray-masses/s) passed={c['n_passed']/c['n_rays']:.3f}", flush=True)

if __name__ == "__main__":
    main()
