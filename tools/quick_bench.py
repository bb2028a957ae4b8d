#!/usr/bin/env python
"""Quick throughput probe of the fused MC kernel (development aid; bench.py is the contract)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from solaraxionraytracing_b200 import raytracer as rt, abi

def run(name, exp, det, stage, tel, nrays, precision=0, masses=None):
    t0 = time.time()
    fs = rt.initFullSetup(exp, det, stage, tel)
    t1 = time.time()
    with rt.RayTracer(fs) as tr:
        if precision: tr.set_precision(precision)
        if masses is not None: tr.set_axion_masses(masses)
        tr.trace_mc(1_000_000, 1); tr.synchronize()
        for n in nrays:
            tr.reset_image(); tr.synchronize()
            t2 = time.time()
            tr.trace_mc(n, 299792458); tr.synchronize()
            dt = time.time() - t2
            res = tr.read_image()
            c = res.counters[0]
            print(f"{name} prec={precision} n={n:.1e} {dt*1e3:9.2f} ms  {n/dt:.3e} rays/s  passed={c['n_passed']/n:.3f} "
                  f"sum_w={c['sum_w']:.4e} setup={t1-t0:.1f}s", flush=True)

if __name__ == "__main__":
    prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    run("cast_llnl", "CAST", "InGrid2018", "vacuum", "LLNL", [10_000_000, 100_000_000], prec)
    run("babyiaxo_xmm", "BabyIAXO", "InGridIAXO", "vacuum", "XMM", [10_000_000, 100_000_000], prec)
