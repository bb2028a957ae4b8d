#!/usr/bin/env python
"""End-to-end time of one calculateFluxFractions-sized call (reset + trace + read image back) versus the ray count:
BASELINE configs[0] is the reference's own 1e6-ray run. Development aid."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from solaraxionraytracing_b200 import raytracer as rt, tables
fs = rt.initFullSetup("CAST", "InGrid2018", "vacuum", "LLNL", emission=tables.synthetic_emission(1968, 1500, "abc"),
                      reflectivity=tables.synthetic_reflectivity(4, 1000, 1000))
with rt.RayTracer(fs) as tr:
    tr.set_precision(2)
    for n in (10**4, 10**5, 10**6, 10**7, 10**8):
        for _ in range(3):
            tr.reset_image(); tr.trace_mc(n, 1); tr.read_image()
        reps = 20
        t = time.perf_counter()
        for k in range(reps):
            tr.reset_image(); tr.trace_mc(n, 1, first_ray=k * n); res = tr.read_image()
        dt = (time.perf_counter() - t) / reps
        t = time.perf_counter()
        for k in range(reps):
            tr.reset_image(); tr.trace_mc(n, 1, first_ray=k * n)
        tr.synchronize()
        dk = (time.perf_counter() - t) / reps
        print(f"n={n:.0e}: e2e {dt*1e6:9.1f} us/call = {n/dt:.3e} rays/s; trace only {dk*1e6:9.1f} us = {n/dk:.3e} rays/s", flush=True)
