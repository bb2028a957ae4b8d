#!/usr/bin/env python
"""Compaction on/off for the two telescope families (development aid)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from solaraxionraytracing_b200 import raytracer as rt
for name, args in (("cast_llnl", ("CAST", "InGrid2018", "vacuum", "LLNL")), ("babyiaxo_xmm", ("BabyIAXO", "InGridIAXO", "vacuum", "XMM"))):
    fs = rt.initFullSetup(*args)
    with rt.RayTracer(fs) as tr:
        tr.set_precision(1)
        res = {}
        for mode in (0, 1):
            tr.set_compaction(mode)
            tr.trace_mc(10_000_000, 1); tr.synchronize()
            best = 1e9
            for _ in range(3):
                tr.reset_image(); tr.synchronize()
                t = time.perf_counter(); tr.trace_mc(1_000_000_000, 299792458); tr.synchronize(); best = min(best, time.perf_counter() - t)
            r = tr.read_image()
            res[mode] = r
            print(f"{name} compaction={mode}: {1e9/best:.3e} rays/s  passed={r.counters[0]['n_passed']} sum_w={r.counters[0]['sum_w']:.9e}", flush=True)
        assert res[0].counters[0]["n_exit"] == res[1].counters[0]["n_exit"]
        print("   identical exit counters; image rel diff", abs(res[0].image - res[1].image).sum() / res[0].image.sum())
