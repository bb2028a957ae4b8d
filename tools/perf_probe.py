#!/usr/bin/env python
"""Throughput + result checksum of the fused fast kernel on the two telescope families at full table sizes
(development aid; bench.py is the contract). Prints exit-counter checksums so that an optimisation can be checked for
being result-neutral against an earlier build:  python tools/perf_probe.py [libsart variant .so] [precision]"""
import hashlib, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
args = sys.argv[1:]
lib = [a for a in args if a.endswith(".so")]
if lib:
    os.environ["SART_LIB"] = lib[0]
precs = [int(a) for a in args if a.isdigit()] or [1]
alias = "alias" in args      # sart_set_sampler(SART_SAMPLER_ALIAS): precision 2 only
from solaraxionraytracing_b200 import raytracer as rt, tables

def probe(name, args, em, ncoat, n):
    fs = rt.initFullSetup(*args, emission=em, reflectivity=tables.synthetic_reflectivity(ncoat, 1000, 1000))
    with rt.RayTracer(fs) as tr:
      for prec in precs:
        tr.set_precision(prec)
        tr.set_sampler(1 if alias and prec == 2 else 0)
        tr.trace_mc(n // 10, 1); tr.synchronize()
        best = 1e9
        for _ in range(3):
            tr.reset_image(); tr.synchronize()
            t = time.perf_counter(); tr.trace_mc(n, 299792458); tr.synchronize(); best = min(best, time.perf_counter() - t)
        r = tr.read_image()
        c = r.counters[0]
        h = hashlib.sha1(json.dumps(c["n_exit"], sort_keys=True).encode()).hexdigest()[:10]
        if len(precs) > 1:
            print("   ", {k: v for k, v in c["n_exit"].items() if v})
        print(f"{name} prec {prec}{' alias' if alias and prec == 2 else ''}: {n/best:.4e} rays/s ({best*1e3:.2f} ms)  exit-hash {h} passed {c['n_passed']} till_window {c['n_passed_till_window']} "
              f"sum_w {c['sum_w']:.12e} img {r.image.sum():.12e}", flush=True)

em_abc = tables.synthetic_emission(1968, 1500, "abc"); em_prim = tables.synthetic_emission(1968, 1500, "primakoff")
probe("cast_llnl   ", ("CAST", "InGrid2018", "vacuum", "LLNL"), em_abc, 4, 10**9)
probe("babyiaxo_xmm", ("BabyIAXO", "InGridIAXO", "vacuum", "XMM"), em_prim, 1, 10**9)
