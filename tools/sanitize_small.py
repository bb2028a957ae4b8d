#!/usr/bin/env python
"""Small runs of every kernel family, one after the other: a crash / illegal-address canary for every launch path (compute-sanitizer is closed on this pool)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_config
from oracle import oracle as orc
from solaraxionraytracing_b200 import raytracer as rt, tables

for cfg in ("cast_llnl", "babyiaxo_xmm", "babyiaxo_gas", "cast_abrixas"):
    setup, tb = make_config(cfg)
    origin, exit_xy, energy = orc.sample_rays(setup, tb, 0, 3000, 1)
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        for mode in (0, 1, 2):
            tr.set_precision(mode)
            for comp in (0, 1):
                tr.set_compaction(comp)
                tr.reset_image(); tr.trace_mc(20_001, 5, first_ray=2**32 - 777); tr.read_image()
            tr.traceAxionWrapper(5000, 5)
            tr.trace_presampled(origin, exit_xy, energy)
            tr.angular_scan(np.array([0.0, 0.03]), 4097, 5)
        tr.set_axion_masses(np.linspace(0.004, 0.02, 37))
        for mode in (0, 1, 2):
            tr.set_precision(mode); tr.reset_image(); tr.trace_mc(9001, 5); tr.read_image()
        tr.set_axion_masses([0.0853]); tr.enable_radial_hist(1000); tr.set_precision(2); tr.reset_image(); tr.trace_mc(9001, 5)
        tr.read_radial_hist()
    print(cfg, "ok", flush=True)
em = rt.calculateEmissionRates(tables.SolarModel(*[a[:64] for a in (lambda s: (s.radius, s.temp_K, s.rho_gcm3, s.mass_fractions))(tables.solar_model_packaged())]),
                               tuple(__import__("solaraxionraytracing_b200.abi", fromlist=["x"]).EM_PROCESSES), nElems=50)
rt.buildCdfs(em)
print("emission ok")
