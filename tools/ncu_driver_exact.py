#!/usr/bin/env python
"""Development aid: fused launches of the exact (FP64) pipeline on the full-size CAST+LLNL workload, for ncu."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_config
from solaraxionraytracing_b200 import raytracer as rt
cfg = sys.argv[1] if len(sys.argv) > 1 else "cast_llnl"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2 * 10**7
setup, tb = make_config(cfg, nR=1968, nE=1500, nAng=1000, nEn=1000)
with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
    tr.set_precision(0)
    tr.trace_mc(n, 299792458); tr.synchronize()
    tr.trace_mc(n, 299792458, first_ray=n); tr.synchronize()
