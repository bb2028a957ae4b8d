#!/usr/bin/env python
"""Development aid: re-traced fraction per decision group (libsart_g<k>.so variants of tools/build_unc_variants.sh)."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
names = ["bore/pipes", "opaque/spider", "shell", "mirror1", "mirror2", "nickel", "angle", "window/chip", "strips", "slow root"]
code = r'''
import sys
sys.path.insert(0, "%s"); sys.path.insert(0, "%s/tests")
from helpers import make_config
from solaraxionraytracing_b200 import raytracer as rt
FULL = dict(nR=1968, nE=1500, nAng=1000, nEn=1000)
for cfg in ("cast_llnl", "babyiaxo_xmm"):
    setup, tb = make_config(cfg, **FULL)
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        tr.set_precision(2)
        tr.reset_image(); tr.trace_mc(10**8, 299792458)
        c = tr.read_image().counters[0]
        print(cfg, "%%.3e" %% (c["n_retraced"] / 1e8), end="  ")
print()
''' % (ROOT, ROOT)
for g, name in enumerate(names):
    env = dict(os.environ, SART_LIB=str(ROOT / "solaraxionraytracing_b200" / f"libsart_g{g}.so"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"{name:14s}", r.stdout.strip() or r.stderr[-300:], flush=True)
