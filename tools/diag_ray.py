#!/usr/bin/env python
"""Development aid: the records of single Monte Carlo rays in precision modes 0 and 2 side by side."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from helpers import make_config
from solaraxionraytracing_b200 import abi, raytracer as rt
cfg = sys.argv[1]
rays = [int(a) for a in sys.argv[2:]]
setup, tb = make_config(cfg)
with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
    for ray in rays:
        out = {}
        for mode in (0, 2):
            tr.set_precision(mode)
            tr.set_retrace(0)
            out[mode] = tr.traceAxionWrapper(1, 299792458, first_ray=ray)
        for name in ("code", "shell", "x", "y", "w", "energy", "reflect", "transMagnet", "yaw", "alpha1", "alpha2", "pathCB", "r", "deviationDet", "transProbArgon"):
            print(f"ray {ray} {name:15s} exact {getattr(out[0], name)[0]!r:28}  f32 {getattr(out[2], name)[0]!r}")
        e = out[0].energy[0]
        i = int(np.argmin(np.abs(np.maximum(tb.energies, 0.03) - e)))
        print("energy index", i, "refl table min/max at that energy column:", tb.reflectivity[..., :].min(), tb.reflectivity.max(), "window/strongback/gas at E:",
              np.interp(e, *tb.windowTransmission), np.interp(e, *tb.strongbackTransmission), np.interp(e, *tb.gasAbsorption))
