#!/usr/bin/env python
"""Development aid: list the rays where the fast and exact pipelines disagree most."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from helpers import make_config
from solaraxionraytracing_b200 import abi, raytracer as rt

cfg = sys.argv[1] if len(sys.argv) > 1 else "babyiaxo_xmm"
n = 2_000_000
setup, tb = make_config(cfg, nR=1968, nE=1500, nAng=1000, nEn=1000)
with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
    ex = tr.traceAxionWrapper(n, 299792458)
    tr.set_precision(1)
    fa = tr.traceAxionWrapper(n, 299792458)
ok = (ex.exit_code == 0) & (fa.exit_code == 0)
d = np.hypot(ex.x - fa.x, ex.y - fa.y); d[~ok] = 0
idx = np.argsort(-d)[:12]
print("quantiles of |d| (mm):", [float(np.quantile(d[ok], q)) for q in (0.5, 0.9, 0.99, 0.999, 0.9999, 1.0)])
for i in idx:
    print(f"ray {i}: d={d[i]:.3e} exact=({ex.x[i]:.6f},{ex.y[i]:.6f}) fast=({fa.x[i]:.6f},{fa.y[i]:.6f}) shell={ex.shell[i]} "
          f"a1={ex.alpha1[i]:.5f} a2={ex.alpha2[i]:.5f} pathCB={ex.pathCB[i]:.3f} r={ex.r[i]:.4f} devDet={ex.deviationDet[i]:.4f} E={ex.energy[i]:.3f}")
# correlation with alpha2/alpha1 ratio
big = ok & (d > 5e-3)
print("n big:", big.sum(), " alpha1 range", ex.alpha1[big].min() if big.any() else None, ex.alpha1[big].max() if big.any() else None)
print("   alpha2 range", ex.alpha2[big].min() if big.any() else None, ex.alpha2[big].max() if big.any() else None)
print("   shells", np.unique(ex.shell[big], return_counts=True))
