#!/usr/bin/env python
"""Development aid: cost split of the margin tests vs the FP64 re-trace, re-traced fraction, and per-field errors of the
FP32 records against the oracle at full table size."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from helpers import make_config
from oracle import oracle as orc
from solaraxionraytracing_b200 import abi, raytracer as rt

FULL = dict(nR=1968, nE=1500, nAng=1000, nEn=1000)
for cfg in sys.argv[1:] or ["cast_llnl", "babyiaxo_xmm"]:
    setup, tb = make_config(cfg, **FULL)
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        tr.set_precision(2)
        n = 10**9
        for mode, scale in ((1, 1.0), (0, 1.0), (1, 0.25), (1, 0.0)):
            tr.set_retrace(mode, scale)
            tr.reset_image(); tr.trace_mc(n // 10, 1); tr.synchronize()
            best = 1e9
            for _ in range(2):
                tr.reset_image(); tr.synchronize()
                t = time.perf_counter(); tr.trace_mc(n, 299792458); tr.synchronize(); best = min(best, time.perf_counter() - t)
            c = tr.read_image().counters[0]
            print(f"{cfg} retrace={mode} scale={scale}: {best*1e3:.2f} ms  retraced {c['n_retraced']/n:.3e} unresolved {c['n_unresolved']}", flush=True)
        tr.set_retrace(1, 1.0)
        m = 1_000_000
        origin, exit_xy, energy = orc.sample_rays(setup, tb, 0, m, 299792459)
        ref = orc.trace_presampled(setup, tb, origin, exit_xy, energy)
        gpu = tr.trace_presampled(origin, exit_xy, energy)
        ok = ((ref.code & 0xff) == 0) & ((gpu.code & 0xff) == 0)
        print(cfg, "passed", ok.sum(), "code mismatches", int((gpu.code != ref.code).sum()))
        for name in ("w", "reflect", "transMagnet", "transProbArgon", "alpha1", "alpha2", "pathCB", "energy", "x", "y"):
            a, b = getattr(gpu, name)[ok], getattr(ref, name)[ok]
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
            print(f"   {name:15s} rel err median {np.median(err):.2e} 99% {np.quantile(err, 0.99):.2e} max {err.max():.2e}  frac>1e-3 {np.mean(err > 1e-3):.3e}")
        bad = np.flatnonzero(ok)[np.abs(gpu.w[ok] / ref.w[ok] - 1) > 1e-2][:8]
        for i in bad:
            print("   ray", i, "E", ref.energy[i], "shell", ref.shell[i], "a1", ref.alpha1[i], gpu.alpha1[i], "refl", ref.reflect[i], gpu.reflect[i],
                  "tm", ref.transMagnet[i], gpu.transMagnet[i], "w", ref.w[i], gpu.w[i], "xy", ref.x[i], ref.y[i])
