#!/usr/bin/env python
"""Development aid: per-ray comparison of the fast pipeline against the exact one on the GPU."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
from helpers import make_config  # noqa: E402
from solaraxionraytracing_b200 import abi, raytracer as rt  # noqa: E402


def cmp(cfg, n=2_000_000, seed=299792458):
    setup, tb = make_config(cfg, nR=1968, nE=1500, nAng=1000, nEn=1000)
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        ex = tr.traceAxionWrapper(n, seed)
        tr.set_precision(1)
        fa = tr.traceAxionWrapper(n, seed)
        tr.reset_image(); tr.trace_mc(n, seed); rf = tr.read_image()
        tr.set_precision(0); tr.reset_image(); tr.trace_mc(n, seed); re_ = tr.read_image()
    ce, cf = ex.exit_code, fa.exit_code
    mism = ce != cf
    print(f"== {cfg}: n={n} exit-code mismatches {mism.sum()} ({mism.mean():.2e})")
    if mism.any():
        pairs, cnt = np.unique(np.stack([ce[mism], cf[mism]]), axis=1, return_counts=True)
        for (a, b), c in zip(pairs.T, cnt):
            print(f"     exact {abi.EXIT_NAMES[a]:>16} -> fast {abi.EXIT_NAMES[b]:<16} {c}")
    ok = (~mism) & (ce == 0)
    if ok.any():
        dx, dy = np.abs(ex.x[ok] - fa.x[ok]), np.abs(ex.y[ok] - fa.y[ok])
        dw = np.abs(fa.w[ok] / ex.w[ok] - 1)
        print(f"   passed both: {ok.sum()}  |dx| max {dx.max():.3e} p99 {np.quantile(dx, 0.99):.3e} mean {dx.mean():.3e};"
              f" |dy| max {dy.max():.3e} mean {dy.mean():.3e}")
        print(f"   rel dw max {dw.max():.3e} p99 {np.quantile(dw, 0.99):.3e} mean {dw.mean():.3e};"
              f" energy equal {np.array_equal(ex.energy.astype(np.float32), fa.energy.astype(np.float32))};"
              f" shell equal {np.array_equal(ex.shell[ok], fa.shell[ok])}")
    e, f = re_.counters[0], rf.counters[0]
    print("   counters exact:", {k: v for k, v in e["n_exit"].items() if v}, "tillW", e["n_passed_till_window"], "sum_w %.6e" % e["sum_w"])
    print("   counters fast :", {k: v for k, v in f["n_exit"].items() if v}, "tillW", f["n_passed_till_window"], "sum_w %.6e" % f["sum_w"])
    tot = re_.image.sum()
    if tot > 0:
        print("   image L1 diff / total: %.3e ; sum ratio %.8f" % (np.abs(rf.image - re_.image).sum() / tot, rf.image.sum() / tot))


if __name__ == "__main__":
    for cfg in (sys.argv[1:] or ["cast_llnl", "babyiaxo_xmm", "babyiaxo_gas", "cast_abrixas"]):
        cmp(cfg)
