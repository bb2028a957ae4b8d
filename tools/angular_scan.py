#!/usr/bin/env python
"""performAngularScan (src/raytracer.nim:2778-2815) on the GPU: relative flux of the XMM optic versus telescope_turned_y,
printed next to the McXtrace curve the reference overlays (resources/McXtrace_angular_xmm.csv, packaged in
tests/golden/angular_scan_reference_curves.npz).

  python tools/angular_scan.py [--rays 10000000] [--experiment BabyIAXO] [--precision fast|exact] [--xray]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from solaraxionraytracing_b200 import abi, raytracer as rt, tables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=float, default=1e7)
    ap.add_argument("--experiment", default="BabyIAXO")
    ap.add_argument("--detector", default="InGridIAXO")
    ap.add_argument("--precision", default="fast")
    ap.add_argument("--xray", action="store_true", help="parallel X-ray test source instead of the Sun")
    ap.add_argument("--flags", default="")
    ap.add_argument("--chip", type=float, default=0.0, help="chip edge length in mm (default: the reference's 14 mm)")
    ap.add_argument("--energy", type=float, default=0.0, help="X-ray source energy in keV (with --xray)")
    a = ap.parse_args()
    curves = np.load(ROOT / "tests" / "golden" / "angular_scan_reference_curves.npz")
    angles = curves["mcxtrace_angle_deg"]
    flags = rt.flags_from_cli(xrayTest=a.xray, **{f: True for f in a.flags.split(",") if f})
    setup = rt.newExperimentSetup(a.experiment, a.detector, "vacuum", "XMM", flags)
    if a.xray:
        setup.testSource.parallel = 1
        if a.energy > 0:
            setup.testSource.energy = a.energy
    if a.chip > 0:
        setup.consts.chipXMax = setup.consts.chipYMax = a.chip
    t0 = time.perf_counter()
    em = rt.calculateEmissionRates(processes=("primakoff",))
    rc, dc = rt.buildCdfs(em)
    tb = tables.TableSet(energies=em.energies, fluxRadiusCDF=rc, diffFluxCDFs=dc,
                         reflectivity=tables.gold_reflectivity_packaged(), **tables.detector_tables_packaged())
    t1 = time.perf_counter()
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        tr.set_precision(1 if a.precision == "fast" else 0)
        t2 = time.perf_counter()
        fl, cnt, _ = tr.angular_scan(angles, int(a.rays))
        t3 = time.perf_counter()
    rel = fl / fl.max()
    print(f"tables {t1 - t0:.2f} s, create {t2 - t1:.2f} s, scan of {angles.size} x {int(a.rays):.1e} rays {t3 - t2:.3f} s")
    print("angle[deg]  rel.flux   McXtrace   passed")
    for ang, r, m, c in zip(angles, rel, curves["mcxtrace_rel"], cnt):
        print(f"{ang:8.3f}  {r:9.4f}  {m:9.4f}  {c['n_passed']:10d}")


if __name__ == "__main__":
    main()
