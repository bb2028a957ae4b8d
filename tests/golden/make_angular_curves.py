#!/usr/bin/env python
"""Generates tests/golden/angular_scan_reference_curves.npz from the reference checkout: the two external curves
`performAngularScan` overlays on its result (src/raytracer.nim:2805-2810) — the McXtrace simulation of the XMM optic
(resources/McXtrace_angular_xmm.csv: angle [deg], flux, relative flux) and ESA's XMM vignetting curve
(resources/xmm_newton_angular_effective_area.csv: angle [arcmin], effective area). They are the only externally pinned
full-run numbers the reference holds for the ray-tracing path. Run:  python tests/golden/make_angular_curves.py"""
import sys
from pathlib import Path

import numpy as np

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference") / "resources"
mc = np.loadtxt(ref / "McXtrace_angular_xmm.csv", delimiter=",", skiprows=1)
x = np.loadtxt(ref / "xmm_newton_angular_effective_area.csv", delimiter=",", comments="#")
np.savez_compressed(Path(__file__).resolve().parent / "angular_scan_reference_curves.npz",
                    mcxtrace_angle_deg=mc[:, 0], mcxtrace_flux=mc[:, 1], mcxtrace_rel=mc[:, 2],
                    xmm_angle_arcmin=x[:, 0], xmm_eff_area=x[:, 1])
print("written", mc.shape, x.shape)
