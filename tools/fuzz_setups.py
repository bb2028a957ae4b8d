#!/usr/bin/env python
"""Randomised differential test of the FP32 pipeline against the exact one (development aid; the permanent versions of what
it found live in tests/test_gpu_retrace.py): random telescope turns, detector shifts, ignore* flags, X-ray sources and
hole patterns on top of the five base setups, N rays each in precision 0 and 2 (both compaction settings), every integer
counter compared.   python tools/fuzz_setups.py [n_setups] [rays] [seed]"""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_config
from solaraxionraytracing_b200 import abi, raytracer as rt



def random_setup(rng):
    """One random setup; consumes the generator exactly as the runs recorded in DESIGN.md did."""
    base = str(rng.choice(["cast_llnl", "babyiaxo_xmm", "cast_abrixas", "babyiaxo_gas", "cast_xmm"]))
    flags = 0
    desc = [base]
    xray = rng.random() < 0.3
    for name, bit in (("ignoreDetWindow", abi.CF_IGNORE_DET_WINDOW), ("ignoreGasAbs", abi.CF_IGNORE_GAS_ABS),
                      ("ignoreConvProb", abi.CF_IGNORE_CONV_PROB), ("ignoreReflection", abi.CF_IGNORE_REFLECTION)):
        if rng.random() < 0.25:
            flags |= bit; desc.append(name)
    if xray:
        flags |= abi.CF_XRAY_TEST
    setup, tb = make_config(base, flags=flags)
    if rng.random() < (1.0 if os.environ.get("FUZZ_TURN") else 0.5):
        setup.telescope.telescope_turned_x = float(rng.uniform(-0.3, 0.3))
        setup.telescope.telescope_turned_y = float(rng.uniform(-0.3, 0.3))
        desc.append("turned %.3f %.3f" % (setup.telescope.telescope_turned_x, setup.telescope.telescope_turned_y))
    if rng.random() < 0.3:
        setup.detectorInstall.lateralShift = float(rng.uniform(-3, 3))
        setup.detectorInstall.transversalShift = float(rng.uniform(-3, 3))
        desc.append("shift %.2f %.2f" % (setup.detectorInstall.lateralShift, setup.detectorInstall.transversalShift))
    if xray:
        s = setup.testSource
        s.parallel = int(rng.random() < 0.5)
        s.energy = float(rng.uniform(0.5, 8.0))
        s.radius = float(rng.uniform(2.0, 25.0))
        s.offAxisUp = float(rng.uniform(-30, 30)) if rng.random() < 0.5 else 0.0
        s.offAxisLeft = float(rng.uniform(-30, 30)) if rng.random() < 0.5 else 0.0
        s.distance = float(rng.uniform(2000.0, 12000.0)); s.lengthCol = float(rng.uniform(0.2, 0.8)) * s.distance
        desc.append("xray par=%d E=%.2f r=%.1f off=(%.1f, %.1f) d=%.0f col=%.0f" % (s.parallel, s.energy, s.radius, s.offAxisUp,
                                                                                   s.offAxisLeft, s.distance, s.lengthCol))
    if rng.random() < 0.3:   # geometry of the beam line and the detector
        setup.pipes.pipesTurned = float(rng.uniform(0.0, 4.0)); desc.append("pipesTurned %.2f" % setup.pipes.pipesTurned)
    if rng.random() < 0.3:
        setup.detector.theta = float(rng.uniform(0.0, 90.0)); desc.append("theta %.1f" % setup.detector.theta)
    if rng.random() < 0.2:
        setup.detector.radiusWindow *= float(rng.uniform(0.5, 1.5)); desc.append("rWin %.2f" % setup.detector.radiusWindow)
    if rng.random() < 0.2:
        setup.detectorInstall.distanceWindowFocalPlane = float(rng.uniform(-30.0, 30.0))
        desc.append("dWinFocal %.1f" % setup.detectorInstall.distanceWindowFocalPlane)
    if rng.random() < 0.2:
        setup.magnet.radiusCB *= float(rng.uniform(0.6, 1.0)); desc.append("rCB %.1f" % setup.magnet.radiusCB)
    if rng.random() < 0.2:
        setup.pipes.cb2vt3_radius *= float(rng.uniform(0.55, 1.0)); setup.pipes.vt3xrt_radius *= float(rng.uniform(0.55, 1.0))
        desc.append("pipes r %.1f %.1f" % (setup.pipes.cb2vt3_radius, setup.pipes.vt3xrt_radius))
    if rng.random() < 0.2:
        setup.telescope.optics_entrance[0] += float(rng.uniform(-3, 3)); setup.telescope.optics_entrance[1] += float(rng.uniform(-3, 3))
        desc.append("oe %.2f %.2f" % (setup.telescope.optics_entrance[0], setup.telescope.optics_entrance[1]))
    if rng.random() < 0.15 and not (flags & abi.CF_IGNORE_REFLECTION):
        setup.telescope.reflKind = abi.RK_EFFECTIVE_AREA
        x = np.linspace(0.2, 9.0, 30); tb.telescopeTransmission = (x, 0.4 + 0.1 * np.sin(x)); desc.append("effArea")
    if setup.telescope.kind == abi.TK_XMM and rng.random() < 0.4:
        setup.telescope.holeType = int(rng.integers(1, 6)); setup.telescope.numberOfHoles = int(rng.integers(1, 8))
        setup.telescope.holeInOptics = float(rng.uniform(0.5, 12.0))
        desc.append("hole %d x%d R=%.1f" % (setup.telescope.holeType, setup.telescope.numberOfHoles, setup.telescope.holeInOptics))
    return setup, tb, desc


def main():
    n_setups = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
    rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    bad = 0
    for it in range(n_setups):
        setup, tb, desc = random_setup(rng)
        try:
            with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
                seed = int(rng.integers(1, 2**62))
                if os.environ.get("FUZZ_DIAG") and int(os.environ["FUZZ_DIAG"]) == it:   # locate the differing rays of one setup
                    first = int(rng.choice([0, 17, 2**32 - 12345, 2**40 + 3]))
                    chunk = 10_000_000
                    for k in range(n // chunk):
                        tr.set_precision(0); ex = tr.traceAxionWrapper(chunk, seed, first_ray=first + k * chunk, optional=False)
                        tr.set_precision(2); fa = tr.traceAxionWrapper(chunk, seed, first_ray=first + k * chunk, optional=False)
                        for i in np.flatnonzero(ex.code != fa.code):
                            print("DIAG ray", first + k * chunk + int(i), "seed", seed, "exact code", hex(int(ex.code[i])), "f32 code", hex(int(fa.code[i])),
                                  "exact x,y", ex.x[i], ex.y[i], flush=True)
                            for scale in (1.0, 1.5, 2.0, 4.0, 16.0):
                                tr.set_retrace(1, scale)
                                one = tr.traceAxionWrapper(1, seed, first_ray=first + k * chunk + int(i), optional=False)
                                print("   budgets x", scale, "-> f32 code", hex(int(one.code[0])), flush=True)
                            for name in ("latS", "latT", "latA", "detS", "detT", "detA", "rho", "discRel", "zrel", "nick", "sinA", "cond", "spider", "entK"):
                                os.environ["SART_TOL_BOOST"] = name + "=4"
                                tr.set_retrace(1, 0.5); tr.set_retrace(1, 1.0)   # a changed scale makes the library derive the budgets anew
                                one = tr.traceAxionWrapper(1, seed, first_ray=first + k * chunk + int(i), optional=False)
                                print("   4 x", name, "-> f32 code", hex(int(one.code[0])), flush=True)
                            os.environ.pop("SART_TOL_BOOST")
                            tr.set_retrace(1, 0.5); tr.set_retrace(1, 1.0)
                    print("DIAG done:", "; ".join(desc)); sys.exit(0)
                if os.environ.get("FUZZ_DIAG"):
                    rng.choice([0, 17, 2**32 - 12345, 2**40 + 3]); continue
                first = int(rng.choice([0, 17, 2**32 - 12345, 2**40 + 3]))
                if os.environ.get("FUZZ_MASSES") and setup.stage == abi.SK_GAS:   # a mass scan around the resonance instead
                    masses = np.sort(np.concatenate([[0.008235101411623404], rng.uniform(0.002, 0.03, 4)]))
                    tr.set_axion_masses(masses); desc.append("masses " + " ".join("%.5f" % m for m in masses))
                    m1 = n // 4
                    tr.trace_mc(m1, seed, first_ray=first); e = tr.read_image().counters
                    tr.set_precision(2); tr.reset_image(); tr.trace_mc(m1, seed, first_ray=first); f = tr.read_image().counters
                    diff = {}
                    for k in range(len(masses)):
                        for key, v in e[k]["n_exit"].items():
                            if f[k]["n_exit"][key] != v: diff[(k, key)] = (f[k]["n_exit"][key], v)
                        if f[k]["n_passed_till_window"] != e[k]["n_passed_till_window"]: diff[(k, "till")] = 1
                        if e[k]["sum_w"] > 0 and abs(f[k]["sum_w"] / e[k]["sum_w"] - 1) > 1e-5: diff[(k, "flux")] = f[k]["sum_w"] / e[k]["sum_w"]
                    bad += bool(diff)
                    print("%3d %s  mass scan  %s" % (it, "DIFF" if diff else "ok  ", "; ".join(desc)), flush=True)
                    if diff: print("     ", diff, flush=True)
                    continue
                tr.trace_mc(n, seed, first_ray=first); e = tr.read_image().counters[0]
                tr.set_precision(2)
                res = []
                for compact in (0, 1):
                    tr.set_compaction(compact); tr.reset_image(); tr.trace_mc(n, seed, first_ray=first); f = tr.read_image().counters[0]
                    diff = {k: (f["n_exit"][k], v) for k, v in e["n_exit"].items() if f["n_exit"][k] != v}
                    for key in ("n_passed_till_window", "n_interp_clamped"):
                        if f[key] != e[key]: diff[key] = (f[key], e[key])
                    if f["n_unresolved"]: diff["unresolved"] = f["n_unresolved"]
                    res.append((diff, f["n_retraced"] / n))
            ok = not res[0][0] and not res[1][0]
            bad += not ok
            print("%3d %s  passed=%.4f retraced=%.2e  %s" % (it, "ok  " if ok else "DIFF", e["n_passed"] / n, res[0][1], "; ".join(desc)), flush=True)
            if not ok: print("     ", res[0][0], res[1][0], flush=True)
        except rt.SartError as ex:
            print("%3d skip (%s): %s" % (it, str(ex)[:80], "; ".join(desc)), flush=True)
    print("setups with differing counters:", bad)


if __name__ == "__main__":
    main()
