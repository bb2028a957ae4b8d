"""XMM's central blocker with a hole pattern (rt:1674-1688, lineIntersectsObject rt:494-527) and the effective-area
"reflectivity" (rkEffectiveArea, rt:1553-1562) in all three pipelines against the CPU oracle.

Neither is used by a BASELINE configuration; round 1 had them in the exact pipeline only (untested) and the throughput
pipelines refused such setups. Tolerances: exit codes identical in modes 0 and 2 (mode 2 re-traces what FP32 cannot
decide), <= 1e-4 of the rays different in mode 1 (its stated classification tolerance); weights 1e-10 relative in mode 0,
99 % of the rays within 2e-4 in modes 1 and 2 (FP32 weight factors).
"""
import numpy as np
import pytest

from helpers import make_config
from solaraxionraytracing_b200 import abi

pytestmark = pytest.mark.gpu
SEED = 299792458
HOLES = [(abi.HT_CROSS, 5, 3.0), (abi.HT_STAR, 3, 2.5), (abi.HT_CIRCLE, 5, 6.0), (abi.HT_SQUARE, 1, 20.0),
         (abi.HT_DIAMOND, 4, 7.0)]


@pytest.fixture(scope="module")
def rt():
    from solaraxionraytracing_b200 import raytracer
    if raytracer.lib.sart_device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests need a B200")
    return raytracer


def _central_rays(oracle, setup, tb, n, seed):
    """Pre-sampled solar rays whose exit-disc point lies within 70 mm of the axis: all of them meet XMM's central blocker
    (radius 64.7 mm) or its rim, instead of the 3 % of a ray set spread over BabyIAXO's 350 mm bore."""
    origin, exit_xy, energy = oracle.sample_rays(setup, tb, 0, n, seed)
    r = np.hypot(exit_xy[0], exit_xy[1])
    scale = np.minimum(1.0, 70.0 / np.maximum(r, 1e-9)) * np.random.default_rng(seed).uniform(0.0, 1.0, n) ** 0.5
    return origin, np.ascontiguousarray(exit_xy * scale), energy


@pytest.mark.parametrize("hole", HOLES, ids=["cross", "star", "circle", "square", "diamond"])
def test_xmm_hole_patterns_all_pipelines(rt, oracle, hole):
    setup, tb = make_config("babyiaxo_xmm")
    setup.telescope.holeType, setup.telescope.numberOfHoles, setup.telescope.holeInOptics = hole
    n = 2_000_000
    origin, exit_xy, energy = _central_rays(oracle, setup, tb, n, SEED + hole[0])
    ref = oracle.trace_presampled(setup, tb, origin, exit_xy, energy, optional=False)
    closed, _ = make_config("babyiaxo_xmm")   # htNone: the same rays against the closed blocker
    ref_closed = oracle.trace_presampled(closed, tb, origin, exit_xy, energy, optional=False)
    opened = (ref_closed.exit_code == abi.EXIT_OPAQUE) & (ref.exit_code != abi.EXIT_OPAQUE)
    assert opened.sum() > n // 500, "the hole pattern must open the blocker for a visible share of the rays"
    assert (ref.exit_code == abi.EXIT_OPAQUE).sum() > n // 10
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        for mode in (0, 1, 2):
            tr.set_precision(mode)
            gpu = tr.trace_presampled(origin, exit_xy, energy, optional=False)
            mism = np.flatnonzero(gpu.code != ref.code)
            if mode == 1:
                assert mism.size <= n * 1e-4, (mode, mism.size)
            else:
                assert mism.size == 0, (mode, [(int(i), int(gpu.code[i]), int(ref.code[i])) for i in mism[:10]])
        # the fused Monte Carlo kernels take the same branch: integer counters of mode 2 equal the exact pipeline's
        tr.set_precision(0); tr.reset_image(); tr.trace_mc(5_000_000, SEED); e = tr.read_image().counters[0]
        tr.set_precision(2); tr.reset_image(); tr.trace_mc(5_000_000, SEED); f = tr.read_image().counters[0]
        assert f["n_exit"] == e["n_exit"] and f["n_unresolved"] == 0


def _eff_area_config(cfg):
    setup, tb = make_config(cfg)
    setup.telescope.reflKind = abi.RK_EFFECTIVE_AREA
    # a transmission curve with structure inside the solar energy range and a grid that ends at 9 keV, so that the
    # energies above carry the clamped flag (eval_linear1d)
    x = np.linspace(0.2, 9.0, 45)
    y = 0.55 * np.exp(-0.5 * ((x - 1.5) / 2.5) ** 2) + 0.05 + 0.02 * np.sin(3.0 * x)
    tb.telescopeTransmission = (x, y)
    return setup, tb


@pytest.mark.parametrize("cfg", ["cast_llnl", "babyiaxo_xmm"])
def test_effective_area_all_pipelines(rt, oracle, cfg):
    setup, tb = _eff_area_config(cfg)
    n = 1_000_000
    origin, exit_xy, energy = oracle.sample_rays(setup, tb, 0, n, SEED + 3)
    ref = oracle.trace_presampled(setup, tb, origin, exit_xy, energy, optional=True)
    passed = ref.exit_code == abi.EXIT_PASSED
    assert passed.sum() > n // 10
    assert (ref.code[passed] & abi.FLAG_INTERP_CLAMPED).any() and not (ref.code[passed] & abi.FLAG_INTERP_CLAMPED).all()
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        for mode in (0, 1, 2):
            tr.set_precision(mode)
            gpu = tr.trace_presampled(origin, exit_xy, energy, optional=True)
            mism = np.flatnonzero(gpu.code != ref.code)
            if mode == 1:
                assert mism.size <= n * 1e-4, (mode, mism.size)
            else:
                assert mism.size == 0, (mode, [(int(i), int(gpu.code[i]), int(ref.code[i])) for i in mism[:10]])
            both = passed & (gpu.exit_code == abi.EXIT_PASSED)
            dw = np.abs(gpu.w[both] / ref.w[both] - 1.0)
            dr = np.abs(gpu.reflect[both] / ref.reflect[both] - 1.0)
            print(cfg, "mode", mode, "weight diff 99% / max", np.quantile(dw, 0.99), dw.max(), "reflect max", dr.max())
            if mode == 0:
                assert dw.max() <= 1e-10 and dr.max() <= 1e-12
            else:
                assert np.quantile(dw, 0.99) <= 2e-4 and dr.max() <= 2e-5
        # fused Monte Carlo run: flux of the throughput modes against the exact pipeline on the same Philox rays
        m = 4_000_000
        tr.set_precision(0); tr.reset_image(); tr.trace_mc(m, SEED); e = tr.read_image().counters[0]
        for mode in (1, 2):
            tr.set_precision(mode); tr.reset_image(); tr.trace_mc(m, SEED); f = tr.read_image().counters[0]
            assert abs(f["sum_w"] / e["sum_w"] - 1.0) < 3e-4, (mode, f["sum_w"], e["sum_w"])
            assert f["n_interp_clamped"] == e["n_interp_clamped"] or mode == 1
            if mode == 2:
                assert f["n_exit"] == e["n_exit"]
