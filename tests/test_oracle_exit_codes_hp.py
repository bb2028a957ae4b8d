"""The oracle's exit codes against 60-digit arithmetic (tests/hp_trace.py: classify), clipped rays included.

The oracle and the exact CUDA pipeline are two statements of the same reading of traceAxion; this is the independent
check of that reading's geometry: every decision of rt:1813-2147 re-derived from the reference's formulas in mpmath. For a
ray that stays further than 1e-2 mm from every decision boundary — far above the reference's own f64 rounding noise — the
exit code does not depend on the arithmetic, so the oracle must return the 60-digit one. (A ray that misses mirror 1 is
carried on by the reference with pointMirror1 = pointExitCB, rt:655-658, and ends as nickel or no_mirror_hit depending on
that garbage: either is accepted there.)"""
import numpy as np
import pytest

from helpers import make_config
from solaraxionraytracing_b200 import abi

import hp_trace


@pytest.mark.parametrize("cfg,n,kinds,turn", [("cast_llnl", 2500, 4, None), ("babyiaxo_xmm", 900, 5, None), ("cast_abrixas", 1500, 4, None),
                                              ("cast_llnl", 800, 3, (0.05, -0.04)), ("babyiaxo_xmm", 600, 4, (-0.18, 0.25))])
def test_oracle_exit_codes_agree_with_60_digit_geometry(oracle, cfg, n, kinds, turn):
    setup, tb = make_config(cfg)
    if turn:   # the frame change of a turned telescope (rt:1888-1905), as performAngularScan uses it
        setup.telescope.telescope_turned_x, setup.telescope.telescope_turned_y = turn
    origin, exit_xy, energy = oracle.sample_rays(setup, tb, 0, n, 424242)
    ref = oracle.trace_presampled(setup, tb, origin, exit_xy, energy, optional=False)
    code = ref.code & abi.CODE_MASK
    seen, checked = set(), 0
    for i in range(n):
        hp, margin, ambiguous = hp_trace.classify(setup, origin[:, i], exit_xy[:, i])
        if margin < 1e-2:
            continue
        got = int(code[i])
        if got == abi.EXIT_ZERO_WEIGHT:      # reached the detector window with weight 0: geometry says "passed"
            got = abi.EXIT_PASSED
        if ambiguous:
            assert got in (abi.EXIT_NO_MIRROR_HIT, abi.EXIT_NICKEL), (i, got, hp)
        else:
            assert got == hp, (i, got, hp, margin)
        seen.add(hp)
        checked += 1
    assert checked > 0.9 * n
    assert len(seen) >= kinds, seen      # clipped rays of several kinds, not only passed ones


@pytest.mark.parametrize("hole", [(abi.HT_CROSS, 5, 3.0), (abi.HT_STAR, 3, 2.5), (abi.HT_CIRCLE, 5, 6.0), (abi.HT_SQUARE, 1, 20.0),
                                  (abi.HT_DIAMOND, 4, 7.0)], ids=["cross", "star", "circle", "square", "diamond"])
def test_oracle_xmm_hole_patterns_agree_with_60_digit_geometry(oracle, hole):
    """XMM's central blocker with a hole pattern (rt:1674-1688, lineIntersectsObject rt:494-527), re-derived in mpmath: rays aimed
    at the blocker, the oracle's exit code against the 60-digit one for every ray further than 1e-2 mm from a decision boundary."""
    setup, tb = make_config("babyiaxo_xmm")
    setup.telescope.holeType, setup.telescope.numberOfHoles, setup.telescope.holeInOptics = hole
    n = 700
    origin, exit_xy, energy = oracle.sample_rays(setup, tb, 0, n, 99 + hole[0])
    r = np.hypot(exit_xy[0], exit_xy[1])
    exit_xy = np.ascontiguousarray(exit_xy * np.minimum(1.0, 70.0 / np.maximum(r, 1e-9)) * np.random.default_rng(hole[0]).uniform(0, 1, n) ** 0.5)
    ref = oracle.trace_presampled(setup, tb, origin, exit_xy, energy, optional=False)
    code = ref.code & abi.CODE_MASK
    checked, opened = 0, 0
    for i in range(n):
        hp, margin, ambiguous = hp_trace.classify(setup, origin[:, i], exit_xy[:, i])
        if margin < 1e-2:
            continue
        got = abi.EXIT_PASSED if int(code[i]) == abi.EXIT_ZERO_WEIGHT else int(code[i])
        if ambiguous:
            assert got in (abi.EXIT_NO_MIRROR_HIT, abi.EXIT_NICKEL), (i, got, hp)
        else:
            assert got == hp, (i, got, hp, margin)
        checked += 1
        opened += hp != abi.EXIT_OPAQUE
    assert checked > 0.8 * n and opened > 10 and checked - opened > 50, (checked, opened)
