"""Pins the CPU oracle against every known-answer value the reference holds for this path.
No per-ray golden outputs exist in the reference; these are the formulas' documented values."""
import ctypes as C
import json
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_setup
from solaraxionraytracing_b200 import abi

GOLDEN = Path(__file__).resolve().parent / "golden"


def _llnl():
    return ref_setup.make_setup(abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, 0)


def test_eff_photon_mass_table(oracle):
    """axionMass/axionMassforMagnet.nim:116-119: p(mbar) at T = 100 K -> m_gamma (eV)."""
    L = oracle.lib()
    for p, m in ((36.61, 0.0853), (109.8, 0.1477), (183.05, 0.1907), (366.1, 0.2698)):
        assert abs(L.oracle_effPhotonMass2(p, 10.0, 0.35, 100.0) - m) < 1e-4
    # independent of the bore volume (length/radius cancel): am:51-61
    assert L.oracle_effPhotonMass2(36.61, 3.0, 0.1, 100.0) == pytest.approx(L.oracle_effPhotonMass2(36.61, 10.0, 0.35, 100.0), rel=1e-14)


def test_eff_photon_mass_iaxo_4K(oracle):
    """axionMass/axionMass.org:802-806: 14.3345 / 43.0034 mbar at 4.2 K -> 0.26048 / 0.45117 eV."""
    L = oracle.lib()
    assert abs(L.oracle_effPhotonMass2(14.3345, 10.0, 0.3, 4.2) - 0.26048) < 2e-4
    assert abs(L.oracle_effPhotonMass2(43.0034, 10.0, 0.3, 4.2) - 0.45117) < 3e-4


def test_babyiaxo_eff_mass_relation(oracle):
    """axionMass/axionMass.nim:75-87: m_gamma = 1.94081e-2 * sqrt(4 pi n_e[1e?]) form agrees to 1e-5 at 4.2 K
    through the ideal-gas density (p = 1 mbar -> 0.0688 eV)."""
    assert abs(oracle.lib().oracle_effPhotonMass2(1.0, 10.0, 0.3, 4.2) - 0.068800) < 5e-5


def test_vacuum_conversion_probability(oracle):
    """conversionProb(9 T, 1e-12 GeV^-1, 9.26 m) = 1.70e-21 (rt:363-365 with unchained's natural units)."""
    s = _llnl()
    p = oracle.lib().oracle_conversionProb(C.byref(s), 9.0, 1e-12, 9260.0)
    assert p == pytest.approx(1.70e-21, rel=5e-3)
    # (g B L / 2)^2 scaling
    assert oracle.lib().oracle_conversionProb(C.byref(s), 4.5, 1e-12, 9260.0) == pytest.approx(p / 4, rel=1e-12)


def test_vacuum_mass_limit_and_resonance(oracle):
    """axionMass.nim:105-107: coherence is lost at q L = pi, i.e. m_a = sqrt(2 pi E / L) = 1.61e-2 eV at 4.2 keV, 20 m;
    at m_a = m_gamma the gas conversion probability is maximal and symmetric in q (am:75-100)."""
    E, Lm = 4.2e3, 20.0
    assert math.sqrt(2 * math.pi * E / (Lm / 1.97e-7)) == pytest.approx(1.61e-2, abs=1e-3)
    L = oracle.lib()
    mg = L.oracle_effPhotonMass2(1.0, 10.0, 0.3, 4.2)
    on = L.oracle_axionConversionProb2(mg, 4.2, 1.0, 4.2, 10.0, 0.3, 1e-12, 2.0)
    off = L.oracle_axionConversionProb2(mg * 1.05, 4.2, 1.0, 4.2, 10.0, 0.3, 1e-12, 2.0)
    assert on > off > 0


def test_intensity_suppression_is_beer_lambert(oracle):
    L = oracle.lib()
    a = L.oracle_intensitySuppression2(3.0, 10.0, 1.0, 100.0, 100.0, 293.15)
    b = L.oracle_intensitySuppression2(3.0, 20.0, 2.0, 100.0, 100.0, 293.15)
    assert 0 < b < a < 1 and b == pytest.approx(a * a, rel=1e-12)


def test_window_strip_geometry(oracle):
    """rt:1431-1462 with r = 7 mm, 4 strips, open ratio 0.838 (also calculateWindowValues.nim:7-33)."""
    w, d = C.c_double(), C.c_double()
    oracle.lib().oracle_calc_window_vals(7.0, 4, 0.838, C.byref(w), C.byref(d))
    assert w.value == pytest.approx(0.500418, abs=1e-6) and d.value == pytest.approx(2.299582, abs=1e-6)
    # strongback where 1.1498 < |y| < 1.6502 or 3.9498 < |y| < 4.4502 (rt:2167-2169)
    assert 0.5 * d.value == pytest.approx(1.1498, abs=1e-4) and 0.5 * d.value + w.value == pytest.approx(1.6502, abs=1e-4)
    assert 1.5 * d.value + w.value == pytest.approx(3.9498, abs=1e-4)


def test_length_telescope(oracle):
    """rt:1883-1884 gives 454.055 mm, matching optics_exit z = 454.0 (rt:1260)."""
    s = _llnl()
    assert oracle.lib().oracle_length_telescope(C.byref(s)) == pytest.approx(454.055, abs=2e-3)


def test_on_axis_ray_through_eighth_shell(oracle):
    """TestMirrors.nim:80-121 scenario: a ray parallel to the axis into LLNL shell index 7 (r1 = 83.405,
    xSep = 4.284, beta = 0.767 deg, l = 225): alpha1 = alpha2 = beta within 0.001 deg, exit 4 beta off axis."""
    r1, xsep, beta, l = 83.405, 4.284, 0.767, 225.0
    rho = r1 - 1.0   # hits the cone 1 mm below its entrance radius
    pcb = np.array([rho, 0.0, -100.0]); pxrt = np.array([rho, 0.0, 0.0])
    alpha = np.zeros(2); direc = np.zeros(3)
    dp = abi.c_double_p
    oracle.lib().oracle_test_mirrors(r1, xsep, beta, l, pcb.ctypes.data_as(dp), pxrt.ctypes.data_as(dp),
                                     alpha.ctypes.data_as(dp), direc.ctypes.data_as(dp))
    assert abs(alpha[0] - beta) < 1e-3 and abs(alpha[1] - beta) < 1e-3
    off_axis = math.degrees(math.atan2(math.hypot(direc[0], direc[1]), direc[2]))
    assert off_axis == pytest.approx(4 * beta, abs=2e-3)
    assert direc[0] < 0   # towards the axis


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        oracle.lib().oracle_philox(c, k, o)
        assert tuple(o) == want
    u = oracle.ray_uniforms(299792458, 12345)
    assert np.all((u > 0) & (u < 1)) and len(set(u.tolist())) == 6


def test_cdf_build_properties(oracle):
    """rt:2679-2705: monotone, ends at exactly 1, radius CDF weights rows by their integrated flux."""
    from solaraxionraytracing_b200 import tables
    em = tables.synthetic_emission(64, 50, "primakoff")
    rc, dc = oracle.build_cdfs(em.radii, em.energies, em.emRates)
    assert rc[-1] == 1.0 and np.all(dc[:, -1] == 1.0)
    assert np.all(np.diff(rc) >= 0) and np.all(np.diff(dc, axis=1) >= 0)
    flux = (em.emRates * em.energies[None, :] ** 2 * em.radii[:, None] ** 2).sum(axis=1)
    assert np.allclose(rc, np.cumsum(flux) / flux.sum(), rtol=1e-12)


def test_heatmap_matches_numpy(oracle):
    rng = np.random.default_rng(1)
    n = 5000
    x, y, w = rng.uniform(0, 14, n), rng.uniform(0, 14, n), rng.uniform(0, 1, n)
    out = np.zeros((256, 256))
    bad = oracle.lib().oracle_prepare_heatmap(256, 256, 0.0, 14.0, 0.0, 14.0, n, oracle._dp(x), oracle._dp(y),
                                              oracle._dp(w), 1.0, oracle._dp(out))
    assert bad == 0
    ref, _, _ = np.histogram2d(y, x, bins=256, range=((0, 14), (0, 14)), weights=w)   # image[y, x] rt:842
    assert np.allclose(out, ref, rtol=1e-12)


def test_oracle_regression_rays(oracle):
    """Committed per-ray fixture (tests/regression/make_regression.py): the oracle's own output of an earlier date, a
    guard against accidental changes of the oracle — not a golden vector of the reference (it holds none for this path)."""
    from helpers import make_config
    z = np.load(GOLDEN.parent / "regression" / "oracle_rays_v1.npz")
    meta = json.loads(str(z["meta"]))
    for cfg in meta["configs"]:
        setup, tb = make_config(cfg)
        r = oracle.trace_mc_rays(setup, tb, meta["first_ray"], meta["n"], meta["seed"])
        assert np.array_equal(r.code, z[f"{cfg}_code"]), cfg
        assert np.array_equal(r.shell, z[f"{cfg}_shell"])
        for name in ("x", "y", "w"):
            assert np.allclose(getattr(r, name), z[f"{cfg}_{name}"], rtol=1e-9, atol=1e-9), (cfg, name)
