"""The oracle's restatement of the opacity-free emission-rate generator (src/readOpacityFile.nim:297-466, 598-860),
pinned on what can be pinned without the Nim program or its un-shipped output file:
  * fNew (:312-326, adaptive Gauss in the reference) against scipy's QUADPACK on the same integrand;
  * the Primakoff spectrum integrated over the AGSS09 Sun against its published shape: maximum at 3.0 keV, mean energy
    4.2 keV (CAST, JCAP 04 (2007) 010: dPhi/dE ~ E^2.481 exp(-E/1.205));
  * bfield (:328-351) known values; process bits are additive."""
import numpy as np
import pytest

from solaraxionraytracing_b200 import abi, tables


def test_fnew_against_quadpack(oracle):
    from scipy import integrate

    def inner(t, y):
        return 0.5 * (y * y / (t * t + y * y) + np.log(t * t + y * y))

    def outer(x, w, y):
        r = np.sqrt(x * x + w)
        return x * np.exp(-x * x) * (inner(r + x, y) - inner(r - x, y))

    L = oracle.lib()
    for w, y in [(0.08, 0.2), (1.0, 0.3), (5.0, 0.05), (19.0, 1.0), (0.001, 0.15), (300.0, 0.2)]:
        ref = integrate.quad(outer, 0, np.inf, args=(w, y), epsabs=1e-14, epsrel=1e-13)[0]
        assert L.oracle_fNew(w, y) == pytest.approx(ref, rel=1e-9)   # numericalnim's own tolerance is 1e-8


def test_primakoff_spectrum_shape(oracle):
    sm = tables.solar_model_packaged()
    assert sm.temp_K.shape == (1968,) and sm.mass_fractions.shape == (1968, 29)
    assert np.allclose(sm.radius, 0.0015 + 0.0005 * np.arange(1968))
    E = np.linspace(1e-3, 15.0, 1500)
    em = oracle.emission_rates(sm.temp_K, sm.rho_gcm3, sm.mass_fractions, E, abi.EM_PRIMAKOFF)
    assert np.isfinite(em).all() and (em >= 0).all()
    flux = (em * E[None, :] ** 2 * sm.radius[:, None] ** 2).sum(axis=0)     # the weighting of rt:2683-2686
    assert abs(E[flux.argmax()] - 3.0) < 0.15
    assert abs((flux * E).sum() / flux.sum() - 4.2) < 0.15
    analytic = E ** 2.481 * np.exp(-E / 1.205)
    sel = (E > 1.0) & (E < 10.0)
    ratio = flux[sel] / analytic[sel]
    assert ratio.std() / ratio.mean() < 0.05


def test_processes_are_additive_and_bfield(oracle):
    sm = tables.solar_model_packaged()
    sl = slice(0, 1968, 200)
    E = np.linspace(0.05, 15.0, 40)
    parts = [oracle.emission_rates(sm.temp_K[sl], sm.rho_gcm3[sl], sm.mass_fractions[sl], E, 1 << b) for b in range(6)]
    both = oracle.emission_rates(sm.temp_K[sl], sm.rho_gcm3[sl], sm.mass_fractions[sl], E, 63)
    assert np.allclose(sum(parts), both, rtol=1e-12, atol=0)
    assert all(np.isfinite(p).all() for p in parts)
    L = oracle.lib()
    unit = 1.0e6 * 1.4440271 * 1.0e-3 * np.sqrt(4.0 * np.pi)
    assert L.oracle_bfield(0.712) * unit == pytest.approx(50.0)     # tachocline peak (readOpacityFile.nim:335)
    assert L.oracle_bfield(0.96) * unit == pytest.approx(4.0)       # outer layers
    assert L.oracle_bfield(0.8) == 0.0
