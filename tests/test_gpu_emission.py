"""sart_emission_rates (GPU) against the oracle's restatement of readOpacityFile.nim's opacity-free processes.
Tolerance: 1e-7 relative per cell for every process (the reference integrates fNew adaptively to 1e-8; the kernel uses
a fixed 96-node Gauss-Legendre rule), cells below 1e-300 of the table maximum compared absolutely. Then the CDF build
on the generated table is bit-identical to the oracle's on the same table."""
import numpy as np
import pytest

from solaraxionraytracing_b200 import abi, tables

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt():
    from solaraxionraytracing_b200 import raytracer
    if raytracer.lib.sart_device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests need a B200")
    return raytracer


@pytest.mark.parametrize("process", sorted(abi.EM_PROCESSES))
def test_each_process_matches_oracle(rt, oracle, process):
    sm = tables.solar_model_packaged()
    sub = tables.SolarModel(sm.radius[:400], sm.temp_K[:400], sm.rho_gcm3[:400], sm.mass_fractions[:400])
    got = rt.calculateEmissionRates(sub, (process,), nElems=300)
    want = oracle.emission_rates(sub.temp_K, sub.rho_gcm3, sub.mass_fractions, got.energies, abi.EM_PROCESSES[process])
    assert got.emRates.shape == want.shape == (400, 300)
    scale = np.abs(want).max()
    assert scale > 0
    big = np.abs(want) > 1e-30 * scale
    assert np.allclose(got.emRates[big], want[big], rtol=1e-7, atol=0), np.max(np.abs(got.emRates[big] / want[big] - 1))
    assert np.all(np.abs(got.emRates[~big] - want[~big]) <= 1e-30 * scale)


def test_full_size_primakoff_table_and_cdfs(rt, oracle):
    sm = tables.solar_model_packaged()
    got = rt.calculateEmissionRates(sm, ("primakoff",))          # 1968 x 1500, the reference's grid
    want = oracle.emission_rates(sm.temp_K, sm.rho_gcm3, sm.mass_fractions, got.energies, abi.EM_PRIMAKOFF)
    # the Primakoff bracket (:384-392) subtracts two nearly equal logarithmic terms at high energy: libm vs CUDA math
    # differences of 1 ulp are amplified there
    nz = want != 0.0
    assert np.max(np.abs(got.emRates[nz] / want[nz] - 1.0)) < 1e-7
    # temperature carry-over of readOpacityFile.nim:686-690 never triggers for AGSS09, but zero cells below the plasma
    # frequency do (primakoff returns 0 for omega < omega_pl): same cells on both sides
    assert np.array_equal(got.emRates == 0.0, want == 0.0)
    rc, dc = rt.buildCdfs(got)
    rc_o, dc_o = oracle.build_cdfs(got.radii, got.energies, got.emRates)
    assert np.array_equal(rc, rc_o) and np.array_equal(dc, dc_o)


def test_all_processes_together(rt, oracle):
    sm = tables.solar_model_packaged()
    sl = slice(0, 1968, 16)
    sub = tables.SolarModel(sm.radius[sl], sm.temp_K[sl], sm.rho_gcm3[sl], sm.mass_fractions[sl])
    got = rt.calculateEmissionRates(sub, tuple(abi.EM_PROCESSES), nElems=500)
    # NB: radius index enters bfield() through 0.0015 + 0.0005 R, so a strided model changes the long-plasmon term on
    # both sides identically
    want = oracle.emission_rates(sub.temp_K, sub.rho_gcm3, sub.mass_fractions, got.energies, 63)
    assert np.allclose(got.emRates, want, rtol=1e-7, atol=1e-300)


def test_rejects_bad_arguments(rt):
    sm = tables.solar_model_packaged()
    with pytest.raises(ValueError):
        rt.calculateEmissionRates(sm, ("plasmon_T",))
    import ctypes as C
    z = np.zeros(4)
    p = z.ctypes.data_as(abi.c_double_p)
    assert rt.lib.sart_emission_rates(0, 1, p, p, p, 1, p, 0, 1e-13, 1e-12, 1e-15, p) == -1   # SART_ERR_ARG
    assert rt.lib.sart_emission_rates(0, 1, p, p, p, 1, p, 1 << 7, 1e-13, 1e-12, 1e-15, p) == -1
