"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this. The product
package (solaraxionraytracing_b200) never does. It borrows the boundary's POD types (abi.py mirrors include/sart.h).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from solaraxionraytracing_b200 import abi

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    deps = [HERE / "oracle.c", HERE / "oracle_emission.c", HERE / "Makefile", HERE.parent / "include" / "sart.h"]
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        subprocess.run(["make", "-C", str(HERE), "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        dp, ip = abi.c_double_p, abi.c_int32_p
        S, T, R = C.POINTER(abi.Setup), C.POINTER(abi.Tables), C.POINTER(abi.RayOut)
        sig = {
            "oracle_philox": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
            "oracle_ray_uniforms": (None, [C.c_uint64, C.c_uint64, dp]),
            "oracle_trace_presampled": (C.c_int, [S, T, C.c_size_t, dp, dp, dp, R]),
            "oracle_trace_mc_rays": (C.c_int, [S, T, C.c_uint64, C.c_size_t, C.c_uint64, R]),
            "oracle_sample_rays": (C.c_int, [S, T, C.c_uint64, C.c_size_t, C.c_uint64, dp, dp, dp]),
            "oracle_sample_words": (C.c_int, [S, T, C.c_size_t, C.POINTER(C.c_uint32), dp, dp, dp]),
            "oracle_trace_words": (C.c_int, [S, T, C.c_size_t, C.POINTER(C.c_uint32), R]),
            "oracle_trace_mc": (C.c_int, [S, T, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, dp, dp, dp,
                                          C.POINTER(abi.Counters)]),
            "oracle_prepare_heatmap": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                                 C.c_size_t, dp, dp, dp, C.c_double, dp]),
            "oracle_build_cdfs": (C.c_int, [C.c_int, C.c_int, dp, dp, dp, dp, dp]),
            "oracle_calc_window_vals": (None, [C.c_double, C.c_int, C.c_double, dp, dp]),
            "oracle_length_telescope": (C.c_double, [S]),
            "oracle_test_mirrors": (None, [C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, dp, dp]),
            "oracle_density": (C.c_double, [C.c_double, C.c_double]),
            "oracle_effPhotonMass": (C.c_double, [C.c_double]),
            "oracle_effPhotonMass2": (C.c_double, [C.c_double] * 4),
            "oracle_axionConversionProb2": (C.c_double, [C.c_double] * 8),
            "oracle_intensitySuppression2": (C.c_double, [C.c_double] * 6),
            "oracle_conversionProb": (C.c_double, [S, C.c_double, C.c_double, C.c_double]),
            "oracle_emission_rates": (C.c_int, [C.c_int, dp, dp, dp, C.c_int, dp, C.c_uint32, C.c_double, C.c_double,
                                                C.c_double, dp]),
            "oracle_fNew": (C.c_double, [C.c_double, C.c_double]),
            "oracle_bfield": (C.c_double, [C.c_double]),
            "oracle_primakoff": (C.c_double, [C.c_double] * 9),
            "oracle_num_threads": (C.c_int, []),
            "oracle_set_num_threads": (None, [C.c_int]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(abi.c_double_p)


class RayBatch:
    """Numpy SoA backing of a sart_ray_out_t."""

    def __init__(self, n: int, optional: bool = True):
        self.n = n
        self.x = np.zeros(n); self.y = np.zeros(n); self.w = np.zeros(n)
        self.code = np.zeros(n, dtype=np.int32); self.shell = np.zeros(n, dtype=np.int32)
        for name in abi.RAY_OUT_OPTIONAL:   # optional: True / False, or the names of the optional arrays wanted
            want = (name in optional) if isinstance(optional, (tuple, list, set)) else bool(optional)
            setattr(self, name, np.zeros(n) if want else None)

    def c_struct(self) -> abi.RayOut:
        o = abi.RayOut()
        o.x, o.y, o.w = _dp(self.x), _dp(self.y), _dp(self.w)
        o.code = self.code.ctypes.data_as(abi.c_int32_p)
        o.shell = self.shell.ctypes.data_as(abi.c_int32_p)
        for name in abi.RAY_OUT_OPTIONAL:
            a = getattr(self, name)
            if a is not None:
                setattr(o, name, _dp(a))
        return o

    @property
    def exit_code(self) -> np.ndarray:
        return self.code & abi.CODE_MASK


def build_cdfs(radii, energies, emRates):
    radii = np.ascontiguousarray(radii, dtype=np.float64)
    energies = np.ascontiguousarray(energies, dtype=np.float64)
    emRates = np.ascontiguousarray(emRates, dtype=np.float64)
    nR, nE = emRates.shape
    rc = np.empty(nR); dc = np.empty((nR, nE))
    lib().oracle_build_cdfs(nR, nE, _dp(radii), _dp(energies), _dp(emRates), _dp(rc), _dp(dc))
    return rc, dc


def trace_presampled(setup: abi.Setup, tables, origin_xyz, exit_xy, energy, optional=True) -> RayBatch:
    n = energy.size
    out = RayBatch(n, optional)
    o = out.c_struct(); t = tables.c_struct()
    origin_xyz = np.ascontiguousarray(origin_xyz, dtype=np.float64)
    exit_xy = np.ascontiguousarray(exit_xy, dtype=np.float64)
    energy = np.ascontiguousarray(energy, dtype=np.float64)
    lib().oracle_trace_presampled(C.byref(setup), C.byref(t), n, _dp(origin_xyz), _dp(exit_xy), _dp(energy), C.byref(o))
    return out


def trace_mc_rays(setup: abi.Setup, tables, first_ray: int, n: int, seed: int, optional=True) -> RayBatch:
    out = RayBatch(n, optional)
    o = out.c_struct(); t = tables.c_struct()
    lib().oracle_trace_mc_rays(C.byref(setup), C.byref(t), first_ray, n, seed, C.byref(o))
    return out


def sample_rays(setup: abi.Setup, tables, first_ray: int, n: int, seed: int):
    origin = np.empty((3, n)); exit_xy = np.empty((2, n)); energy = np.empty(n)
    t = tables.c_struct()
    lib().oracle_sample_rays(C.byref(setup), C.byref(t), first_ray, n, seed, _dp(origin), _dp(exit_xy), _dp(energy))
    return origin, exit_xy, energy


def _words(words) -> np.ndarray:
    w = np.ascontiguousarray(words, dtype=np.uint32)
    if w.ndim != 2 or w.shape[0] != 6:
        raise ValueError("words must be [6, n] (phi_sun, theta_sun, radius, disc r, disc phi, energy)")
    return w


def sample_words(setup: abi.Setup, tables, words):
    """Emission point, exit-disc point and energy for caller-supplied random words [6, n] (u = (w + 0.5) 2^-32)."""
    w = _words(words)
    n = w.shape[1]
    origin = np.empty((3, n)); exit_xy = np.empty((2, n)); energy = np.empty(n)
    t = tables.c_struct()
    rc = lib().oracle_sample_words(C.byref(setup), C.byref(t), n, w.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(origin),
                                   _dp(exit_xy), _dp(energy))
    assert rc == 0
    return origin, exit_xy, energy


def trace_words(setup: abi.Setup, tables, words, optional=True) -> RayBatch:
    """oracle_trace_mc_rays with the six random words of every ray supplied by the caller instead of Philox."""
    w = _words(words)
    n = w.shape[1]
    out = RayBatch(n, optional)
    o = out.c_struct(); t = tables.c_struct()
    rc = lib().oracle_trace_words(C.byref(setup), C.byref(t), n, w.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(o))
    assert rc == 0
    return out


def trace_mc(setup: abi.Setup, tables, first_ray: int, n_rays: int, seed: int, masses=None):
    m = 1 if masses is None else len(masses)
    img = np.zeros((m, abi.IMAGE_BINS, abi.IMAGE_BINS)); img2 = np.zeros_like(img)
    cnt = (abi.Counters * m)()
    t = tables.c_struct()
    mp = None
    if masses is not None:
        marr = np.ascontiguousarray(masses, dtype=np.float64)
        mp = _dp(marr)
    lib().oracle_trace_mc(C.byref(setup), C.byref(t), first_ray, n_rays, seed, m, mp, _dp(img), _dp(img2), cnt)
    return img, img2, [c.as_dict() for c in cnt]


def ray_uniforms(seed: int, ray: int) -> np.ndarray:
    u = np.empty(6)
    lib().oracle_ray_uniforms(seed, ray, _dp(u))
    return u


def emission_rates(temp, rho, frac, energies, processes: int, g_ae=1e-13, gagamma=1e-12, ganuclei=1e-15) -> np.ndarray:
    """oracle_emission_rates (readOpacityFile.nim:598-860, opacity-free processes) -> emRates [nR, nE]."""
    temp = np.ascontiguousarray(temp, dtype=np.float64); rho = np.ascontiguousarray(rho, dtype=np.float64)
    frac = np.ascontiguousarray(frac, dtype=np.float64); energies = np.ascontiguousarray(energies, dtype=np.float64)
    assert frac.shape == (temp.size, 29)
    out = np.empty((temp.size, energies.size))
    lib().oracle_emission_rates(temp.size, _dp(temp), _dp(rho), _dp(frac), energies.size, _dp(energies), processes,
                                g_ae, gagamma, ganuclei, _dp(out))
    return out
