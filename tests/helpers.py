"""Shared builders for the test-suite: the BASELINE.json configurations as (setup, tables) pairs.

CDFs are built with the ORACLE here (tests only); the product builds them on the GPU (sart_build_cdfs) and
test_gpu_parity checks the two are bit-identical.
"""
from __future__ import annotations

import functools

import numpy as np

from oracle import oracle as orc
from oracle import ref_setup
from solaraxionraytracing_b200 import abi, tables


@functools.lru_cache(maxsize=None)
def _emission(nR: int, nE: int, kind: str):
    em = tables.synthetic_emission(nR, nE, kind)
    rc, dc = orc.build_cdfs(em.radii, em.energies, em.emRates)
    return em, rc, dc


@functools.lru_cache(maxsize=None)
def _refl(nCoat: int, nAng: int, nEn: int):
    return tables.synthetic_reflectivity(nCoat, nAng, nEn)


def make_tables(nCoat: int, nR: int = 246, nE: int = 300, nAng: int = 200, nEn: int = 200, kind: str = "abc"):
    """Small-by-default tables so CPU tests stay fast; pass the full sizes (1968, 1500, 1000, 1000) for parity at
    BASELINE scale."""
    em, rc, dc = _emission(nR, nE, kind)
    det = tables.detector_tables_packaged()
    return tables.TableSet(energies=em.energies, fluxRadiusCDF=rc, diffFluxCDFs=dc,
                           reflectivity=_refl(nCoat, nAng, nEn), **det)


CONFIGS = {
    # name: (experiment, detector, stage, telescope, flags, emission kind)
    "cast_llnl": (abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, 0, "abc"),
    "cast_xmm": (abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_XMM, 0, "primakoff"),
    "babyiaxo_xmm": (abi.ES_BABYIAXO, abi.DK_INGRIDIAXO, abi.SK_VACUUM, abi.TK_XMM, 0, "primakoff"),
    "babyiaxo_gas": (abi.ES_BABYIAXO, abi.DK_INGRIDIAXO, abi.SK_GAS, abi.TK_XMM, 0, "primakoff"),
    "cast_abrixas": (abi.ES_CAST, abi.DK_INGRID2017, abi.SK_VACUUM, abi.TK_ABRIXAS, 0, "abc"),
}


def make_config(name: str, flags: int | None = None, **table_kw):
    ex, dk, sk, tk, fl, kind = CONFIGS[name]
    if flags is not None:
        fl = flags
    setup = ref_setup.make_setup(ex, dk, sk, tk, fl)
    tb = make_tables(setup.telescope.nCoatings, kind=kind, **table_kw)
    return setup, tb
