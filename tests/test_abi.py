"""The C-ABI library loads on a CPU-only machine, exports every symbol include/sart.h declares, its struct layouts
match the ctypes mirror, its host-side setup constructors agree with an independent transcription of the reference,
and it fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_setup
from solaraxionraytracing_b200 import abi

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def rt():
    from solaraxionraytracing_b200 import raytracer
    return raytracer


def test_header_symbols_exported(rt):
    hdr = (ROOT / "include" / "sart.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sart_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(rt.lib, name), f"{name} declared in include/sart.h but not exported by libsart.so"
        assert name in abi.SIGNATURES, f"{name} missing from abi.SIGNATURES"


def test_struct_layouts(rt):
    assert rt.lib.sart_sizeof_setup() == C.sizeof(abi.Setup)
    assert rt.lib.sart_sizeof_tables() == C.sizeof(abi.Tables)
    assert rt.lib.sart_sizeof_counters() == C.sizeof(abi.Counters)
    assert rt.lib.sart_abi_version() == abi.ABI_VERSION


COMBOS = [(e, d, s, t) for e in (abi.ES_CAST, abi.ES_BABYIAXO)
          for d in (abi.DK_INGRID2017, abi.DK_INGRID2018, abi.DK_INGRIDIAXO)
          for s in (abi.SK_VACUUM, abi.SK_GAS) for t in (abi.TK_LLNL, abi.TK_XMM, abi.TK_ABRIXAS)]


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("flags", [0, abi.CF_XRAY_TEST | abi.CF_IGNORE_GAS_ABS])
def test_host_setup_matches_independent_transcription(rt, combo, flags):
    got = abi.struct_to_dict(rt.newExperimentSetup(*combo, flags))
    want = abi.struct_to_dict(ref_setup.make_setup(*combo, flags))

    def cmp(a, b, path=""):
        assert type(a) is type(b), path
        if isinstance(a, dict):
            assert a.keys() == b.keys()
            for k in a:
                cmp(a[k], b[k], f"{path}.{k}")
        elif isinstance(a, list):
            assert len(a) == len(b)
            for i, (x, y) in enumerate(zip(a, b)):
                cmp(x, y, f"{path}[{i}]")
        elif isinstance(a, float):
            assert a == pytest.approx(b, rel=1e-15, abs=0), path
        else:
            assert a == b, path
    cmp(got, want)


def test_enum_strings_and_errors(rt):
    s = rt.newExperimentSetup("BabyIAXO", "InGridIAXO", "vacuum", "XMM")   # config_default.toml:20-23
    assert s.magnet.radiusCB == 500.0 and s.telescope.nShells == 58 and s.detectorInstall.distanceDetectorXRT == 7500.0
    with pytest.raises(ValueError):
        rt.newExperimentSetup("CAST", "InGrid2018", "vacuum", "Hubble")
    for tk in ("CustomBabyIAXO", "Other"):       # doAssert false in the reference (rt:1232-1234, 1347-1348)
        with pytest.raises(rt.SartError) as e:
            rt.newExperimentSetup("CAST", "InGrid2018", "vacuum", tk)
        assert e.value.code == -3


def test_window_vals(rt):
    w, d = rt.calcWindowVals(7.0, 4, 0.838)
    assert w == pytest.approx(0.500418, abs=1e-6) and d == pytest.approx(2.299582, abs=1e-6)
    assert w + d == pytest.approx(14.0 / 5.0)


def test_uniforms_match_oracle(rt, oracle):
    u = np.empty(6)
    for seed, ray in ((299792458, 0), (1, 2**40 + 17), (2**63 + 5, 123456789)):
        rt.lib.sart_ray_uniforms(seed, ray, u.ctypes.data_as(abi.c_double_p))
        assert np.array_equal(u, oracle.ray_uniforms(seed, ray))


def test_no_cpu_fallback(rt):
    """Without a CUDA device every compute entry point must fail with SART_ERR_CUDA, not fall back."""
    if rt.lib.sart_device_count() > 0:
        pytest.skip("a GPU is visible here")
    from helpers import make_config
    setup, tb = make_config("cast_llnl")
    with pytest.raises(rt.SartError) as e:
        rt.RayTracer(rt.FullRaytraceSetup(setup, tb))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    from solaraxionraytracing_b200 import tables
    with pytest.raises(rt.SartError) as e:
        rt.buildCdfs(tables.synthetic_emission(8, 8))
    assert e.value.code == -2


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (tests, smoke() and bench.py's CPU legs only)."""
    pkg = ROOT / "solaraxionraytracing_b200"
    files = [f for ext in ("*.py", "*.cu", "*.cpp", "*.h", "*.cuh") for f in pkg.rglob(ext)]
    assert len(files) > 10
    for f in files:
        txt = f.read_text()
        for needle in ("import oracle", "from oracle", "liboracle", "oracle/"):
            assert needle not in txt, (f, needle)


def test_integer_cdf_thresholds_are_exact(rt):
    """sart_cdf_thresholds: `cdf[i] < (w + 0.5) 2^-32` <=> `w >= thr[i]`, checked with exact rational arithmetic at the
    words around every threshold, and lowerBound == number of thresholds <= w on a realistic CDF row."""
    from fractions import Fraction
    rng = np.random.default_rng(7)
    cdf = np.sort(np.concatenate([rng.random(200), [0.0, 1.0, 2.0 ** -33, 1 - 2.0 ** -33, 0.5, 0.25 + 2.0 ** -34,
                                                        3.0 / 2 ** 32, 3.5 / 2 ** 32, (2 ** 32 - 1.5) / 2 ** 32]]))
    thr = np.zeros(cdf.size, dtype=np.uint32)
    rt.lib.sart_cdf_thresholds(cdf.ctypes.data_as(abi.c_double_p), cdf.size, thr.ctypes.data_as(C.POINTER(C.c_uint32)))
    for c, t in zip(cdf, thr):
        t = int(t)
        for w in {max(t - 1, 0), t, min(t + 1, 2 ** 32 - 2)}:
            if w >= 2 ** 32 - 1:
                continue                                      # the saturated word is resolved on the f64 table
            below = Fraction(float(c)) < Fraction(2 * w + 1, 2 ** 33)
            assert below == (w >= t) or t == 2 ** 32 - 1, (c, t, w)
    nan = np.array([np.nan]); out = np.zeros(1, dtype=np.uint32)
    rt.lib.sart_cdf_thresholds(nan.ctypes.data_as(abi.c_double_p), 1, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert out[0] == 0xFFFFFFFF                               # empty shells (0/0 rows, rt:2674) are never selected
    from solaraxionraytracing_b200 import tables
    from oracle import oracle as orc
    em = tables.synthetic_emission(40, 300, "abc")
    _, dc = orc.build_cdfs(em.radii, em.energies, em.emRates)
    row = np.ascontiguousarray(dc[7])
    th = np.zeros(row.size, dtype=np.uint32)
    rt.lib.sart_cdf_thresholds(row.ctypes.data_as(abi.c_double_p), row.size, th.ctypes.data_as(C.POINTER(C.c_uint32)))
    w = rng.integers(0, 2 ** 32 - 1, size=200_000, dtype=np.uint64)
    u = (w.astype(np.float64) + 0.5) / 4294967296.0          # exact in f64
    want = np.searchsorted(row, u, side="left")               # std/algorithm lowerBound
    got = np.searchsorted(th.astype(np.uint64), w, side="right")   # number of thresholds <= w
    assert np.array_equal(got, want)


@pytest.mark.parametrize("tel", [abi.TK_LLNL, abi.TK_XMM, abi.TK_ABRIXAS])
def test_shell_lookup_table_equals_the_scan(rt, tel):
    """sart_shell_lookup: the radial table the FP32 kernels use for the shell search (rt:1932-1957: hit shell, glass front
    rt:1942-1944, outside the last shell rt:1934) gives the outcome of the scan it replaces for every radius: at each
    boundary (R1[j], R1[j] + thickness[j]) and the FP32 numbers next to it, at the bucket edges of the table, on a dense
    sweep, and for NaN / zero / huge arguments."""
    setup = rt.newExperimentSetup(abi.ES_BABYIAXO, abi.DK_INGRIDIAXO, abi.SK_VACUUM, tel, 0)
    t = setup.telescope
    n_sh = t.nShells
    r1 = np.array(t.allR1[:n_sh], dtype=np.float64)
    th = np.array(t.allThickness[:n_sh], dtype=np.float64)
    pts = []
    for b in np.concatenate([r1, r1 + th]).astype(np.float32):
        x = b
        for _ in range(4):
            x = np.nextafter(x, np.float32(-np.inf))
        for _ in range(9):
            pts.append(x)
            x = np.nextafter(x, np.float32(np.inf))
    rng = np.random.default_rng(5)
    lo, hi = float(r1[0]) - 5.0, float(r1[-1]) + 5.0
    pts = np.concatenate([np.array(pts, dtype=np.float32), np.linspace(lo, hi, 400_001).astype(np.float32),
                          rng.uniform(lo, hi, 200_000).astype(np.float32),
                          np.array([np.nan, 0.0, 1e-30, 1e9, np.inf], dtype=np.float32)])   # rho is a norm: >= 0 or NaN
    a = np.zeros(pts.size, dtype=np.int32)
    b = np.zeros(pts.size, dtype=np.int32)
    rc = rt.lib.sart_shell_lookup(C.byref(setup), pts.size, pts.ctypes.data_as(C.POINTER(C.c_float)),
                                  a.ctypes.data_as(C.POINTER(C.c_int32)), b.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, rt.lib.sart_last_error()
    assert np.array_equal(a, b), pts[a != b][:10]
    # the outcomes are the expected mix: every shell is hit, glass fronts, outside, and NaN = no mirror hit
    assert set(range(n_sh)) <= set(b.tolist())
    assert (b == 64 + 7).sum() > 0 and (b == 64 + 6).sum() > 0
    assert b[-5] == 64 + 9


def _alias_realized_counts(entries, n):
    """Number of 32-bit words that end at each index under the kernel's lookup (k = (w n) >> 32, coin = low 32 bits of
    w n, index k if coin < share else alias), in exact integer arithmetic."""
    two32 = 1 << 32
    got = [0] * n
    t_alias = [0] * n
    for k in range(n):
        e = int(entries[k])
        share, al = e & 0xFFFFF800, e & 0x7FF
        wlo, whi = (k * two32 + n - 1) // n, ((k + 1) * two32 + n - 1) // n      # words of bucket k
        keep = min(max((k * two32 + share + n - 1) // n - wlo, 0), whi - wlo)    # coin = w n - k 2^32 < share
        got[k] += keep
        got[al] += (whi - wlo) - keep
        if al != k:
            t_alias[al] += 1
    return got, t_alias


def test_alias_table_realizes_the_threshold_distribution(rt):
    """sart_alias_table: the alias sampler draws index i for (thr[i] - thr[i-1]) of the 2^32 words, like the inverse-CDF
    search — up to the flooring of a bucket's share to 2^-21 and one word per bucket: |words(i) - c(i)| <= (1 + number of
    buckets whose alias is i) (2048 / n + 2), total variation below 2e-6, on a solar-model row, a one-point distribution,
    a uniform one, n = 1 and n = 2048."""
    from solaraxionraytracing_b200 import tables
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    em = tables.synthetic_emission(40, 1500, "abc")
    rc_, dc = orc.build_cdfs(em.radii, em.energies, em.emRates)
    cases = [np.ascontiguousarray(dc[7]), np.ascontiguousarray(dc[35]), np.ascontiguousarray(rc_),
             np.concatenate([np.zeros(100), np.ones(200)]), np.arange(1, 301) / 300.0, np.array([1.0]),
             np.sort(np.concatenate([rng.random(2047), [1.0]]))]
    two32 = 1 << 32
    for cdf in cases:
        n = cdf.size
        thr = np.zeros(n, dtype=np.uint32)
        rt.lib.sart_cdf_thresholds(cdf.ctypes.data_as(abi.c_double_p), n, thr.ctypes.data_as(C.POINTER(C.c_uint32)))
        ent = np.zeros(n, dtype=np.uint32)
        rt.lib.sart_alias_table(thr.ctypes.data_as(C.POINTER(C.c_uint32)), n, ent.ctypes.data_as(C.POINTER(C.c_uint32)))
        t = np.maximum.accumulate(thr.astype(np.int64))
        edges = np.concatenate([[0], t[:-1], [two32]])
        want = np.diff(edges)                                   # c(i): words the inverse-CDF search maps to i
        assert want.sum() == two32 and (want >= 0).all()
        got, t_alias = _alias_realized_counts(ent, n)
        assert sum(got) == two32
        assert all((int(e) & 0x7FF) < n for e in ent)
        for i in range(n):
            assert abs(got[i] - int(want[i])) <= (1 + t_alias[i]) * (2048 / n + 2), (n, i, got[i], int(want[i]))
            if want[i] == 0:
                assert got[i] <= t_alias[i] * 3                 # an index without words stays (all but) unreachable
        assert sum(abs(g - int(w)) for g, w in zip(got, want)) / (2 * two32) < 2e-6
    big = np.zeros(4096, dtype=np.uint32)                        # n > 2048: refused (entries zeroed)
    out = np.ones(4096, dtype=np.uint32)
    rt.lib.sart_alias_table(big.ctypes.data_as(C.POINTER(C.c_uint32)), 4096, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert not out.any()


def test_shell_lookup_refuses_unresolvable_shells(rt):
    """A glass 1e-5 mm thick puts two boundaries into one bucket of the (at most 4000-bucket) radial table: the helper
    reports it instead of building a wrong table (the FP32 pipeline then refuses the setup, modes 1 and 0 trace it)."""
    setup = rt.newExperimentSetup(abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, 0)
    setup.telescope.allThickness[3] = 1e-5
    rho = np.array([70.0], dtype=np.float32)
    a = np.zeros(1, dtype=np.int32); b = np.zeros(1, dtype=np.int32)
    rc = rt.lib.sart_shell_lookup(C.byref(setup), 1, rho.ctypes.data_as(C.POINTER(C.c_float)),
                                  a.ctypes.data_as(C.POINTER(C.c_int32)), b.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc != 0 and b"radial lookup table" in rt.lib.sart_last_error()
