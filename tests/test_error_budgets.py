"""CPU tests of the host logic added with the bit-exact classification of precision mode 2: the
error budgets of the FP32 decisions (sart_error_budgets: derive_tolerances as the kernels use it), the flag that skips the
pipe tests where no solar ray can reach the pipe wall, and the command line's refusal to substitute input tables
silently."""
import ctypes as C

import pytest

from oracle import ref_setup
from solaraxionraytracing_b200 import abi


@pytest.fixture(scope="module")
def rt():
    from solaraxionraytracing_b200 import raytracer
    return raytracer


def _budgets(rt, setup, scale=1.0, slope_sum=1e-3, rs=0.2, n_radii=1968):
    lat, det, free = C.c_double(), C.c_double(), C.c_int()
    rt.check(rt.lib.sart_error_budgets(C.byref(setup), n_radii, scale, slope_sum, rs, C.byref(lat), C.byref(det), C.byref(free)))
    return lat.value, det.value, free.value


def test_budgets_are_small_positive_and_scale_linearly(rt):
    llnl = ref_setup.make_setup(abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, 0)
    xmm = ref_setup.make_setup(abi.ES_BABYIAXO, abi.DK_INGRIDIAXO, abi.SK_VACUUM, abi.TK_XMM, 0)
    for setup in (llnl, xmm):
        lat, det, _ = _budgets(rt, setup)
        # far above FP32 rounding of a coordinate (6e-8 x 500 mm), far below the shell spacing (2 mm) and a strip (0.5 mm)
        assert 3e-5 < lat < 5e-3 and 3e-5 < det < 5e-2, (lat, det)
        lat2, det2, _ = _budgets(rt, setup, scale=0.25)
        assert lat2 == pytest.approx(0.25 * lat, rel=1e-5) and det2 == pytest.approx(0.25 * det, rel=1e-5)
        assert _budgets(rt, setup, scale=0.0)[:2] == (0.0, 0.0)
        # the reference's rounding noise grows with the slopes (it is one ulp of the emission point's coordinates)
        assert _budgets(rt, setup, slope_sum=5e-3)[0] > lat and _budgets(rt, setup, slope_sum=5e-3)[1] > det
    # the longer optic carries direction errors further: 7.5 m focal length against 1.5 m
    assert _budgets(rt, xmm)[1] > 2.0 * _budgets(rt, llnl)[1]


def test_pipe_tests_are_skipped_only_where_no_solar_ray_can_reach_the_wall(rt):
    """CAST: bore 21.5 mm, pipes 39.9 mm, slopes <= 4.7e-3 over 735 mm: unreachable. BabyIAXO: bore 500 mm, pipes 370 mm."""
    cast = ref_setup.make_setup(abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, 0)
    iaxo = ref_setup.make_setup(abi.ES_BABYIAXO, abi.DK_INGRIDIAXO, abi.SK_VACUUM, abi.TK_XMM, 0)
    assert _budgets(rt, cast)[2] == 1
    assert _budgets(rt, iaxo)[2] == 0
    xray = ref_setup.make_setup(abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, abi.CF_XRAY_TEST)
    xray.testSource.active = 1
    assert _budgets(rt, xray)[2] == 0        # the X-ray source is not bounded by the solar disc
    narrow = ref_setup.make_setup(abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, 0)
    narrow.pipes.cb2vt3_radius = 24.0        # 21.5 + 4.7e-3 * 735 * 1.5 + 1 > 24
    assert _budgets(rt, narrow)[2] == 0


def test_command_line_refuses_silent_stand_ins(rt, tmp_path):
    """The reference stops when a [Resources] file is missing (rt:2647, 1174); so does the command line, before it touches
    the GPU, unless --allowSynthetic is given."""
    from solaraxionraytracing_b200 import config
    from solaraxionraytracing_b200.__main__ import main
    cfg = tmp_path / "config.toml"
    cfg.write_text(config.DEFAULT_CONFIG.read_text().replace('outputPath = "../out"', f'outputPath = "{tmp_path}/out"'))
    with pytest.raises(SystemExit) as e:
        main(["--config", str(cfg), "--nRays", "1000"])
    assert "allowSynthetic" in str(e.value) and "solar model" in str(e.value)


def _supported(rt, setup):
    why = C.create_string_buffer(256)
    rc = rt.lib.sart_throughput_supported(C.byref(setup), why, len(why))
    return rc, why.value.decode()


def test_throughput_modes_cover_every_reflectivity_kind_and_hole_type(rt):
    """sart_throughput_supported (host-only): since round 2 precision modes 1 and 2 take rkEffectiveArea (rt:1553-1562) and
    every XMM hole type (rt:1674-1688); they still refuse shells that overlap or are out of order and more than 64 holes."""
    for tel in (abi.TK_LLNL, abi.TK_XMM, abi.TK_ABRIXAS):
        exp = abi.ES_CAST if tel == abi.TK_LLNL else abi.ES_BABYIAXO
        setup = ref_setup.make_setup(exp, abi.DK_INGRID2018, abi.SK_VACUUM, tel, 0)
        assert _supported(rt, setup) == (1, "")
        setup.telescope.reflKind = abi.RK_EFFECTIVE_AREA
        assert _supported(rt, setup) == (1, "")
    xmm = ref_setup.make_setup(abi.ES_BABYIAXO, abi.DK_INGRIDIAXO, abi.SK_VACUUM, abi.TK_XMM, 0)
    for hole in (abi.HT_CROSS, abi.HT_STAR, abi.HT_CIRCLE, abi.HT_SQUARE, abi.HT_DIAMOND):
        xmm.telescope.holeType, xmm.telescope.numberOfHoles, xmm.telescope.holeInOptics = hole, 5, 2.0
        assert _supported(rt, xmm) == (1, "")
    xmm.telescope.numberOfHoles = 65
    rc, why = _supported(rt, xmm)
    assert rc == 0 and "64 holes" in why
    llnl = ref_setup.make_setup(abi.ES_CAST, abi.DK_INGRID2018, abi.SK_VACUUM, abi.TK_LLNL, 0)
    llnl.telescope.allR1[3] = llnl.telescope.allR1[2]      # two shells at the same radius
    rc, why = _supported(rt, llnl)
    assert rc == 0 and "shell" in why
    llnl.abi_version = 0
    assert rt.lib.sart_throughput_supported(C.byref(llnl), None, 0) < 0
