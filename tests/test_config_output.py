"""Host logic around the GPU path: the config.toml schema (rt:984-1096), the containment radii of generateResultPlots
(rt:2459-2527), the detector-image CSV of plotHeatmap (rt:863-896) and the solar_model_dataframe.csv format
(rt:2647-2668). No GPU."""
import numpy as np
import pytest

from solaraxionraytracing_b200 import abi, config, output, tables


def test_default_config_is_the_reference_default():
    cfg = config.load_config()
    assert config.parseSetup(cfg) == ("BabyIAXO", "InGridIAXO", "vacuum", "XMM")      # config_default.toml:20-23
    res = config.parseResources(cfg)
    assert res.solarModelFile == "solar_model_dataframe.csv" and res.goldReflFile == "gold_0.25microns_reflectivities.h5"
    for table in ("Magnet", "TestXraySource", "DetectorInstallation"):
        assert cfg[table]["useConfig"] is False
    assert cfg["Magnet"]["radiusCB"] == 350.0 and cfg["TestXraySource"]["lengthCol"] == 0.021


def test_tables_apply_only_with_flag_or_useconfig():
    setup, _ = config.setup_from_config(None, 0)
    assert setup.magnet.radiusCB == 500.0 and setup.detectorInstall.distanceDetectorXRT == 7500.0   # rt:1117, 1402
    assert setup.testSource.active == 0
    setup, _ = config.setup_from_config(None, abi.CF_READ_MAGNET_CONFIG | abi.CF_READ_DET_INSTALL_CONFIG)
    assert setup.magnet.radiusCB == 350.0 and setup.magnet.lengthColdbore == 11300.0
    assert setup.detectorInstall.distanceDetectorXRT == 1485.0
    # --xrayTest takes the source from the file, not from initTestXraySource's per-experiment defaults (rt:1065)
    setup, _ = config.setup_from_config(None, abi.CF_XRAY_TEST)
    assert setup.testSource.active == 1 and setup.testSource.parallel == 0
    assert setup.testSource.energy == 1.0 and setup.testSource.lengthCol == 0.021


def test_useconfig_and_bad_enum(tmp_path):
    txt = config.DEFAULT_CONFIG.read_text().replace('experimentSetup = "BabyIAXO"', 'experimentSetup = "CAST"')
    txt = txt.replace('telescopeSetup = "XMM"', 'telescopeSetup = "LLNL"').replace('detectorSetup = "InGridIAXO"', 'detectorSetup = "InGrid2018"')
    txt = txt.replace("[Magnet]\nuseConfig = false", "[Magnet]\nuseConfig = true").replace("B = 2.0", "B = 8.8")
    p = tmp_path / "config.toml"
    p.write_text(txt)
    setup, _ = config.setup_from_config(p, 0)
    assert setup.experiment == abi.ES_CAST and setup.telescope.kind == abi.TK_LLNL
    assert setup.magnet.B == 8.8 and setup.magnet.radiusCB == 350.0
    p.write_text(txt.replace('"LLNL"', '"Chandra"'))
    with pytest.raises(ValueError):
        config.setup_from_config(p, 0)


def _reference_radii(R, W):
    """Literal loop of rt:2462-2510."""
    order = np.argsort(R, kind="stable")
    pointR, w = R[order], W[order]
    n = len(pointR)
    rnd = lambda x: int(np.floor(x + 0.5))
    r1, r2 = pointR[rnd(n * 0.68) - 1], pointR[rnd(n * 0.955) - 1]
    tot = w.sum()
    s1, s2 = tot * 0.68, tot * 0.955
    k0 = rnd(n * 0.63)
    ws = w[:k0 + 1].sum()
    r1w = r2w = 0.0
    for i in range(k0 + 1, n):
        ws += w[i]
        if ws < s1:
            r1w = pointR[i]
        elif ws < s2 and ws >= s1:
            r2w = pointR[i]
    return r1, r2, r1w, r2w


def test_containment_radii_match_literal_loop_and_histogram():
    rng = np.random.default_rng(5)
    n = 20000
    R = np.abs(rng.normal(0, 1.2, n)) + 0.3 * rng.random(n)
    W = rng.random(n) * np.exp(-R)           # weights fall with radius: weighted radii are smaller than the count ones
    got = output.containment_radii(R, W)
    want = _reference_radii(R, W)
    assert (got.rSigma1, got.rSigma2) == want[:2]
    assert got.rSigma1W == pytest.approx(want[2], rel=1e-12) and got.rSigma2W == pytest.approx(want[3], rel=1e-12)
    assert got.rSigma1W == 0.0               # 68 % of the weight is reached before the 63 % count index: stays 0 (quirk)
    W2 = rng.random(n) * (0.2 + R)           # weights rising with radius
    got2, want2 = output.containment_radii(R, W2), _reference_radii(R, W2)
    assert got2.rSigma1W == pytest.approx(want2[2], rel=1e-12) and got2.rSigma1W > 0
    edges = np.linspace(0, 10, 16385)
    hw, _ = np.histogram(R, edges, weights=W2)
    hn, _ = np.histogram(R, edges)
    h = output.containment_radii_from_hist(edges, hw, hn)
    step = edges[1]
    for a, b in zip((h.rSigma1, h.rSigma2, h.rSigma1W, h.rSigma2W), (got2.rSigma1, got2.rSigma2, got2.rSigma1W, got2.rSigma2W)):
        assert abs(a - b) <= 2 * step


def test_axion_image_csv_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    img = rng.random((256, 256)) * 1e-3
    path = output.write_axion_image_csv(tmp_path / "out" / "axion_image_2018.csv", img, 2.5, 4.0)
    lines = path.read_text().splitlines()
    assert lines[0] == "x,y,photon flux,yr0,yr02,x-position [mm],y-position [mm],xr,xrneg,yr,xr2,xrneg2,yr2"
    assert len(lines) == 1 + 65536
    back, cols = output.read_axion_image_csv(path)
    assert np.array_equal(back, img)
    k = 300                                               # row = y * 256 + x
    assert cols["x"][k] == 300 % 256 and cols["y"][k] == 1
    assert cols["photon flux"][k] == img[1, 44]
    assert cols["x-position [mm]"][k] == pytest.approx(44 * 14.0 / 256)
    assert cols["yr0"][0] == -2.5 and cols["yr0"][-1] == 2.5 and cols["yr02"][-1] == 4.0
    assert cols["xr"][0] == pytest.approx(7.0) and cols["yr2"][0] == pytest.approx(3.0)
    mid = 65536 // 2
    assert cols["xr"][mid] == pytest.approx(7.0 + np.sqrt(2.5 ** 2 - cols["yr0"][mid] ** 2))
    short = output.write_axion_image_csv(tmp_path / "p4.csv", img, 2.5, 4.0, precision=4)
    assert len(short.read_text().splitlines()[1].split(",")[2]) <= 10


def test_result_summary_lines():
    c = {"n_passed": 4, "n_passed_till_window": 6, "n_hit_nickel": 1, "sum_x": 28.0, "sum_y": 26.0, "sum_r": 2.0}
    s = output.result_summary(c, output.ContainmentRadii(1, 2, 0.5, 1.5), np.ones((2, 2)))
    assert s.splitlines()[:3] == ["Passed axions 4", "Passed axions until the Window 6", "Number of X-rays hitting nickel: 1"]
    assert "7.0" in s.splitlines()[3] and "0.5vs 1" in s and s.endswith("The total flux arriving in the detector is: 4.0")


def test_solar_model_dataframe_roundtrip(tmp_path):
    em = tables.synthetic_emission(12, 9, "primakoff")
    p = tables.write_solar_model_dataframe(tmp_path / "solar_model_dataframe.csv", em)
    assert p.read_text().splitlines()[0] == "Radius,Energy [keV],emRates"
    back = tables.read_solar_model_dataframe(p)
    assert np.array_equal(back.radii, em.radii) and np.array_equal(back.energies, em.energies)
    assert np.array_equal(back.emRates, em.emRates)
    rows = p.read_text().splitlines()
    (tmp_path / "shuffled.csv").write_text("\n".join([rows[0]] + rows[:0:-1]) + "\n")     # any row order
    assert np.array_equal(tables.read_solar_model_dataframe(tmp_path / "shuffled.csv").emRates, em.emRates)
    (tmp_path / "short.csv").write_text("\n".join(rows[:-1]) + "\n")
    with pytest.raises(ValueError):
        tables.read_solar_model_dataframe(tmp_path / "short.csv")


def test_cli_parser_has_the_reference_switches():
    from solaraxionraytracing_b200.__main__ import build_parser
    a = build_parser().parse_args(["--ignoreDetWindow", "--xrayTest", "--angularScanMax", "0.3", "--numAngularScanPoints", "14"])
    assert a.ignoreDetWindow and a.xrayTest and not a.magnet and a.angularScanMax == 0.3 and a.numAngularScanPoints == 14


def test_run_table_is_an_optional_superset(tmp_path):
    assert config.parseRun(config.load_config()) == config.Run()          # absent in the reference's default file
    p = tmp_path / "config.toml"
    p.write_text(config.DEFAULT_CONFIG.read_text() + '\n[Run]\nnRays = 5000000\nseed = 7\nmAxion = [0.01, 0.02]\nprecision = "exact"\n')
    r = config.parseRun(config.load_config(p))
    assert (r.nRays, r.seed, r.mAxion, r.precision) == (5_000_000, 7, (0.01, 0.02), "exact")
    p.write_text(config.DEFAULT_CONFIG.read_text() + '\n[Run]\nprecision = "fp8"\n')
    with pytest.raises(ValueError):
        config.parseRun(config.load_config(p))
    # the sampler: the reference's inverse CDF unless asked otherwise
    assert r.sampler == "inverse_cdf"
    p.write_text(config.DEFAULT_CONFIG.read_text() + '\n[Run]\nsampler = "alias"\n')
    assert config.parseRun(config.load_config(p)).sampler == "alias"
    p.write_text(config.DEFAULT_CONFIG.read_text() + '\n[Run]\nsampler = "sobol"\n')
    with pytest.raises(ValueError):
        config.parseRun(config.load_config(p))
    from solaraxionraytracing_b200.__main__ import build_parser
    assert build_parser().parse_args([]).sampler is None and build_parser().parse_args(["--sampler", "alias"]).sampler == "alias"


def test_h5_reflectivity_reader_fails_clearly_without_h5py(tmp_path):
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            tables.reflectivity_from_h5(tmp_path / "llnl_layer_reflectivities.h5")
    else:
        import h5py
        p = tmp_path / "gold.h5"
        with h5py.File(p, "w") as f:
            f["Energy"] = np.linspace(0.03, 15, 7); f["Angles"] = np.linspace(0, 1.5, 5)
            f["Reflectivity"] = np.arange(35.0).reshape(5, 7)
        r, al, el = tables.reflectivity_from_h5(p)
        assert r.shape == (1, 5, 7) and al == (0.0, 1.5) and el == (0.03, 15.0)
