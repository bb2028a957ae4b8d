"""The outputs of a fused GPU run against per-ray records of the same rays: the weighted radial histogram
(sart_enable_radial_hist) reproduces the containment radii generateResultPlots gets by sorting every passed ray
(rt:2459-2527) to within two histogram bins (1.2e-3 mm), and the command line writes the reference's detector-image CSV."""
import numpy as np
import pytest

from helpers import make_config
from solaraxionraytracing_b200 import abi, output

pytestmark = pytest.mark.gpu
SEED = 299792458


@pytest.fixture(scope="module")
def rt():
    from solaraxionraytracing_b200 import raytracer
    if raytracer.lib.sart_device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests need a B200")
    return raytracer


@pytest.mark.parametrize("cfg,mode", [("cast_llnl", 1), ("cast_llnl", 0), ("babyiaxo_xmm", 1)])
def test_radial_hist_gives_the_sorted_rays_radii(rt, cfg, mode):
    setup, tb = make_config(cfg)
    n = 600_000
    with rt.RayTracer(rt.FullRaytraceSetup(setup, tb)) as tr:
        tr.set_precision(mode)
        tr.enable_radial_hist(16384)
        tr.update_setup(setup)            # re-runs the pilot launch: must not leak into the histogram
        tr.set_precision(mode)
        tr.reset_image()
        tr.trace_mc(n, SEED)
        res = tr.read_image()
        edges, hw, hn = tr.read_radial_hist()
        rays = tr.traceAxionWrapper(n, SEED)
        tr.reset_image()
        assert tr.read_radial_hist()[2].sum() == 0
    c = res.counters[0]
    assert int(hn.sum()) == c["n_passed"] == int(rays.passed.sum())
    assert hw.sum() == pytest.approx(c["sum_w"], rel=1e-12)
    exact = output.containment_radii(rays.r[rays.passed], rays.w[rays.passed])
    hist = output.containment_radii_from_hist(edges, hw, hn)
    step = edges[1] - edges[0]
    assert step < 7e-4
    for name in ("rSigma1", "rSigma2", "rSigma1W", "rSigma2W"):
        a, b = getattr(hist, name), getattr(exact, name)
        assert abs(a - b) <= 2 * step, (name, a, b)
    assert 0 < exact.rSigma1 < exact.rSigma2 < 7.0


def test_command_line_writes_the_detector_image(rt, tmp_path, capsys):
    from solaraxionraytracing_b200.__main__ import main
    cfg = tmp_path / "config.toml"
    from solaraxionraytracing_b200 import config
    cfg.write_text(config.DEFAULT_CONFIG.read_text().replace('outputPath = "../out"', f'outputPath = "{tmp_path}/out"'))
    assert main(["--config", str(cfg), "--nRays", "300000", "--ignoreGasAbs", "--allowSynthetic"]) == 0
    out = capsys.readouterr().out
    assert "Passed axions " in out and "The total flux arriving in the detector is: " in out
    img, cols = output.read_axion_image_csv(tmp_path / "out" / "axion_image_BabyIAXO.csv")   # config_default.toml setup
    flux = float(out.split("The total flux arriving in the detector is: ")[1].split()[0])
    assert img.sum() == pytest.approx(flux, rel=1e-12) and flux > 0
    assert cols["yr0"][-1] > 0          # rSigma1W of the run
    assert main(["--config", str(cfg), "--allowSynthetic", "--nRays", "100000", "--xrayTest", "--ignoreDetWindow", "--angularScanMin", "0",
                 "--angularScanMax", "0.04", "--numAngularScanPoints", "3"]) == 0
    scan = np.loadtxt(tmp_path / "out" / "angular_scan_telescope_y.csv", delimiter=",", skiprows=1)
    assert scan.shape == (3, 2) and scan[:, 1].max() == 1.0


def test_command_line_mass_scan(rt, tmp_path, capsys):
    """[Run] mAxion = [...] with the gas stage: one image per axion mass, total flux per mass on stdout."""
    from solaraxionraytracing_b200.__main__ import main
    from solaraxionraytracing_b200 import config
    txt = config.DEFAULT_CONFIG.read_text().replace('outputPath = "../out"', f'outputPath = "{tmp_path}/out"')
    txt = txt.replace('stageSetup = "vacuum"', 'stageSetup = "gas"')
    txt += '\n[Run]\nnRays = 200000\nseed = 3\nmAxion = [0.006, 0.0082, 0.02]\n'
    cfg = tmp_path / "config.toml"
    cfg.write_text(txt)
    assert main(["--config", str(cfg), "--allowSynthetic"]) == 0
    out = capsys.readouterr().out
    assert out.count("m_a = ") == 3
    imgs = np.load(tmp_path / "out" / "axion_images_mass_scan.npy")
    assert imgs.shape == (3, 256, 256) and np.all(imgs.sum(axis=(1, 2)) > 0)
    flux = [float(line.split("total flux")[1]) for line in out.splitlines() if line.startswith("m_a = ")]
    assert np.allclose(flux, imgs.sum(axis=(1, 2)), rtol=1e-5)
    assert max(flux) / min(flux) > 1.5          # the scan crosses the m_a = m_gamma resonance
