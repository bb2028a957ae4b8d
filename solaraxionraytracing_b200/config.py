"""The reference's `config.toml` schema (config/config_default.toml) for the ray-tracing path.

Mirrors the parse procs of src/raytracer.nim:984-1096: `parseSetup` (:1026-1033), `maybeParseMagnetConfig`
(:1035-1055), `maybeParseTestXraySource` (:1057-1080), `maybeParseDetectorInstallation` (:1082-1096) and the
`[Resources]` getters (:988-1024). A table is applied when its CLI flag is set OR its `useConfig` key is true — note
that `--xrayTest` therefore always takes the source from the file, never from the per-experiment defaults of
`initTestXraySource` (:1350-1379). Everything numeric stays in libsart: this module only fills the sart_setup_t POD.
"""
from __future__ import annotations

import tomllib
from dataclasses import dataclass
from pathlib import Path

from . import abi

DEFAULT_CONFIG = Path(__file__).resolve().parent / "data" / "config_default.toml"


@dataclass
class Resources:
    """[Resources] of config.toml (rt:988-1024)."""
    resourcePath: str = "../resources"
    outputPath: str = "../out"
    llnlEfficiency: str = ""
    goldFilePrefix: str = ""
    rawSolarModel: str = ""
    solarModelFile: str = ""
    llnlReflFile: str = ""
    goldReflFile: str = ""


def load_config(path: str | Path | None = None) -> dict:
    """Parses a config.toml; without a path the packaged copy of the reference's config_default.toml is used (the
    reference copies config_default.toml to config.toml on first start, rt:2829-2833)."""
    with open(path or DEFAULT_CONFIG, "rb") as f:
        return tomllib.load(f)


def parseSetup(cfg: dict) -> tuple[str, str, str, str]:
    """parseSetup (rt:1026-1033): (experimentSetup, detectorSetup, stageSetup, telescopeSetup) enum strings.
    Unknown strings raise ValueError like parseEnum."""
    s = cfg["Setup"]
    out = (s["experimentSetup"], s["detectorSetup"], s["stageSetup"], s["telescopeSetup"])
    for v, table in zip(out, (abi.EXPERIMENT_KINDS, abi.DETECTOR_KINDS, abi.STAGE_KINDS, abi.TELESCOPE_KINDS)):
        if v not in table:
            raise ValueError(f"invalid enum value {v!r} in [Setup]; expected one of {sorted(table)}")
    return out


def parseResources(cfg: dict) -> Resources:
    r = cfg.get("Resources", {})
    return Resources(**{k: r[k] for k in Resources.__dataclass_fields__ if k in r})


@dataclass
class Run:
    """[Run] — NOT in the reference, whose values for these are hard-coded `let`s: NumberOfPointsSun (rt:251), the
    `randomize` seed (rt:276) and mAxion (rt:255). Optional table; a strict superset of the reference's schema."""
    nRays: int = 1_000_000
    seed: int = 299792458
    mAxion: tuple = ()          # eV; empty = the reference's 0.0853 eV; several values = a mass scan (gas stage)
    precision: str = "f32"      # f32 | fast | exact
    sampler: str = "inverse_cdf"   # inverse_cdf (the reference's lowerBound) | alias (same distributions, f32 single mass only)


def parseRun(cfg: dict) -> Run:
    r = cfg.get("Run", {})
    m = r.get("mAxion", ())
    if isinstance(m, (int, float)):
        m = (float(m),)
    out = Run(nRays=int(r.get("nRays", 1_000_000)), seed=int(r.get("seed", 299792458)),
              mAxion=tuple(float(x) for x in m), precision=str(r.get("precision", "f32")),
              sampler=str(r.get("sampler", "inverse_cdf")))
    if out.sampler not in ("inverse_cdf", "alias"):
        raise ValueError(f"[Run] sampler = {out.sampler!r}; expected inverse_cdf or alias")
    if out.precision not in ("f32", "fast", "exact"):
        raise ValueError(f"[Run] precision = {out.precision!r}; expected f32, fast or exact")
    if out.nRays < 0 or len(out.mAxion) > abi.MAX_MASSES:
        raise ValueError("[Run] nRays must be >= 0 and at most %d axion masses can be scanned at once" % abi.MAX_MASSES)
    return out


def apply_config(setup: abi.Setup, cfg: dict, flags: int) -> abi.Setup:
    """Overrides the per-experiment defaults in `setup` with the [Magnet], [TestXraySource] and [DetectorInstallation]
    tables where the reference would (flag set or useConfig = true)."""
    m = cfg.get("Magnet", {})
    if (flags & abi.CF_READ_MAGNET_CONFIG) or m.get("useConfig", False):        # rt:1043
        mg = setup.magnet
        mg.B, mg.lengthB, mg.radiusCB = float(m["B"]), float(m["lengthB"]), float(m["radiusCB"])
        mg.lengthColdbore, mg.pGasRoom, mg.tGas = float(m["lengthColdbore"]), float(m["pGasRoom"]), float(m["tGas"])
    x = cfg.get("TestXraySource", {})
    if (flags & abi.CF_XRAY_TEST) or x.get("useConfig", False):                 # rt:1065
        ts = setup.testSource
        ts.active, ts.parallel = int(bool(x["active"])), int(bool(x["parallel"]))
        ts.energy, ts.distance, ts.radius = float(x["energy"]), float(x["distance"]), float(x["radius"])
        ts.offAxisUp, ts.offAxisLeft = float(x["offAxisUp"]), float(x["offAxisLeft"])
        ts.activity, ts.lengthCol = float(x["activity"]), float(x["lengthCol"])
    d = cfg.get("DetectorInstallation", {})
    if (flags & abi.CF_READ_DET_INSTALL_CONFIG) or d.get("useConfig", False):   # rt:1090
        di = setup.detectorInstall
        di.distanceDetectorXRT = float(d["distanceDetectorXRT"])
        di.distanceWindowFocalPlane = float(d["distanceWindowFocalPlane"])
        di.lateralShift, di.transversalShift = float(d["lateralShift"]), float(d["transversalShift"])
    return setup


def setup_from_config(path: str | Path | None = None, flags: int = 0):
    """(sart_setup_t, Resources) for a config file + CLI flag set: parseSetup -> newExperimentSetup/newDetectorSetup
    (C++ constructors in libsart) -> table overrides."""
    from . import raytracer
    cfg = load_config(path)
    es, dk, sk, tk = parseSetup(cfg)
    setup = raytracer.newExperimentSetup(es, dk, sk, tk, flags)
    return apply_config(setup, cfg, flags), parseResources(cfg)
