"""Input tables of the ray-tracing path as numpy arrays + the ctypes view libsart consumes.

The reference reads these in initReflectivity (src/raytracer.nim:1160-1249), newDetectorSetup (rt:1498-1527) and
initFullSetup (rt:2647-2668). Three of its files are not shipped, so besides loaders for the
reference's file formats this module has deterministic synthetic generators of the same shapes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import abi

DATA_DIR = Path(__file__).resolve().parent / "data"


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(abi.c_double_p)


@dataclass
class EmissionTable:
    """`Radius, Energy [keV], emRates` of solar_model_dataframe.csv (rt:2647-2668) in dense form."""
    radii: np.ndarray      # [nR] fraction of the solar radius, ascending
    energies: np.ndarray   # [nE] keV, ascending
    emRates: np.ndarray    # [nR, nE]


@dataclass
class TableSet:
    energies: np.ndarray
    fluxRadiusCDF: np.ndarray
    diffFluxCDFs: np.ndarray          # [nR, nE]
    reflectivity: np.ndarray | None   # [nCoat, nAng, nEn]
    angleLim: tuple[float, float] = (0.0, 1.5)
    reflEnergyLim: tuple[float, float] = (0.03, 15.0)
    strongback: tuple[np.ndarray, np.ndarray] = None   # (E keV, T)
    window: tuple[np.ndarray, np.ndarray] = None
    gasAbsorption: tuple[np.ndarray, np.ndarray] = None
    telescopeTransmission: tuple[np.ndarray, np.ndarray] | None = None
    _keep: list = field(default_factory=list, repr=False)

    def c_struct(self) -> abi.Tables:
        """The sart_tables_t view. The returned struct borrows this object's arrays."""
        t = abi.Tables()
        self._keep.clear()

        def hold(a):
            a = _f64(a)
            self._keep.append(a)
            return a

        if self.energies is not None:
            en, rc, dc = hold(self.energies), hold(self.fluxRadiusCDF), hold(self.diffFluxCDFs)
            assert dc.shape == (rc.size, en.size)
            t.nRadii, t.nEnergies = rc.size, en.size
            t.energies, t.fluxRadiusCDF, t.diffFluxCDFs = _ptr(en), _ptr(rc), _ptr(dc)
        if self.reflectivity is not None:
            r = hold(self.reflectivity)
            assert r.ndim == 3
            t.nCoatings, t.nAngles, t.nReflEnergies = r.shape
            t.reflectivity = _ptr(r)
        t.angleMin, t.angleMax = self.angleLim
        t.reflEnergyMin, t.reflEnergyMax = self.reflEnergyLim
        for name, pair in (("strongbackTransmission", self.strongback), ("windowTransmission", self.window),
                           ("gasAbsorption", self.gasAbsorption), ("telescopeTransmission", self.telescopeTransmission)):
            if pair is None:
                continue
            x, y = hold(pair[0]), hold(pair[1])
            assert x.shape == y.shape and x.ndim == 1
            it = getattr(t, name)
            it.n, it.x, it.y = x.size, _ptr(x), _ptr(y)
        return t


# ---------------------------------------------------------------------------------------------------------------
# Detector chain (rt:1498-1527)

def _read_tsv(path: Path) -> tuple[np.ndarray, np.ndarray]:
    a = np.loadtxt(path, skiprows=1)
    return a[:, 0].copy(), a[:, 1].copy()


def detector_tables_from_resources(resources: str | Path, windowThickness=0.3, alThickness=0.02):
    """Reads the reference's four Henke TSV files and combines them exactly as newDetectorSetup does:
    strongback = Si(200 um) * Al, window = Si3N4 * Al, gas absorption = 1 - T_Ar; energies eV -> keV."""
    res = Path(resources)
    eSiN, tSiN = _read_tsv(res / f"Si3N4Density=3.44Thickness={windowThickness:.1f}microns.tsv")
    eSi, tSi = _read_tsv(res / "SiDensity=2.33Thickness=200.microns.tsv")
    eAr, tAr = _read_tsv(res / "transmission-argon-30mm-1050mbar-295K.tsv")
    eAl, tAl = _read_tsv(res / f"AlDensity=2.7Thickness={alThickness:.2f}microns.tsv")
    return {"strongback": (eSi / 1000.0, tSi * tAl), "window": (eSiN / 1000.0, tSiN * tAl),
            "gasAbsorption": (eAr / 1000.0, 1.0 - tAr)}


def detector_tables_packaged():
    """The same three tables from the packaged fixture (made by tools/make_fixtures.py from the reference's
    resources/*.tsv; the GPU box has no /root/reference)."""
    z = np.load(DATA_DIR / "detector_tables.npz")
    return {"strongback": (z["sb_E"], z["sb_T"]), "window": (z["wd_E"], z["wd_T"]),
            "gasAbsorption": (z["ga_E"], z["ga_A"])}


# ---------------------------------------------------------------------------------------------------------------
# Synthetic stand-ins for the un-shipped tables (shapes of SURVEY.md §8a)

def synthetic_emission(nRadii: int = 1968, nEnergies: int = 1500, kind: str = "abc") -> EmissionTable:
    """Deterministic analytic emission table with the real table's grid: radii 0.0015 + 0.0005 i
    (resources/AGSS09_solar_model_stripped.dat), energies linspace(1e-3, 15, nE) keV
    (src/readOpacityFile.nim:608-609). `abc`: axion-electron-like (bremsstrahlung continuum ~ exp(-E/T) plus a few
    lines), `primakoff`: ~E^2/(exp(E/T)-1) screened. Temperature/density profiles are smooth fits to AGSS09."""
    r = 0.0015 + 0.0005 * np.arange(nRadii, dtype=np.float64)
    E = np.linspace(1e-3, 15.0, nEnergies)
    T = 1.35 * np.exp(-(r / 0.29) ** 1.25) + 0.05           # keV
    rho = np.exp(-r / 0.095)                                # relative density
    Tm, Em = T[:, None], E[None, :]
    if kind == "primakoff":
        ks2 = (8.0 * rho[:, None] ** 0.5 + 0.2) ** 2        # Debye screening scale^2, keV^2
        em = rho[:, None] * Tm * ks2 * Em / np.expm1(Em / Tm) * np.log1p(4.0 * Em * Em / ks2) / (Em * Em + 1e-3)
    else:
        cont = rho[:, None] ** 2 / np.sqrt(Tm) * np.exp(-Em / Tm) / (Em + 0.05)
        lines = np.zeros_like(cont)
        for e0, amp in ((0.653, 0.6), (0.779, 0.5), (0.986, 0.9), (1.865, 0.35), (2.45, 0.25), (6.5, 0.15)):
            lines += amp * np.exp(-0.5 * ((Em - e0) / 0.012) ** 2) * np.exp(-e0 / Tm)
        em = cont * (1.0 + 8.0 * lines)
    em = em * 1e-12
    return EmissionTable(radii=r, energies=E, emRates=np.ascontiguousarray(em))


def synthetic_reflectivity(nCoatings: int = 4, nAngles: int = 1000, nEnergies: int = 1000,
                           angleLim=(0.0, 1.5), energyLim=(0.03, 15.0)) -> np.ndarray:
    """Analytic grazing-incidence reflectivity R(angle deg, E keV) on the reference's grid (1000 angles 0..1.5 deg x
    1000 energies 0.03..15 keV per coating, tools/llnl_layer_reflectivity.nim:50-51): total external reflection below
    a critical angle ~ 1/E, an absorption edge, and a Bragg bump per multilayer recipe."""
    ang = np.linspace(angleLim[0], angleLim[1], nAngles)[:, None]
    E = np.linspace(energyLim[0], energyLim[1], nEnergies)[None, :]
    out = np.empty((nCoatings, nAngles, nEnergies), dtype=np.float64)
    for c in range(nCoatings):
        thc = (0.52 + 0.035 * c) * (8.0 / np.maximum(E, 0.05)) ** 0.95 * 0.1 + 0.02  # critical angle in deg
        x = ang / thc
        ter = 0.97 / (1.0 + x ** 6) ** 0.5 * np.exp(-0.08 * x)
        edge = 1.0 - 0.35 * np.exp(-0.5 * ((E - (2.1 + 0.15 * c)) / 0.12) ** 2)
        d = 3.2 + 0.6 * c                                                          # bilayer period, nm
        bragg_ang = np.degrees(np.arcsin(np.clip(1.2398 / np.maximum(E, 0.05) / (2.0 * d), 0, 1)))
        bump = 0.35 * np.exp(-0.5 * ((ang - bragg_ang) / 0.035) ** 2) * (E > 4.0)
        out[c] = np.clip(ter * edge + bump * (1.0 - ter), 0.0, 1.0)
    return out


def gold_reflectivity_packaged(nAngles: int = 1000, nEnergies: int = 1000, angleLim=(0.0, 1.5),
                               energyLim=(0.03, 15.0)) -> np.ndarray:
    """Gold (0.25 um) grazing-incidence reflectivity on the reference's grid, [1, nAngles, nEnergies].

    The HDF5 the reference reads (rt:1196-1208, gold_0.25microns_reflectivities.h5) is not shipped; what the tree has
    is the Henke download it was made from (resources/reflectivity.zip: 71 angles 0.13..0.83 deg x 500 energies
    30..10000 eV), packed by tools/make_fixtures.py. It is resampled bilinearly (linear in angle and in energy) onto
    the (0..1.5 deg) x (0.03..15 keV) grid of tools/download_henke_files.nim:258-262. Outside the tabulated range:
    below 0.13 deg the reflectivity is continued linearly to R = 1 at grazing angle 0 (total external reflection),
    above 0.83 deg and above 10 keV the edge value is kept."""
    z = np.load(DATA_DIR / "gold_reflectivity_henke.npz")
    a0, e0, R = z["angles_deg"].astype(np.float64), z["energies_eV"].astype(np.float64) / 1000.0, z["R"].astype(np.float64)
    a0 = np.concatenate([[0.0], a0])
    R = np.concatenate([np.ones((1, R.shape[1])), R], axis=0)
    ang = np.linspace(angleLim[0], angleLim[1], nAngles)
    en = np.linspace(energyLim[0], energyLim[1], nEnergies)
    # energy axis first (np.interp clamps at the edges), then the angle axis
    Re = np.stack([np.interp(en, e0, row) for row in R])                       # [72, nEnergies]
    out = np.stack([np.interp(ang, a0, Re[:, j]) for j in range(nEnergies)], axis=1)   # [nAngles, nEnergies]
    return np.ascontiguousarray(out[None, :, :])


@dataclass
class SolarModel:
    """The AGSS09 columns the emission-rate generator reads (readSolarModel.nim:3-7; readOpacityFile.nim:659-700)."""
    radius: np.ndarray          # [nR] fraction of the solar radius (0.0015 + 0.0005 i)
    temp_K: np.ndarray          # [nR]
    rho_gcm3: np.ndarray        # [nR]
    mass_fractions: np.ndarray  # [nR, 29] H1, He4, He3, C12 ... Ni (file order)


def read_solar_model(path: str | Path) -> SolarModel:
    """readSolarModel (src/readSolarModel.nim:3-7) for AGSS09_solar_model_stripped.dat (35 columns, '#' header)."""
    sm = np.loadtxt(path, comments="#")
    if sm.ndim != 2 or sm.shape[1] != 35:
        raise ValueError(f"{path}: expected 35 columns (Mass Radius Temp Rho Pres Lumi + 29 elements)")
    return SolarModel(sm[:, 1].copy(), sm[:, 2].copy(), sm[:, 3].copy(), np.ascontiguousarray(sm[:, 6:35]))


def solar_model_packaged() -> SolarModel:
    """The same columns from the packaged copy (tools/make_fixtures.py; the GPU box has no /root/reference)."""
    z = np.load(DATA_DIR / "agss09_solar_model.npz")
    return SolarModel(z["radius"], z["temp_K"], z["rho_gcm3"], np.ascontiguousarray(z["mass_fractions"]))


def read_solar_model_dataframe(path: str | Path) -> EmissionTable:
    """The emission table the reference reads in initFullSetup (rt:2647-2668): a CSV with the columns `Radius`,
    `Energy [keV]` and `emRates` written by readOpacityFile (one row per (radius, energy), any order). Radii and energies
    are the sorted unique values; every radius must carry the same energies (the reference's doAssert, rt:2667)."""
    with open(path) as f:
        header = [h.strip() for h in f.readline().rstrip("\n").split(",")]
    need = ("Radius", "Energy [keV]", "emRates")
    for n in need:
        if n not in header:
            raise ValueError(f"{path}: column {n!r} missing (found {header})")
    data = np.loadtxt(path, delimiter=",", skiprows=1, usecols=[header.index(n) for n in need], ndmin=2)
    radii, ri = np.unique(data[:, 0], return_inverse=True)
    energies, ei = np.unique(data[:, 1], return_inverse=True)
    if data.shape[0] != radii.size * energies.size:
        raise ValueError(f"{path}: {data.shape[0]} rows is not {radii.size} radii x {energies.size} energies")
    em = np.full((radii.size, energies.size), np.nan)
    em[ri, ei] = data[:, 2]
    if np.isnan(em).any():
        raise ValueError(f"{path}: not every radius has every energy")
    return EmissionTable(radii=radii, energies=energies, emRates=em)


def write_solar_model_dataframe(path: str | Path, table: EmissionTable) -> Path:
    """Writes an EmissionTable in the format readOpacityFile produces (readOpacityFile.nim:848-849 `result.add toDf(...)`)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    R = np.repeat(table.radii, table.energies.size)
    E = np.tile(table.energies, table.radii.size)
    with open(path, "w") as f:
        f.write("Radius,Energy [keV],emRates\n")
        np.savetxt(f, np.column_stack([R, E, table.emRates.reshape(-1)]), delimiter=",", fmt="%.17g")
    return path


def reflectivity_from_h5(path: str | Path, numCoatings: int | None = None):
    """The reflectivity files initReflectivity reads (rt:1160-1232): `/Energy` [keV], `/Angles` [deg] and either
    `/Reflectivity0 .. /Reflectivity{n-1}` (LLNL multilayer recipes, llnl_layer_reflectivities.h5) or `/Reflectivity`
    (gold_0.25microns_reflectivities.h5), each [nAngles, nEnergies]. Returns (reflectivity [nCoat, nAng, nEn],
    (angleMin, angleMax), (energyMin, energyMax)) ready for `TableSet(reflectivity=, angleLim=, reflEnergyLim=)`.
    Needs h5py, which is not part of the build image (the two files are not part of the reference checkout either)."""
    try:
        import h5py
    except ImportError as e:
        raise ImportError("reading the reference's HDF5 reflectivity files needs h5py; alternatively pass the arrays to "
                          "TableSet directly or use gold_reflectivity_packaged() / synthetic_reflectivity()") from e
    with h5py.File(path, "r") as f:
        energies = np.asarray(f["Energy"], dtype=np.float64)
        angles = np.asarray(f["Angles"], dtype=np.float64)
        if "Reflectivity" in f:
            data = [np.asarray(f["Reflectivity"], dtype=np.float64)]
        else:
            n = numCoatings if numCoatings is not None else sum(1 for k in f.keys() if k.startswith("Reflectivity"))
            data = [np.asarray(f[f"Reflectivity{i}"], dtype=np.float64) for i in range(n)]
    refl = np.stack([d.reshape(angles.size, energies.size) for d in data])
    return np.ascontiguousarray(refl), (float(angles.min()), float(angles.max())), (float(energies.min()), float(energies.max()))
