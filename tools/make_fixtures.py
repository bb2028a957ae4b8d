#!/usr/bin/env python
"""Regenerates the small data fixtures this repo ships, from the reference checkout (default /root/reference).

The GPU box has no /root/reference, so the detector-chain tables the reference reads from resources/*.tsv
(src/raytracer.nim:1498-1527) are packed into solaraxionraytracing_b200/data/detector_tables.npz here, combined
exactly as newDetectorSetup combines them. Run in the build container:  python tools/make_fixtures.py
"""
import sys
import zipfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from solaraxionraytracing_b200 import tables  # noqa: E402


def main(ref: str = "/root/reference") -> None:
    res = Path(ref) / "resources"
    out = ROOT / "solaraxionraytracing_b200" / "data"
    out.mkdir(exist_ok=True)
    t = tables.detector_tables_from_resources(res)
    np.savez_compressed(out / "detector_tables.npz",
                        sb_E=t["strongback"][0], sb_T=t["strongback"][1],
                        wd_E=t["window"][0], wd_T=t["window"][1],
                        ga_E=t["gasAbsorption"][0], ga_A=t["gasAbsorption"][1])
    print("wrote", out / "detector_tables.npz")

    # AGSS09 solar model (resources/AGSS09_solar_model_stripped.dat, read by readSolarModel.nim:3-7): the columns the
    # emission-rate generator needs — Radius, Temp, Rho and the 29 mass fractions H1..Ni in file order.
    sm = np.loadtxt(res / "AGSS09_solar_model_stripped.dat", comments="#")
    assert sm.shape == (1968, 35)
    np.savez_compressed(out / "agss09_solar_model.npz", radius=sm[:, 1], temp_K=sm[:, 2], rho_gcm3=sm[:, 3],
                        mass_fractions=sm[:, 6:35])
    print("wrote", out / "agss09_solar_model.npz")

    # Gold reflectivity (Henke, 0.25 um Au): 71 angle files x 500 energies inside resources/reflectivity.zip
    # (the HDF5 the reference reads, rt:1196-1208, is a missing blob). Packed as float32 [71, 500].
    zf = zipfile.ZipFile(res / "reflectivity.zip")
    rows = {}
    for name in zf.namelist():
        base = Path(name).name
        if base.endswith("degGold0.25microns"):
            ang = float(base.split("deg")[0])
            txt = zf.read(name).decode("latin-1").splitlines()
            vals = []
            for line in txt:
                p = line.split()
                if len(p) >= 2:
                    try:
                        vals.append((float(p[0]), float(p[1])))
                    except ValueError:
                        pass
            rows[ang] = np.array(vals)
    angs = np.array(sorted(rows))
    E = rows[angs[0]][:, 0]
    R = np.stack([rows[a][:, 1] for a in angs])
    assert all(np.allclose(rows[a][:, 0], E) for a in angs)
    np.savez_compressed(out / "gold_reflectivity_henke.npz", angles_deg=angs, energies_eV=E, R=R.astype(np.float32))
    print("wrote", out / "gold_reflectivity_henke.npz", R.shape)


if __name__ == "__main__":
    main(*sys.argv[1:])
