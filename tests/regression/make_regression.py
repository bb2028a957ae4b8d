#!/usr/bin/env python
"""Generates tests/regression/oracle_rays_v1.npz: per-ray outputs of the CPU oracle for 2000 Philox rays of each test
configuration. NOT golden data: the reference ships no per-ray vectors and cannot be executed here (Nim), so this
fixture is the oracle's own output and only guards the oracle against accidental changes (what pins the oracle to the
reference is tests/test_oracle_known_answers.py). Run:  python tests/regression/make_regression.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_config  # noqa: E402
from oracle import oracle as orc  # noqa: E402

meta = {"configs": ["cast_llnl", "babyiaxo_xmm", "babyiaxo_gas", "cast_abrixas"], "first_ray": 0, "n": 2000,
        "seed": 299792458}
out = {"meta": json.dumps(meta)}
for cfg in meta["configs"]:
    setup, tb = make_config(cfg)
    r = orc.trace_mc_rays(setup, tb, meta["first_ray"], meta["n"], meta["seed"])
    for name in ("x", "y", "w", "code", "shell"):
        out[f"{cfg}_{name}"] = getattr(r, name)
np.savez_compressed(Path(__file__).resolve().parent / "oracle_rays_v1.npz", **out)
print("written")
