#!/usr/bin/env python
"""CPU-only: a ray of a tools/fuzz_setups.py setup under 60-digit arithmetic (tests/hp_trace.py: exit code and distance of the
nearest decision boundary in mm) next to the oracle's f64 outcome.  python tools/fuzz_ray_hp.py <seed> <turn 0/1> <setup index> <ray>..."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import fuzz_setups, hp_trace
from oracle import oracle as orc
rng = np.random.default_rng(int(sys.argv[1]))
if sys.argv[2] == "1": os.environ["FUZZ_TURN"] = "1"
idx = int(sys.argv[3])
for it in range(idx + 1):
    setup, tb, desc = fuzz_setups.random_setup(rng)
    seed = int(rng.integers(1, 2**62)); first = int(rng.choice([0, 17, 2**32 - 12345, 2**40 + 3]))
print("setup", idx, "; ".join(desc), "seed", seed, "first", first)
for ray in map(int, sys.argv[4:]):
    O, E, en = orc.sample_rays(setup, tb, ray, 1, seed)
    rec = orc.trace_presampled(setup, tb, O, E, en, optional=True)
    code, margin, amb = hp_trace.classify(setup, O[:, 0], E[:, 0])
    print("ray", ray, "oracle code", hex(int(rec.code[0])), "shell", int(rec.shell[0]), "alpha1/2", rec.alpha1[0], rec.alpha2[0],
          "| 60-digit code", code, "nearest boundary %.3e mm" % margin, "ambiguous" if amb else "")
