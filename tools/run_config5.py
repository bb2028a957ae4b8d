#!/usr/bin/env python
"""BASELINE config 5: BabyIAXO / InGridIAXO / vacuum / XMM (= config_default.toml as shipped), N rays sharded over the
GPUs of one box by contiguous global-ray-index ranges, one NCCL all-reduce of the detector image + counters.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/run_config5.py --rays 1e11
  python tools/run_config5.py --rays 1e9            # single GPU

Tables: Primakoff emission rates generated on the GPU from AGSS09 (1968 x 1500), Henke gold reflectivity on the
reference's 1000 x 1000 grid. Prints one JSON line (device time = max over ranks) and, with --check N, re-traces the
first N rays on rank 0 alone and verifies that the sharded integer counters of that prefix are identical.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=float, default=1e11)
    ap.add_argument("--chunk", type=float, default=4e9, help="rays per launch per GPU")
    ap.add_argument("--check", type=float, default=0, help="also verify shard-invariance on this many rays")
    ap.add_argument("--out", default="")
    ap.add_argument("--precision", type=int, default=2, help="0 exact, 1 fast, 2 f32")
    a = ap.parse_args()
    import torch
    from solaraxionraytracing_b200 import multi_gpu, output, raytracer as rt, tables
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    em = rt.calculateEmissionRates(processes=("primakoff",), device=local)
    rc, dc = rt.buildCdfs(em, local)
    tb = tables.TableSet(energies=em.energies, fluxRadiusCDF=rc, diffFluxCDFs=dc,
                         reflectivity=tables.gold_reflectivity_packaged(), **tables.detector_tables_packaged())
    setup = rt.newExperimentSetup("BabyIAXO", "InGridIAXO", "vacuum", "XMM", 0)
    fs = rt.FullRaytraceSetup(setup, tb)
    tr = rt.RayTracer(fs, local)
    tr.set_precision(a.precision)
    setup_s = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(tr.stream, device=local)
    views = multi_gpu.device_views(tr, local)

    def run(total: int):
        first, count = multi_gpu.shard(total, rank, world)
        with torch.cuda.stream(stream):
            tr.reset_image()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            done = 0
            while done < count:
                n = int(min(a.chunk, count - done))
                tr.trace_mc(n, 299792458, first_ray=first + done)
                done += n
            if dist is not None:
                multi_gpu.allreduce_device(tr, local, views=views)
            e1.record(stream)
            torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), tr.read_image()

    run(int(1e8))                                  # warm-up
    total = int(a.rays)
    ms, res = run(total)
    c = res.counters[0]
    ok = None
    if a.check:
        nchk = int(a.check)
        _, sharded = run(nchk)
        if rank == 0:
            tr.reset_image(); tr.trace_mc(nchk, 299792458); alone = tr.read_image()
            ok = (alone.counters[0]["n_exit"] == sharded.counters[0]["n_exit"]
                  and bool(np.allclose(alone.image, sharded.image, rtol=1e-9, atol=0)))
    if rank == 0:
        line = {"config": "5: BabyIAXO+InGridIAXO+vacuum+XMM (config_default.toml), Primakoff from AGSS09, Henke gold",
                "n_gpus": world, "precision": a.precision, "rays": total, "device_ms": ms, "rays_per_s": total / (ms * 1e-3), "setup_s": round(setup_s, 2),
                "n_rays": c["n_rays"], "n_passed": c["n_passed"], "passed_fraction": c["n_passed"] / max(1, c["n_rays"]),
                "sum_w": c["sum_w"], "rel_mc_error_total_flux": float(np.sqrt(c["sum_w2"]) / c["sum_w"]) if c["sum_w"] else None,
                "n_exit": c["n_exit"], "shard_invariant": ok}
        print(json.dumps(line), flush=True)
        if a.out:
            output.generateResultPlots(res, setup.detector.windowYear, a.out, suffix=f"_{world}gpu", echo=lambda s: None)
    tr.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
