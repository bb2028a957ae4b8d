"""N > 1 host logic on CPU: world_size-2 gloo run of the product's shard + merge code, with the oracle standing in for
the per-rank GPU trace. Philox keys on the global ray index, so two half runs merged must equal one full run."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_shard_partition_properties():
    from solaraxionraytracing_b200.multi_gpu import shard
    for n in (0, 1, 7, 1000, 10**11 + 3):
        for world in (1, 2, 3, 4, 8):
            parts = [shard(n, r, world, first_ray=5) for r in range(world)]
            assert parts[0][0] == 5
            assert sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1                     # contiguous, no overlap
            counts = [c for _, c in parts]
            assert max(counts) - min(counts) <= 1        # balanced
    with pytest.raises(ValueError):
        shard(10, 2, 2)


def test_counter_array_round_trip():
    from solaraxionraytracing_b200 import abi
    from solaraxionraytracing_b200.multi_gpu import N_COUNTER_INTS, N_COUNTER_WORDS, arrays_to_counters, counters_to_arrays
    c = abi.Counters()
    c.n_rays = 11; c.n_exit[0] = 5; c.n_exit[8] = 2; c.n_passed = 5; c.n_passed_till_window = 6; c.n_hit_nickel = 2
    c.n_interp_clamped = 1; c.sum_w = 1.5; c.sum_w2 = 2.5; c.sum_x = 3.5; c.sum_y = 4.5; c.sum_r = 5.5
    d = c.as_dict()
    ints, flts = counters_to_arrays([d])
    assert ints.shape == (1, N_COUNTER_INTS) and flts.shape == (1, N_COUNTER_WORDS - N_COUNTER_INTS)
    # same order as the C struct, so device-side all-reduces of the raw words mean the same thing
    raw = np.frombuffer(bytes(c), dtype=np.int64)
    assert np.array_equal(raw[:N_COUNTER_INTS], ints[0])
    assert np.array_equal(np.frombuffer(bytes(c), dtype=np.float64)[N_COUNTER_INTS:], flts[0])
    assert arrays_to_counters(ints, flts)[0] == d


def _worker(rank: int, world: int, port: int, n: int, out_path: str):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from helpers import make_config
    from oracle import oracle as orc
    from solaraxionraytracing_b200 import multi_gpu
    dist.init_process_group("gloo", rank=rank, world_size=world)
    setup, tb = make_config("cast_llnl")
    first, count = multi_gpu.shard(n, rank, world)
    img, img2, cnt = orc.trace_mc(setup, tb, first, count, 42)      # stand-in for the rank's GPU trace
    img, img2, cnt = multi_gpu.merge_host(img, img2, cnt)
    if rank == 0:
        np.savez(out_path, img=img, img2=img2, n_rays=cnt[0]["n_rays"], n_passed=cnt[0]["n_passed"],
                 sum_w=cnt[0]["sum_w"], nickel=cnt[0]["n_exit"]["nickel"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_equal_single_run(tmp_path, oracle):
    import torch.multiprocessing as mp
    from helpers import make_config
    n = 60_000
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    z = np.load(out)
    setup, tb = make_config("cast_llnl")
    img, img2, cnt = oracle.trace_mc(setup, tb, 0, n, 42)
    assert int(z["n_rays"]) == n == cnt[0]["n_rays"]
    assert int(z["n_passed"]) == cnt[0]["n_passed"]
    assert int(z["nickel"]) == cnt[0]["n_exit"]["nickel"]
    assert np.allclose(z["img"], img, rtol=1e-12, atol=0)
    assert np.allclose(z["img2"], img2, rtol=1e-12, atol=0)
    assert abs(float(z["sum_w"]) / cnt[0]["sum_w"] - 1) < 1e-12


class _FakeTracer:
    """Stands in for a RayTracer in performAngularScan: the flux of a scan point is a function of its angle and of
    the global ray range it was asked to trace, so a wrong shard offset or a point traced twice shows up in the sum."""

    def angular_scan(self, angles, n_rays, seed, want_images=False, first_ray=0):
        a = np.asarray(angles, dtype=np.float64)
        first = first_ray + n_rays * np.arange(a.size)
        return np.cos(a) * 1e-3 + first.astype(np.float64) * 1e-12 + seed * 1e-15, [{}] * a.size, None


def _scan_worker(rank: int, world: int, port: int, out_path: str):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from solaraxionraytracing_b200 import raytracer as rt
    dist.init_process_group("gloo", rank=rank, world_size=world)
    angles, rel, flux = rt.performAngularScan(None, 0.0, 0.3, 7, nRays=1000, seed=5, tracer=_FakeTracer(), rank=rank, world=world)
    if rank == 1:
        np.savez(out_path, angles=angles, rel=rel, flux=flux)
    dist.barrier()
    dist.destroy_process_group()


def test_angular_scan_sharded_by_scan_points(tmp_path):
    """performAngularScan over 2 ranks (gloo): contiguous blocks of scan points, scan point i traces the global rays
    [i n, (i+1) n) whichever rank owns it, fluxes summed over ranks — equal to the unsharded scan on every rank."""
    import torch.multiprocessing as mp
    from solaraxionraytracing_b200 import raytracer as rt
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "scan.npz")
    mp.spawn(_scan_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    angles, rel, flux = rt.performAngularScan(None, 0.0, 0.3, 7, nRays=1000, seed=5, tracer=_FakeTracer())
    assert np.array_equal(z["angles"], angles) and np.allclose(angles, np.linspace(0.0, 0.3, 7))
    assert np.allclose(z["flux"], flux, rtol=1e-15) and np.allclose(z["rel"], rel, rtol=1e-15)
    assert rel.max() == 1.0
