"""Sharding of one Monte Carlo run over the GPUs of a box: one process per GPU, torch.distributed for the plumbing.

Rays are independent and ray i depends only on (seed, i) (Philox counter = global ray index), so the run is cut into
contiguous index ranges, one per rank, with no data-path exchange; the only collective is the final sum of the
detector image, the sum-of-squares image and the counters (512 KiB + 512 KiB + 224 B per axion mass) — one
all-reduce over NVLink (NCCL) on the device buffers libsart exposes. The reference's only parallel construct is the
Weave parallelFor over rays inside one process (src/raytracer.nim:2234-2244).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


def shard(n_rays: int, rank: int, world: int, first_ray: int = 0) -> tuple[int, int]:
    """Contiguous, balanced partition of [first_ray, first_ray + n_rays): returns (first, count) of `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_rays, world)
    count = base + (1 if rank < rem else 0)
    first = first_ray + rank * base + min(rank, rem)
    return first, count


N_COUNTER_WORDS = C.sizeof(abi.Counters) // 8
N_COUNTER_INTS = abi.Counters.sum_w.offset // 8     # the leading u64 fields; the rest are f64 sums


def counters_to_arrays(counters: list[dict]) -> tuple[np.ndarray, np.ndarray]:
    """Counter dicts (one per axion mass) -> (int64 [M, n_int], float64 [M, n_float]) in struct order."""
    ints, flts = [], []
    for c in counters:
        ex = [c["n_exit"].get(name, 0) for name in abi.EXIT_NAMES] + [0] * (16 - abi.N_EXIT_CODES)
        ints.append([c["n_rays"], *ex, c["n_passed"], c["n_passed_till_window"], c["n_hit_nickel"],
                     c["n_interp_clamped"], c.get("n_retraced", 0), c.get("n_unresolved", 0)])
        flts.append([c["sum_w"], c["sum_w2"], c["sum_x"], c["sum_y"], c["sum_r"]])
    return np.asarray(ints, dtype=np.int64), np.asarray(flts, dtype=np.float64)


def arrays_to_counters(ints: np.ndarray, flts: np.ndarray) -> list[dict]:
    out = []
    for i, f in zip(ints, flts):
        d = {"n_rays": int(i[0]), "n_exit": {name: int(i[1 + k]) for k, name in enumerate(abi.EXIT_NAMES)},
             "n_passed": int(i[17]), "n_passed_till_window": int(i[18]), "n_hit_nickel": int(i[19]),
             "n_interp_clamped": int(i[20]), "n_retraced": int(i[21]), "n_unresolved": int(i[22]), "sum_w": float(f[0]), "sum_w2": float(f[1]), "sum_x": float(f[2]),
             "sum_y": float(f[3]), "sum_r": float(f[4])}
        out.append(d)
    return out


def merge_host(image: np.ndarray, image_w2: np.ndarray, counters: list[dict], group=None):
    """All-reduce(sum) of host-side results over `group` (any torch.distributed backend; gloo on CPU).
    Returns (image, image_w2, counters) holding the whole run on every rank."""
    import torch
    import torch.distributed as dist
    ints, flts = counters_to_arrays(counters)
    t_img, t_img2 = torch.from_numpy(np.ascontiguousarray(image)), torch.from_numpy(np.ascontiguousarray(image_w2))
    t_i, t_f = torch.from_numpy(ints), torch.from_numpy(flts)
    for t in (t_img, t_img2, t_i, t_f):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t_img.numpy(), t_img2.numpy(), arrays_to_counters(t_i.numpy(), t_f.numpy())


def comm_init(tracer, rank: int, world: int, device: int, group=None):
    """Gives `tracer` its NCCL communicator (sart_comm_init_rank): rank 0 draws the 128-byte NCCL id and broadcasts it
    over the already initialised torch.distributed group — the only thing torch ships; the collective itself is
    sart_allreduce behind the C-ABI."""
    import torch
    import torch.distributed as dist
    uid = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(tracer.comm_unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        uid = uid.to(f"cuda:{device}")
    dist.broadcast(uid, src=0, group=group)
    tracer.comm_init_rank(world, rank, bytes(uid.cpu().numpy().tobytes()))


class _DevArray:
    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_views(tracer, device: int):
    """Zero-copy torch views of a RayTracer's device-resident image, w^2 image and counters (ints, floats)."""
    import torch
    img, img2, n = tracer.image_dev()
    m = tracer.n_masses
    dev = f"cuda:{device}"
    cnt = tracer.counters_dev()
    views = [torch.as_tensor(_DevArray(img, n, "<f8"), device=dev), torch.as_tensor(_DevArray(img2, n, "<f8"), device=dev)]
    if m == 1:
        views.append(torch.as_tensor(_DevArray(cnt, N_COUNTER_INTS, "<i8"), device=dev))
        views.append(torch.as_tensor(_DevArray(cnt + 8 * N_COUNTER_INTS, N_COUNTER_WORDS - N_COUNTER_INTS, "<f8"), device=dev))
    else:
        # interleaved int/float words per mass: reduce each mass record's two halves separately
        for k in range(m):
            base = cnt + k * 8 * N_COUNTER_WORDS
            views.append(torch.as_tensor(_DevArray(base, N_COUNTER_INTS, "<i8"), device=dev))
            views.append(torch.as_tensor(_DevArray(base + 8 * N_COUNTER_INTS, N_COUNTER_WORDS - N_COUNTER_INTS, "<f8"), device=dev))
    return views


def allreduce_device(tracer, device: int, group=None, views=None):
    """One NCCL all-reduce(sum) per buffer on the tracer's stream (call inside `torch.cuda.stream(ext_stream)`)."""
    import torch.distributed as dist
    for t in (views or device_views(tracer, device)):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def calculateFluxFractionsSharded(raytraceSetup, nRays: int, seed: int = 299792458, device: int | None = None,
                                  group=None):
    """calculateFluxFractions (rt:2755-2776) over all ranks of an initialised torch.distributed NCCL group: every
    rank traces its shard on its GPU, the images are summed with one all-reduce, every rank returns the full result."""
    import torch
    import torch.distributed as dist
    from . import raytracer as rt
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = torch.cuda.current_device() if device is None else device
    first, count = shard(nRays, rank, world)
    with rt.RayTracer(raytraceSetup, device) as tr:
        for mode in (2, 1):      # the fastest pipeline this setup supports; precision 0 otherwise
            try:
                tr.set_precision(mode)
                break
            except rt.SartError:
                pass
        comm_init(tr, rank, world, device, group)
        tr.reset_image()
        tr.trace_mc(count, seed, first_ray=first)
        tr.allreduce()
        return tr.read_merged()
