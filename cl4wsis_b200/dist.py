"""Multi-GPU plumbing: images shard across ranks with no data-path collective;
the only collective is a tiny reduction of run statistics, mirroring the reference's
``distributed.reduce`` of scalars (train.py:575-576, modules/utils.py:40-41)."""
import os

import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world_size):
    """Contiguous block partition: the first ``n_items % world_size`` ranks get one extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend=None):
    """Join the process group described by RANK/WORLD_SIZE/MASTER_* (torchrun); no-op at world size 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world)
    return rank, local, world


def reduce_stats(n_images, elapsed_s, checksum_mask, checksum_ids, device="cpu"):
    """-> dict with the job-wide totals: images SUM, elapsed MAX over ranks, checksums SUM.
    A single all_reduce(SUM) of a 4-vector plus one all_reduce(MAX); 40 bytes on the wire."""
    sums = torch.tensor([float(n_images), float(checksum_mask), float(checksum_ids)], dtype=torch.float64, device=device)
    tmax = torch.tensor([float(elapsed_s)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return {"images": sums[0].item(), "elapsed_s": tmax[0].item(), "checksum_mask": sums[1].item(),
            "checksum_ids": sums[2].item()}


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


_AFFINITY_NOTE = "unchanged"


def pin_to_local_cpus(local_rank):
    """Bind this process to the CPU cores NVML reports as local to its GPU (the cores of the GPU's NUMA node), so that
    the pinned host buffers of the end-to-end path are allocated from, and copied out of, node-local memory.  When
    several ranks share the same core set (one NUMA node for all GPUs), each rank takes its own contiguous slice, which
    keeps the ranks' submit threads from migrating over one another.  Best effort: returns a description, never raises."""
    global _AFFINITY_NOTE
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            _AFFINITY_NOTE = "NVML affinity empty; unchanged"
            return _AFFINITY_NOTE
        world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
        if world > 1 and len(allowed) >= 2 * world:
            per = len(allowed) // world
            mine = allowed[local_rank * per:(local_rank + 1) * per]
        else:
            mine = allowed
        os.sched_setaffinity(0, mine)
        _AFFINITY_NOTE = f"cores {mine[0]}-{mine[-1]} ({len(mine)} of the {len(allowed)} NVML-local cores of GPU {local_rank})"
    except Exception as e:  # noqa: BLE001
        _AFFINITY_NOTE = f"unchanged ({type(e).__name__})"
    return _AFFINITY_NOTE


def affinity_note():
    return _AFFINITY_NOTE
