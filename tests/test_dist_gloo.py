"""world_size-2 gloo test of the multi-GPU plumbing (sharding + the one stats reduction)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from cl4wsis_b200 import dist as cdist
    r, _, w = cdist.init_from_env("gloo")
    lo, hi = cdist.shard_bounds(n_items, r, w)
    items = torch.arange(n_items, dtype=torch.float64)[lo:hi]
    cdist.barrier()
    out = cdist.reduce_stats(hi - lo, 1.0 + r, float(items.sum()), float((items * 2).sum()))
    q.put((r, lo, hi, out))
    dist.destroy_process_group()


def test_shard_bounds_partition():
    from cl4wsis_b200.dist import shard_bounds
    for n, w in [(10582, 8), (16, 2), (5, 8), (0, 4), (1323 * 8 - 2, 8)]:
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_stats_reduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n = _free_port(), 101
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 51, 51, 101)
    for _, _, _, out in res:
        assert out["images"] == n and out["elapsed_s"] == 2.0
        assert out["checksum_mask"] == sum(range(n)) and out["checksum_ids"] == 2 * sum(range(n))


def test_reduce_stats_single_process():
    from cl4wsis_b200.dist import reduce_stats
    out = reduce_stats(16, 0.5, 3.0, 4.0)
    assert out == {"images": 16.0, "elapsed_s": 0.5, "checksum_mask": 3.0, "checksum_ids": 4.0}
