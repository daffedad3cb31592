"""BASELINE config 5: one synthetic SBD-sized pass (10 582 images, C21, 512x512) of PAMR + centre NMS + grouping,
images sharded contiguously over the ranks (cl4wsis_b200.dist.shard_bounds), batches of 16, no data-path
collective; one all_reduce of (images, max elapsed, checksums) at the end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sbd_pass.py

Inputs: each rank generates one batch (seed 1234 + rank, bench.synth_inputs) and rolls it along the batch
dimension from batch to batch; the tail batch is run full-size and only its real images are counted."""
import json
import os
import sys

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the JSON line
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cl4wsis_b200 as cl4  # noqa: E402
from cl4wsis_b200 import dist as cdist  # noqa: E402

N_IMAGES = int(os.environ.get("SBD_IMAGES", "10582"))


def main():
    rank, local, world = cdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = bench.WORKLOADS["voc_b16_c21_512"]
    B = cfg["B"]
    lo, hi = cdist.shard_bounds(N_IMAGES, rank, world)
    n_mine = hi - lo
    n_batches = (n_mine + B - 1) // B
    img, mask, heat, off = (t.to(dev) for t in bench.synth_inputs(cfg, 1234 + rank))
    step = cl4.PseudoLabelStep(B, cfg["C"], cfg["H"], cfg["W"], num_iter=cfg["T"], dilations=cfg["dil"],
                               threshold=cfg["thr"], nms_kernel=cfg["nms"], max_centers=256, device=dev)
    for _ in range(3):
        step.run(img, mask, heat, off)
    torch.cuda.synchronize()
    cdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ck_mask = torch.zeros((), dtype=torch.float64, device=dev)
    ck_ids = torch.zeros((), dtype=torch.float64, device=dev)
    e0.record()
    for k in range(n_batches):
        s = k % B
        xi, mi, hi_, oi = (torch.roll(t, s, 0) for t in (img, mask, heat, off)) if s else (img, mask, heat, off)
        refined, ids, counts, _ = step.run(xi, mi, hi_, oi)
        real = min(B, n_mine - k * B)
        ck_mask += refined[:real].sum(dtype=torch.float64)
        ck_ids += ids[:real].sum(dtype=torch.float64)
    e1.record()
    torch.cuda.synchronize()
    stats = cdist.reduce_stats(n_mine, e0.elapsed_time(e1) / 1e3, float(ck_mask), float(ck_ids), device=dev)
    cdist.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps({"workload": "sbd_pass_10582_c21_512", "n_gpus": world, "images": stats["images"],
                          "pass_seconds": stats["elapsed_s"], "images_per_s": stats["images"] / stats["elapsed_s"],
                          "checksum_mask": stats["checksum_mask"], "checksum_ids": stats["checksum_ids"],
                          "note": "timed region includes the batch rolls and the checksum reductions"}), flush=True)


if __name__ == "__main__":
    main()
