#!/usr/bin/env python
"""Overlap-tile inference benchmark (BASELINE config 4): a synthetic 4096x4096 image is reflect-padded,
cut into 11 x 11 = 121 tiles of 572x572 (388x388 outputs), every rank takes a contiguous range of tiles
(no collective on the data path), uploads only the image rows they read, runs forward + centre-crop + argmax
(tester.py:27-30) on the B200 kernels, and the uint8 masks are gathered on rank 0 and stitched (b200.tiling.TileSegmenter).  One JSON line on rank 0:
Mpix/s = original-image pixels / time.   python bench_infer.py [--gpus N] [--steps K] [--warmup W] [--batch 16]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))


def _claim_stdout():
    """stdout carries exactly ONE JSON line: keep the real fd for it and route everything else that
    writes to fd 1 (e.g. NCCL's "NCCL version ..." banner) to stderr."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


JSON_OUT = _claim_stdout()


def measure(net, dev, world, rank, steps, warmup, batch=16, size=4096, group=None):
    """BASELINE config 4 on the given module: returns the result dict on rank 0 (None elsewhere).  Device-timed value:
    the image band is resident, cut -> forward -> argmax per tile batch, max over ranks.  e2e: a stream of pinned host
    images through TileSegmenter.segment_many -- band upload, forward, mask gather + stitch on rank 0, mask download,
    pipelined over two slots -- images per second x original pixels."""
    import torch
    import torch.distributed as dist
    from b200 import ops, tiling
    H = W = size
    g = torch.Generator().manual_seed(2)
    image = torch.rand(H, W, generator=g).pin_memory()
    seg = tiling.TileSegmenter(net, H, W, 572, batch, rank, world, group=group, device=dev)
    n_tiles = seg.ny * seg.nx

    def sync():
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()

    band = seg.band[0]
    band.copy_(seg.band_of(image))
    W_ = max(3, warmup)
    for _ in range(W_):
        seg.segment_band(band)
    sync()
    ops.STATS.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        seg.segment_band(band)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        ms = t.item()
    launches = ops.STATS.launches

    # end to end.  One untimed pass first: the gather's NCCL channels are set up lazily.
    for _ in seg.segment_many([image] * 2):
        pass
    sync()
    ke = max(steps, 4)
    t0 = time.perf_counter()
    mask = None
    for _, m in seg.segment_many([image] * ke):
        mask = m
    sync()
    e2e_s = (time.perf_counter() - t0) / ke
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        e2e_s = t.item()
    if rank != 0:
        return None
    mpix = H * W / 1e6
    return {
        "metric": "inference Mpix/s (overlap-tile, original-image pixels)", "value": mpix * steps / (ms * 1e-3),
        "unit": "Mpix/s", "n_gpus": world, "steps": steps, "warmup": W_, "ms_per_image": ms / steps,
        "higher_is_better": True, "scaling": "strong", "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{H}x{W} image -> {n_tiles} tiles of 572x572 (388x388 out), contiguous tile ranges over {world} "
                               f"GPU(s) ({seg.cap} tiles on the fullest rank), batches of {batch}, forward + crop + argmax, "
                               f"random-init weights", "baseline_config": 4},
        "tiles_per_s": n_tiles * steps / (ms * 1e-3),
        "fwd_tflops": 300.86e9 * n_tiles * steps / (ms * 1e-3) / 1e12,
        "e2e": {"value": mpix / e2e_s, "unit": "Mpix/s", "images": ke,
                "h2d_bytes_per_step": seg.rows * W * 4, "d2h_bytes_per_step": H * W,
                "what": "pinned host image in (each rank uploads the band of rows its tiles read), stitched uint8 mask on "
                        "the host of rank 0 out; upload / forward / gather + stitch / download pipelined over two slots"},
        "gpu_launches": launches, "mask_foreground_frac": float(mask.float().mean()),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=4096)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from network import Unet

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = Unet().to(dev)
    out = measure(net, dev, world, rank, args.steps, args.warmup, args.batch, args.size)
    if rank == 0:
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
