#!/usr/bin/env python
"""Overlap-tile inference benchmark (BASELINE config 4): a synthetic 4096x4096 image is reflect-padded,
cut into 11 x 11 = 121 tiles of 572x572 (388x388 outputs), the tiles are dealt round-robin to the ranks
(no collective on the data path), each rank runs forward + centre-crop + argmax (tester.py:27-30) on the
B200 kernels, and the uint8 masks are gathered on rank 0 and stitched.  One JSON line on rank 0:
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
    from b200 import ops, tiling

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = Unet().to(dev)
    H = W = args.size
    g = torch.Generator().manual_seed(2)
    image = torch.rand(H, W, generator=g).pin_memory()
    tile_out, halo, ny, nx, _, _ = tiling.plan(H, W)
    n_tiles = ny * nx
    mine = list(range(rank, n_tiles, world))                      # round-robin shard, independent units

    def run(img_dev):
        tiles = tiling.cut_tiles(img_dev)[mine]
        out = torch.empty(len(mine), tile_out, tile_out, dtype=torch.uint8, device=dev)
        for b0 in range(0, len(mine), args.batch):
            t = tiles[b0:b0 + args.batch]
            am = out[b0:b0 + t.shape[0]]
            net.engine.forward(t, train=False, argmax_out=am, want_logits=False)
        return out

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    img_dev = image.to(dev)
    for _ in range(max(3, args.warmup)):
        run(img_dev)
    sync()
    ops.STATS.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = run(img_dev)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    launches = ops.STATS.launches

    # end to end: host image in -> stitched mask on host (rank 0).  One untimed pass first: the gather's NCCL
    # channels and the pinned result buffer are set up lazily.
    def e2e_once():
        out = run(image.to(dev, non_blocking=True))
        if world > 1:
            cap = -(-n_tiles // world)                                    # gather needs equal shapes: pad to the max share
            padded = torch.zeros(cap, tile_out, tile_out, dtype=torch.uint8, device=dev)
            padded[:out.shape[0]] = out
            parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
            dist.gather(padded, parts, dst=0)
            if rank == 0:
                full = torch.empty(n_tiles, tile_out, tile_out, dtype=torch.uint8, device=dev)
                for r in range(world):
                    cnt = len(range(r, n_tiles, world))
                    full[r::world] = parts[r][:cnt]
        else:
            full = out
        return tiling.stitch(full, H, W).cpu() if rank == 0 else None

    mask = e2e_once()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mask = e2e_once()
    sync()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if rank == 0:
        mpix = H * W / 1e6
        print(json.dumps({
            "metric": "inference Mpix/s (overlap-tile, original-image pixels)", "value": mpix * args.steps / (ms * 1e-3),
            "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_image": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{H}x{W} image -> {n_tiles} tiles of 572x572 (388x388 out), sharded round-robin over {world} GPU(s), "
                                   f"batches of {args.batch}, forward + crop + argmax, random-init weights", "baseline_config": 4},
            "tiles_per_s": n_tiles * args.steps / (ms * 1e-3),
            "fwd_tflops": 300.86e9 * n_tiles * args.steps / (ms * 1e-3) / 1e12,
            "e2e": {"value": mpix / e2e_s, "unit": "Mpix/s", "h2d_bytes_per_step": H * W * 4, "d2h_bytes_per_step": H * W},
            "gpu_launches": launches, "mask_foreground_frac": float(mask.float().mean()),
        }), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
