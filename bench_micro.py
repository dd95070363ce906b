#!/usr/bin/env python
"""Micro-benchmarks of the HBM-bound kernels of the path against the measured HBM roofline
(BASELINE config 5 = weighted loss fwd/bwd at batch 64, 388x388, 2 classes; plus pool, skip write,
head, conv11c, SGD).  One JSON line per kernel:  python bench_micro.py [--reps 20]

achieved = ALGORITHMIC bytes / CUDA-event time; every working set is > the 126 MB L2
except where noted, and the buffers are rotated so that no launch re-reads what the previous one left
in L2."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    from b200 import ops

    dev = torch.device("cuda:0")
    peak = 6538.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)

    def run(name, nbytes, fns, note=""):
        for f in fns:
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.reps):
            fns[i % len(fns)]()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / args.reps * 1e3
        gbs = nbytes / us / 1e3
        print(json.dumps({"kernel": name, "us": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 1),
                          "achieved_GBs": round(gbs, 1), "peak_GBs": peak, "frac": round(gbs / peak, 3), "note": note}), flush=True)

    torch.manual_seed(0)
    # ---- BASELINE config 5: loss microbench, batch 64, 388x388, fp32 logits, uint8 labels, fp32 weight map
    B, L = 64, 388
    npx = B * L * L
    zs = [torch.randn(B, 2, L, L, device=dev) * 10 for _ in range(3)]
    lab = (torch.rand(B, L, L, device=dev) > 0.6).to(torch.uint8)
    wm = 1 + 10 * torch.rand(B, L, L, device=dev)
    dz = [torch.empty_like(zs[0]) for _ in range(3)]
    out = torch.zeros(1, device=dev)
    ws = ops.loss_workspace(dev)
    for kind, nm in ((ops.LOSS_BCE, "bce_with_logits"), (ops.LOSS_SOFTMAX_CE, "softmax_ce")):
        run(f"loss_{nm}_fwd", npx * (8 + 1 + 4), [lambda i=i: ops.loss_fwd_bwd(kind, zs[i], lab, wm, None, out, ws, 1.0) for i in range(3)],
            "13 B/pixel: fp32 logits x2 + u8 label + fp32 weight")
        run(f"loss_{nm}_fwd_bwd", npx * (8 + 1 + 4 + 8), [lambda i=i: ops.loss_fwd_bwd(kind, zs[i], lab, wm, dz[i], out, ws, 1.0) for i in range(3)],
            "21 B/pixel: + fp32 dlogits x2")
    del zs, dz

    # ---- pools of the 572 network at 16 tiles (fwd: 8 B in + 2 B out + 1 B idx (+2 B skip write) per output element)
    for (h, c, hu) in ((568, 64, 392), (280, 128, 200), (136, 256, 104), (64, 512, 56)):
        n = 16
        hp = h // 2
        coff = int((hp - hu) * 0.5)
        xs = [torch.randn(n, h, h, c, device=dev).clamp_min(0).to(torch.bfloat16) for _ in range(2)]
        y = torch.empty(n, hp, hp, c, dtype=torch.bfloat16, device=dev)
        idx = torch.empty(n, hp, hp, c, dtype=torch.uint8, device=dev)
        cat = torch.zeros(n, hu, hu, 2 * c, dtype=torch.bfloat16, device=dev)
        nout = n * hp * hp * c
        run(f"pool_fwd+skip_write {h}x{h}x{c}", nout * (8 + 2 + 1 + 2),
            [lambda i=i: ops.pool2x2_fwd(xs[i], y, idx, cat[..., :c], coff) for i in range(2)],
            "crop_and_concat placement fused: 13 B per pooled element")
        dy = torch.randn(n, hp, hp, c, device=dev).to(torch.bfloat16)
        dcat = torch.randn(n, hu, hu, 2 * c, device=dev).to(torch.bfloat16)
        dxs = [torch.empty(n, h, h, c, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        run(f"pool_bwd+skip_grad_sum {h}x{h}x{c}", nout * (2 + 2 + 1 + 2 + 8),
            [lambda i=i: ops.pool2x2_bwd(dy, dcat[..., :c], coff, idx, y, dxs[i]) for i in range(2)],
            "dy + dcat window + idx + pooled read, 4 positions written: 15 B per pooled element")
        dbias = torch.empty(ops.partial_rows(ops.PROD_POOL, n, h, h, c), c, device=dev)
        run(f"pool_bwd+skip_grad_sum+bias_grad {h}x{h}x{c}", nout * (2 + 2 + 1 + 2 + 8),
            [lambda i=i: ops.pool2x2_bwd(dy, dcat[..., :c], coff, idx, y, dxs[i], dbias_part=dbias) for i in range(2)],
            "the variant the training step runs: also leaves the partial rows of the bias gradient of the conv that fed the pool")
        del xs, dxs, cat, dcat

    # ---- conv11c, head, SGD
    n, s = 16, 572
    img = torch.rand(n, 1, s, s, device=dev)
    w1, b1 = torch.randn(64, 1, 3, 3, device=dev), torch.randn(64, device=dev)
    ys = [torch.empty(n, s - 2, s - 2, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    run("conv11c_fwd (1->64, bias+ReLU)", n * (s * s * 4 + (s - 2) ** 2 * 128), [lambda i=i: ops.conv1_fwd(img, w1, b1, ys[i]) for i in range(2)])
    c1part = torch.empty(ops.partial_rows(ops.PROD_CONV1, n, s, s), ops.CONV1_STRIDE, device=dev)
    run("conv11c_wgrad", n * (s * s * 4 + (s - 2) ** 2 * 128), [lambda i=i: ops.conv1_wgrad(img, ys[i], part=c1part) for i in range(2)])
    x = torch.randn(n, 388, 388, 64, device=dev).clamp_min(0).to(torch.bfloat16)
    x2 = x.clone()
    wf, bf_ = torch.randn(2, 64, 1, 1, device=dev), torch.randn(2, device=dev)
    lg = torch.empty(n, 2, 388, 388, device=dev)
    run("head_fwd (64->2)", n * 388 * 388 * (128 + 8), [lambda: ops.head_fwd(x, wf, bf_, lg, None), lambda: ops.head_fwd(x2, wf, bf_, lg, None)])
    dxh = torch.empty_like(x)
    hpart = torch.empty(ops.partial_rows(ops.PROD_HEAD, n, 388, 388), ops.HEAD_STRIDE, device=dev)
    run("head_bwd", n * 388 * 388 * (128 + 8 + 128), [lambda: ops.head_bwd(x, wf, lg, dxh, part=hpart), lambda: ops.head_bwd(x2, wf, lg, dxh, part=hpart)])
    g2 = torch.randn(n, 388, 388, 64, device=dev).to(torch.bfloat16)
    cs = torch.empty(ops.partial_rows(ops.PROD_COLSUM, n, 388, 388, 64), 64, device=dev)
    run("colsum (bias grad)", n * 388 * 388 * 128, [lambda: ops.colsum(x, part=cs), lambda: ops.colsum(g2, part=cs)])
    nel = 31_030_658
    pa, ga, ma = (torch.randn(nel + 6, device=dev) for _ in range(3))
    run("sgd_momentum (31.03 M params)", nel * 20, [lambda: ops.sgd_momentum(pa, ga, ma, 1e-4, 0.99, False)],
        "working set 372 MB > L2")

    # ---- loss weight maps (functions.py:7-78) for 16 label tiles of 388x388, on the device vs the reference's
    # CPU route (OpenCV per sample, restated in dl-unet_b200/functions.py) on the host cores of this box
    import time
    from b200 import weightmap
    from b200.synth import synthetic_batch
    import functions as Fn
    _, labels, _ = synthetic_batch(16, 572, seed=5)
    lab_cpu = labels[:, 0].contiguous()
    lab = lab_cpu.to(dev)
    wout = torch.empty(16, 388, 388, device=dev)
    px = 16 * 388 * 388
    for check in (False,):
        for f in (lambda: weightmap.weighted_map(lab, check=check, out=wout),):
            f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                f()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / args.reps * 1e3
    t0 = time.perf_counter()
    wref = Fn.weighted_map(lab_cpu[:4])
    cpu_us = (time.perf_counter() - t0) / 4 * 16 * 1e6
    same = float((wout[:4].cpu() - wref).abs().max())
    # algorithmic bytes: label 8 B in + fp32 map out per pixel (the candidate / component planes are internal)
    print(json.dumps({"kernel": "weighted_map 16 x 388x388 (6 launches)", "us": round(us, 2), "algorithmic_MB": round(px * 12 / 1e6, 1),
                      "achieved_GBs": round(px * 12 / us / 1e3, 1), "peak_GBs": peak, "frac": round(px * 12 / us / 1e3 / peak, 4),
                      "tiles_per_s": round(16 / (us * 1e-6), 1), "cpu_reference_us": round(cpu_us, 1),
                      "cpu_tiles_per_s": round(16 / (cpu_us * 1e-6), 2), "speedup_vs_cpu": round(cpu_us / us, 1),
                      "max_abs_diff_vs_cpu": same,
                      "note": "latency-bound integer kernels (union-find, column scan, pruned row scan); CPU = OpenCV route on "
                              f"{os.cpu_count()} host cores (single-threaded per sample, as the reference)"}), flush=True)


if __name__ == "__main__":
    main()
