#!/usr/bin/env python
"""Benchmark of the U-Net training hot path (BASELINE.json metric: 572x572 train tiles/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = forward + weighted BCE loss + backward (+ gradient all-reduce at N > 1) + SGD-momentum
update over one batch of synthetic 572x572 tiles:
  * N = 1: BASELINE config 2 -- batch 16, bf16 operands / fp32 accumulate, on 1 x B200;
  * N > 1: BASELINE config 3 -- global batch 128 split evenly (64 / 32 / 16 tiles per GPU at 2/4/8).
The step is b200.train.TrainStep, the call dl-unet_b200/trainer.py makes per batch (CUDA-graph replay at N = 1).
`value` is device-resident throughput (inputs already in HBM), `e2e` runs the same call with pinned HOST buffers
uploaded every step (copy stream, double-buffered) and the loss read back every step.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
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

METRIC = "572x572 train tiles/sec (fwd+bwd)"
UNIT = "tiles/s"
TILE = 572
FLOPS_PER_TRAIN_TILE = 902.2e9


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1399.0), d.get("hbm_gbs", 6538.6), "measured"
    return 1400.0, 6650.0, "fallback"  # B200_PROFILING.md fallback (sustained bf16, HBM)


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_steps(steps, warmup, threads=None):
    """The reference's CPU implementation of the path (trainer.py:54-78 restated in oracle/): fp32, batch 1,
    one 572x572 tile per step (BASELINE config 1), all host threads.  Returns (tiles/s, cores, sample)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import unet_oracle as O
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.init_state_dict(0)
    images, labels, weight = O.synthetic_batch(1, TILE, seed=1)
    mom = {}
    for _ in range(warmup):
        O.train_step(sd, mom, images, labels, weight)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(sd, mom, images, labels, weight)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return 1.0 / dt, cores, dt


def cudnn_steps(per_gpu, steps, warmup, device, mode):
    """The "existing Blackwell kernel" bar: the reference's own operator sequence
    (network.py:131-190 + trainer.py:54-78, restated in oracle/unet_oracle.py) dispatched by stock PyTorch to
    cuDNN / ATen CUDA kernels on the same B200, same batch and tile size.  mode "fp32": the reference as written
    (NCHW fp32, TF32 off as torch defaults); "tf32": same with torch's TF32 switches on; "bf16": autocast(bf16) +
    channels_last, the usual mixed-precision recipe.  Returns (tiles/s, ms per step)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import unet_oracle as O
    torch.backends.cudnn.benchmark = True
    tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    sd = {k: v.to(device) for k, v in O.init_state_dict(0).items()}
    images, labels, weight = (t.to(device) for t in O.synthetic_batch(per_gpu, TILE, seed=1000))
    if mode == "bf16":
        images = images.contiguous(memory_format=torch.channels_last)
        sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
    mom = {}

    def one():
        if mode == "bf16":
            params = {k: v.detach().requires_grad_(True) for k, v in sd.items()}
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = O.forward(params, images)
            preds = O.center_crop_preds(logits.float(), labels.shape[-1])
            loss = O.loss_bce(preds, labels, weight)
            grads = torch.autograd.grad(loss, list(params.values()))
            with torch.no_grad():
                ks = list(sd.keys())
                if not mom:
                    for k, g in zip(ks, grads):
                        mom[k] = g.clone()
                else:
                    torch._foreach_mul_([mom[k] for k in ks], 0.99)
                    torch._foreach_add_([mom[k] for k in ks], list(grads))
                torch._foreach_add_([sd[k] for k in ks], [mom[k] for k in ks], alpha=-1e-4)
            return loss
        return O.train_step(sd, mom, images, labels, weight)[0]

    for _ in range(max(3, warmup)):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    torch.backends.cudnn.allow_tf32 = True     # torch defaults
    torch.backends.cuda.matmul.allow_tf32 = False
    return per_gpu / (ms * 1e-3), ms


def run_cudnn(args):
    """--impl cudnn: stock PyTorch + cuDNN on one B200 (rank 0 only), same metric / config as our arm."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    per_gpu = args.batch_per_gpu or 16
    res = {}
    for mode in ("bf16", "tf32", "fp32"):
        v, ms = cudnn_steps(per_gpu, max(1, min(args.steps, 10)), args.warmup, dev, mode)
        res[mode] = {"value": v, "ms_per_step": ms}
        torch.cuda.empty_cache()
    best = max(res, key=lambda k: res[k]["value"])
    print(json.dumps({
        "impl": "cudnn", "metric": METRIC, "value": res[best]["value"], "unit": UNIT, "n_gpus": 1,
        "steps": max(1, min(args.steps, 10)), "warmup": max(3, args.warmup), "ms_per_step": res[best]["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": best, "data": "synthetic",
        "config": dict(workload_config(per_gpu, per_gpu, 1),
                       what="the reference's operator sequence (oracle/unet_oracle.py = network.py:131-190 + "
                            "trainer.py:54-78) on stock torch %s + cuDNN %s, device-resident inputs" %
                            (torch.__version__, torch.backends.cudnn.version())),
        "modes": res, "gpu_launches": 0,
    }), file=JSON_OUT, flush=True)


def workload_config(per_gpu, global_batch, world):
    """The `config` object both arms report (BASELINE config 2 at N = 1, config 3 at N > 1)."""
    return {"workload": f"original U-Net (64-1024 ch, valid convs) 572x572->388x388 training step "
                        f"(fwd + weighted BCE + bwd + SGD-momentum), {per_gpu} tiles per GPU, global batch "
                        f"{global_batch}, random-init weights (reference init, seed 0), weight map w0=10 sigma=5",
            "baseline_config": 2 if world == 1 else 3, "tiles_per_gpu": per_gpu, "global_batch": global_batch,
            "parallelism": f"dp{world}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    v, cores, dt = cpu_reference_steps(steps, warm)
    sample = f"{steps} full train steps (fwd+weighted BCE+bwd+SGD) on one 572x572 tile, batch 1, fp32, after {warm} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(16 if args.gpus == 1 else 128 // args.gpus, 16 if args.gpus == 1 else 128, args.gpus),
                       sample="each step = ONE tile of that workload (batch 1, fp32: BASELINE config 1) on the host cores; the "
                              "reference is pure PyTorch (no build, cannot travel): its path is restated in "
                              "oracle/unet_oracle.py and pinned bit-exactly against /root/reference by "
                              "oracle/validate_against_reference.py"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=JSON_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cudnn"])
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="override the BASELINE batch (diagnostics only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock torch + cuDNN arm (N = 1)")
    ap.add_argument("--no-inference", action="store_true", help="skip the overlap-tile inference measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "cudnn":
        return run_cudnn(args)

    import torch
    import torch.distributed as dist
    from network import Unet
    from b200 import ops, synth
    from b200 import train as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    per_gpu = args.batch_per_gpu or (16 if world == 1 else 128 // world)
    global_batch = per_gpu * world

    torch.manual_seed(0)
    net = Unet().to(dev)
    opt = T.FusedSGD(net, lr=1e-4, momentum=0.99)
    reducer = None
    images, labels, weight = synth.synthetic_batch(per_gpu, TILE, seed=1000 + rank)
    labels_u8 = labels[:, 0].to(torch.uint8)
    h_img, h_lab, h_w = images.pin_memory(), labels_u8.pin_memory(), weight.pin_memory()
    d_img, d_lab, d_w = h_img.to(dev), h_lab.to(dev), h_w.to(dev)
    gb = global_batch if world > 1 else None

    # The call a user makes: b200.train.TrainStep, the hot-loop body of trainer.py:54-78 (zero_grad, forward, crop +
    # one-hot + weighted BCE, backward, SGD-momentum) as one call -- what dl-unet_b200/trainer.py::training runs per
    # batch.  On one GPU it replays a CUDA graph of the step; with the all-reducer attached (N > 1) it issues the
    # same kernel schedule eagerly so that the gradient buckets can be reduced while backward runs.
    # N > 1: the step incl. its bucketed all-reduces is captured into one CUDA graph (UB_DP_GRAPH=0: eager schedule)
    dp_graph = world > 1 and os.environ.get("UB_DP_GRAPH", "1") == "1"
    tstep = T.TrainStep(net, opt, kind="bce", graph=(world == 1 or dp_graph), global_batch=gb, graph_with_allreduce=dp_graph)
    tstep_eager = T.TrainStep(net, opt, kind="bce", graph=False, global_batch=gb)

    def step(img, lab, wmap):
        return tstep(img, lab, wmap)

    step(d_img, d_lab, d_w)  # builds the arenas
    if world > 1:
        # gradient exchange: our own NVLS / peer-memory all-reduce on the compute stream (default), or the bucketed NCCL
        # all-reduce on a side stream (UB_DP_ALLREDUCE=nccl, also the fallback where symmetric memory is unavailable)
        from b200.dist import make_reducer
        reducer = make_reducer(net)
        reducer.broadcast_parameters(0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident timed region
    for _ in range(W):
        step(d_img, d_lab, d_w)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step(d_img, d_lab, d_w)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = global_batch * K / (ms * 1e-3)

    # ---- the same K steps once more, issued kernel by kernel with CUDA events around every tensor-pipe launch:
    # per-family kernel times for the roofline object and the launch count of a step (a graph replay launches the
    # same kernels, but hides them from host-side accounting)
    tstep_eager(d_img, d_lab, d_w)
    barrier()
    ops.STATS.reset()
    ops.STATS.timing = True
    for _ in range(K):
        tstep_eager(d_img, d_lab, d_w)
    barrier()
    ops.STATS.timing = False
    launches = ops.STATS.launches
    fam_all = ops.STATS.summarize()
    fam_ig = ops.STATS.summarize("igemm").get("igemm", (0.0, 0.0, 1e-9, 0))
    fam_wg = ops.STATS.summarize("wgrad").get("wgrad", (0.0, 0.0, 1e-9, 0))

    # ---- where the time goes at N > 1: the same per-GPU batch without the exchange step (reducer detached: eager
    # schedule, and as the CUDA-graph replay a single GPU would run), against the timed region above
    breakdown = None
    if world > 1:
        def timed(fn, n):
            for _ in range(2):
                fn(d_img, d_lab, d_w)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn(d_img, d_lab, d_w)
            b.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / n

        reducer.detach()
        kb = max(3, min(K, 10))
        t_eager = timed(tstep_eager, kb)
        tgraph = T.TrainStep(net, opt, kind="bce", graph=True, global_batch=gb)
        t_graph = timed(tgraph, kb)
        tgraph.graphs.clear()
        reducer.attach()
        reducer.broadcast_parameters(0)
        t_total = ms / K
        breakdown = {"ms_per_step": t_total, "compute_only_eager_ms": t_eager, "compute_only_graph_ms": t_graph,
                     "exposed_allreduce_ms": t_total - t_eager, "exposed_allreduce_frac": (t_total - t_eager) / t_total,
                     "eager_vs_graph_ms": t_eager - t_graph,
                     "n1_equivalent_tiles_per_s": per_gpu / (t_graph * 1e-3),
                     "allreduce_bytes_per_step": net.engine.store.total * 4,
                     "allreduce": ("own kernels over symmetric memory, " + ("NVLS multimem.ld_reduce / multimem.st" if reducer.multicast else "peer loads / stores")
                                   + ", on the compute stream after backward") if hasattr(reducer, "multicast") else
                                  f"NCCL, buckets of {reducer.bucket_elems * 4 / (1 << 20):.0f} MB on a high-priority side stream",
                     "schedule": "one CUDA graph incl. the exchange" if dp_graph else "eager kernel schedule",
                     "what": "compute_only_*: the same per-GPU batch with the all-reducer detached (max over ranks): eager = "
                             "the schedule the timed region uses minus the all-reduces, graph = what one GPU alone runs; "
                             "exposed_allreduce = timed step - compute_only_eager"}

    # ---- end-to-end: pinned host inputs in, loss out, every step, through the public API.  Every step's inputs cross
    # PCIe inside the timed region; the copy of step i+1 is issued on a copy stream while step i computes (what an
    # input pipeline does), and the loss of every step is read back to the host.
    ke = max(1, K)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [tuple(torch.empty_like(t) for t in (d_img, d_lab, d_w)) for _ in range(2)]   # double-buffered staging
    consumed = [None, None]

    def upload(k):
        dst = slots[k % 2]
        with torch.cuda.stream(copy_stream):
            if consumed[k % 2] is not None:
                copy_stream.wait_event(consumed[k % 2])            # the step that read this slot has finished
            for d, h in zip(dst, (h_img, h_lab, h_w)):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dst, ev

    def e2e_pass(n):
        marks = [time.perf_counter()]
        nxt = upload(0)
        for i in range(n):
            (b_img, b_lab, b_w), ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < n:
                nxt = upload(i + 1)
            l = step(b_img, b_lab, b_w)
            consumed[i % 2] = torch.cuda.Event()
            consumed[i % 2].record(torch.cuda.current_stream())
            l.item()
            marks.append(time.perf_counter())
        return marks

    e2e_pass(2)                          # untimed: first use of the copy stream / staging slots
    barrier()
    t0 = time.perf_counter()
    marks = e2e_pass(ke)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = global_batch * ke / e2e_s
    e2e_step_ms = sorted((b - a) * 1e3 for a, b in zip(marks, marks[1:]))
    h2d = h_img.numel() * 4 + h_lab.numel() + h_w.numel() * 4

    # ---- inference half of the BASELINE metric (config 4: overlap-tile, 4096 x 4096, tiles sharded over the ranks)
    inference = None
    if not args.no_inference:
        import bench_infer
        torch.cuda.empty_cache()
        inference = bench_infer.measure(net, dev, world, rank, steps=max(2, min(K, 5)), warmup=3)

    # ---- the "existing Blackwell kernel" bar: stock torch + cuDNN on this same GPU (rank 0, N = 1 only)
    gpu_baseline = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        torch.cuda.empty_cache()
        try:
            v, ms_c = cudnn_steps(per_gpu, 3, 3, dev, "bf16")
            gpu_baseline = {"value": v, "unit": UNIT, "ms_per_step": ms_c, "kind": "torch %s + cuDNN %s, autocast(bf16) + "
                            "channels_last, the reference's operator sequence (oracle/unet_oracle.py), same batch, "
                            "device-resident inputs; `python bench.py --impl cudnn` also reports tf32 and fp32" %
                            (torch.__version__, torch.backends.cudnn.version())}
        except Exception as e:  # noqa: BLE001 -- a baseline failure must not lose the measurement
            gpu_baseline = {"unavailable": repr(e)[:200]}

    if rank == 0:
        tf_peak, hbm_peak, peak_kind = load_peaks()
        ig, wg = fam_ig, fam_wg
        ach = ig[0] / (ig[2] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("igemm_dram_bytes_per_launch"), tj.get("source")
        # every family on its own roof(s): tensor-pipe families by FLOPs, HBM families by algorithmic bytes; the
        # up-convolutions carry both and are reported against the one that binds
        fams = {}
        for name, (fl, by, t_ms, n) in sorted(fam_all.items()):
            e = {"ms_per_step": t_ms / K, "launches_per_step": n / K}
            if fl > 0:
                e["tflops"] = fl / (t_ms * 1e-3) / 1e12
                e["tensor_frac"] = e["tflops"] / tf_peak
            if by > 0:
                e["gbs"] = by / (t_ms * 1e-3) / 1e9
                e["hbm_frac"] = e["gbs"] / hbm_peak
            if fl > 0 and by > 0:
                e["bound"] = "hbm" if e["hbm_frac"] > e["tensor_frac"] else "tensor"
            fams[name] = e
        timed_ms = sum(v[2] for v in fam_all.values()) / K
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(per_gpu, global_batch, world),
                           l2="working set per step (>5 GB of activations) is far larger than the 126 MB L2; no flush needed",
                           timing="CUDA events on the launch stream, barrier+synchronize both sides, max over ranks",
                           step_api="b200.train.TrainStep" + (" (CUDA-graph replay)" if world == 1 else
                                                             " (CUDA-graph replay incl. the gradient all-reduce)" if dp_graph else
                                                             " (eager schedule + gradient all-reduce)"),
                           roofline_timing="per-launch CUDA events over a second pass of the same K steps issued eagerly"),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": ke,
                    "step_ms_min_median_max": [e2e_step_ms[0], e2e_step_ms[len(e2e_step_ms) // 2], e2e_step_ms[-1]]},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "igemm family: igemm_kernel / igemm2_kernel (3x3 conv fwd + dgrad incl. the fused pool / head / bias-sum epilogues, up-conv fwd + dgrad)",
                         "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                         "peak_kind": f"{peak_kind} sustained bf16 (MEASURED_PEAKS.json)", "traffic": traffic,
                         "traffic_source": traffic_src,
                         "launches_timed": ig[3], "ms_per_step": ig[2] / K,
                         "whole_step": {"achieved": FLOPS_PER_TRAIN_TILE * per_gpu * K / (ms * 1e-3) / 1e12,
                                        "frac": FLOPS_PER_TRAIN_TILE * per_gpu * K / (ms * 1e-3) / 1e12 / tf_peak},
                         "wgrad_kernel": {"achieved": wg[0] / (wg[2] * 1e-3) / 1e12,
                                          "frac": wg[0] / (wg[2] * 1e-3) / 1e12 / tf_peak, "ms_per_step": wg[2] / K,
                                          "deterministic": True},
                         "hbm_peak": hbm_peak, "hbm_peak_kind": f"{peak_kind} copy bandwidth (MEASURED_PEAKS.json)",
                         "families": fams,
                         "timed_ms_per_step": timed_ms,
                         "untimed_ms_per_step": ms / K - timed_ms},
            "loss": float(loss.item()),
            "inference": inference,
            "gpu_baseline": gpu_baseline,
            "scaling_breakdown": breakdown,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, dt = cpu_reference_steps(3, 1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "3 train steps on one 572x572 tile, batch 1, fp32 (BASELINE config 1), "
                                             "oracle/unet_oracle.py on all host threads, after 1 warm-up"}
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if world > 1:
        tstep.graphs.clear()            # captured NCCL work must be gone before the communicator is
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
