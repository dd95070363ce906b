"""Per-layer timing table at the benchmark batch (CUDA events, isolated kernels):
python tests/prof_layers.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402
from b200.geometry import Geometry, LAYERS  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = Geometry(572)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def rnd(*shape):
    return torch.randn(*shape, device=DEV).to(torch.bfloat16)


tot = {}
print(f"{'layer':10s} {'Hin':>4s} {'Cin':>5s} {'Cout':>5s} | {'fwd us':>8s} {'TF/s':>7s} | {'dgrad us':>8s} {'TF/s':>7s} | {'wg0 us':>8s} {'TF/s':>7s} | {'wg2 us':>8s} {'TF/s':>7s} | {'wg5 us':>8s} {'TF/s':>7s}")
for (name, kind, cin, cout) in LAYERS:
    if kind == "conv":
        lvl, pos, side = int(name[4]), int(name[5]), name[6]
        hout = ((g.enc1 if pos == 1 else g.enc2) if side == "c" else (g.dec1 if pos == 1 else g.dec2))[lvl]
        hin = hout + 2
        x, y = rnd(B, hin, hin, cin), torch.empty(B, hout, hout, cout, dtype=torch.bfloat16, device=DEV)
        dy, dx = rnd(B, hout, hout, cout), torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=DEV)
        wf, wd = rnd(cout, 9 * cin) / 30, rnd(cin, 9 * cout) / 30
        bias = torch.zeros(cout, device=DEV)
        fl = 2.0 * B * hout * hout * cin * cout * 9

        def wg():   # the slab count depends on the kernel the current option selects
            dwp = ops.wgrad_workspace(ops.PACK_CONV_FWD, x, cout)
            return timeit(lambda: ops.conv3x3_wgrad(x, dy, dwp))

        t_f = timeit(lambda: ops.conv3x3_fwd(x, wf, bias, y))
        t_d = timeit(lambda: ops.conv3x3_dgrad(dy, wd, dx, x))
        ops.set_option(1, 0)
        t_w0 = wg()
        ops.set_option(1, 2)
        t_w2 = wg()
        ops.set_option(1, 4)
        t_w5 = wg()
        ops.set_option(1, 5)
        t_w4 = wg() if cout == 64 else 0.0
        ops.set_option(1, 3)
        t_auto = wg()
    elif kind == "up":
        lvl = int(name[-1])
        hin = g.up[lvl] // 2
        x = rnd(B, hin, hin, cin)
        cat, dcat = torch.zeros(B, 2 * hin, 2 * hin, 2 * cout, dtype=torch.bfloat16, device=DEV), rnd(B, 2 * hin, 2 * hin, 2 * cout)
        dx = torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=DEV)
        wf, wd = rnd(4 * cout, cin) / 30, rnd(cin, 4 * cout) / 30
        bias = torch.zeros(cout, device=DEV)
        dwp = ops.wgrad_workspace(ops.PACK_UP_FWD, x, cout)
        fl = 2.0 * B * hin * hin * cin * cout * 4
        t_f = timeit(lambda: ops.upconv2x2_fwd(x, wf, bias, cat[..., cout:]))
        t_d = timeit(lambda: ops.upconv2x2_dgrad(dcat[..., cout:], wd, dx, x))
        t_w0 = t_w2 = t_w5 = t_auto = timeit(lambda: ops.upconv2x2_wgrad(x, dcat[..., cout:], dwp))
        t_w4 = 0.0
    else:
        continue
    for k, v in (("fwd", t_f), ("dgrad", t_d), ("wg0", t_w0), ("wg2", t_w2), ("wg5", t_w5), ("wgbest", min(t_w0, t_w2, t_w5)), ("wgauto", t_auto)):
        tot[k] = tot.get(k, 0.0) + v
    print(f"{name:10s} {hin:4d} {cin:5d} {cout:5d} | {t_f:8.1f} {fl / t_f / 1e6:7.1f} | {t_d:8.1f} {fl / t_d / 1e6:7.1f} | "
          f"{t_w0:8.1f} {fl / t_w0 / 1e6:7.1f} | {t_w2:8.1f} {fl / t_w2 / 1e6:7.1f} | {t_w5:8.1f} {fl / t_w5 / 1e6:7.1f} | w4 {t_w4:7.1f} auto {t_auto:7.1f}", flush=True)
print("totals (ms):", {k: round(v / 1e3, 2) for k, v in tot.items()})

# conv12e + finalconv: fused epilogue head vs conv -> bf16 store -> head kernel
x = rnd(B, 390, 390, 64)
y = torch.empty(B, 388, 388, 64, dtype=torch.bfloat16, device=DEV)
wf = rnd(64, 9 * 64) / 30
bias = torch.zeros(64, device=DEV)
hw, hb = torch.randn(2, 64, device=DEV) / 8, torch.zeros(2, device=DEV)
lg = torch.empty(B, 2, 388, 388, device=DEV)
am = torch.empty(B, 388, 388, dtype=torch.uint8, device=DEV)
t_plain = timeit(lambda: ops.conv3x3_fwd(x, wf, bias, y))
t_head = timeit(lambda: ops.head_fwd(y, hw, hb, lg, None))
t_fused = timeit(lambda: ops.conv3x3_fwd_head(x, wf, bias, y, hw, hb, lg, None))
t_inf = timeit(lambda: ops.conv3x3_fwd_head(x, wf, bias, None, hw, hb, None, am))
print(f"conv12e+head: conv {t_plain:.1f} us + head {t_head:.1f} us = {t_plain + t_head:.1f} us; fused (train) {t_fused:.1f} us; "
      f"fused, no activation store, argmax only (inference) {t_inf:.1f} us")

