"""Kernel micro-benchmarks for ncu: python tests/prof_kernels.py <case> ; cases: wg52 (conv52c wgrad),
wg12 (conv12c wgrad), fw12 (conv12c fwd), fw42 (conv42c fwd), wg22 (conv22c wgrad)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

DEV = "cuda:0"
CASES = {"wg52": (16, 30, 1024, 1024), "wg42": (16, 66, 512, 512), "wg41": (16, 68, 256, 512), "wg41e": (16, 56, 1024, 512),
         "fw52": (16, 30, 1024, 1024), "fw41e": (16, 56, 1024, 512), "wg12": (16, 570, 64, 64), "wg22": (16, 282, 128, 128),
         "fw12": (16, 570, 64, 64), "fw42": (16, 66, 512, 512), "fw22": (16, 282, 128, 128),
         "fwA": (16, 200, 64, 64), "fwB": (16, 200, 128, 128),
         "dg12": (16, 570, 64, 64), "dg11e": (16, 392, 128, 64), "up1": (16, 196, 128, 64), "up1d": (16, 196, 128, 64)}
case = sys.argv[1]
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, h, cin, cout = CASES[case]
torch.manual_seed(0)
x = torch.randn(n, h, h, cin, device=DEV).to(torch.bfloat16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if case.startswith("dg"):
    dy = torch.randn(n, h - 2, h - 2, cout, device=DEV).to(torch.bfloat16)
    wd = torch.randn(cin, 9 * cout, device=DEV).to(torch.bfloat16) / 30
    dx = torch.empty(n, h, h, cin, dtype=torch.bfloat16, device=DEV)
    xm = x.clamp_min(0)
    msk = xm if mode != 9 else None
    ops.conv3x3_dgrad(dy, wd, dx, msk)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.conv3x3_dgrad(dy, wd, dx, msk)
    e1.record()
elif case.startswith("up"):
    cat = torch.zeros(n, 2 * h, 2 * h, 2 * cout, dtype=torch.bfloat16, device=DEV)
    wf = torch.randn(4 * cout, cin, device=DEV).to(torch.bfloat16) / 30
    wdd = torch.randn(cin, 4 * cout, device=DEV).to(torch.bfloat16) / 30
    b = torch.zeros(cout, device=DEV)
    dx = torch.empty(n, h, h, cin, dtype=torch.bfloat16, device=DEV)
    fn = (lambda: ops.upconv2x2_fwd(x, wf, b, cat[..., cout:])) if case == "up1" else (lambda: ops.upconv2x2_dgrad(cat[..., cout:], wdd, dx, x))
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
elif case.startswith("wg"):
    dy = torch.randn(n, h - 2, h - 2, cout, device=DEV).to(torch.bfloat16)
    ops.set_option(1, mode)
    dwp = ops.wgrad_workspace(ops.PACK_CONV_FWD, x, cout)
    print("slabs", dwp.shape[0])
    ops.conv3x3_wgrad(x, dy, dwp)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.conv3x3_wgrad(x, dy, dwp)
    e1.record()
else:
    ops.set_option(2, mode)   # for forward cases "mode" selects the conv kernel (0 = igemm, 2 = igemm2)
    w = torch.randn(cout, 9 * cin, device=DEV).to(torch.bfloat16) / 30
    b = torch.zeros(cout, device=DEV)
    y = torch.empty(n, h - 2, h - 2, cout, dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_fwd(x, w, b, y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.conv3x3_fwd(x, w, b, y)
    e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
fl = 2.0 * n * (h - 2) ** 2 * cin * cout * 9 if not case.startswith('up') else 2.0 * n * h * h * cin * cout * 4
print(f"{case} mode={mode}: {us:.1f} us  {fl / us / 1e6:.1f} TF/s")
