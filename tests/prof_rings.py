"""CTA-pair kernel: 5 operand stages + 4 output staging buffers vs 3 + 8 (option 14): python tests/prof_rings.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def timeit(fn, reps=10):
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


print(f"{'layer':8s} {'pass':8s} | {'<5,4> us':>9s} | {'<3,8> us':>9s} | auto")
for name, hin, cin, cout in [("upconv4", 28, 1024, 512), ("upconv3", 52, 512, 256), ("upconv2", 100, 256, 128), ("upconv1", 196, 128, 64)]:
    x = rnd(B, hin, hin, cin).clamp_min(0)
    cat = torch.zeros(B, 2 * hin, 2 * hin, 2 * cout, dtype=torch.bfloat16, device=DEV)
    wf, wd = rnd(4 * cout, cin) / 30, rnd(cin, 4 * cout) / 30
    bias = torch.zeros(cout, device=DEV)
    dx = torch.empty_like(x)
    for what in ("fwd", "dgrad", "dgrad+b"):
        res = []
        for mode in (1, 2, 0):
            ops.set_option(14, mode)
            if what == "fwd":
                t = timeit(lambda: ops.upconv2x2_fwd(x, wf, bias, cat[..., cout:]))
            elif what == "dgrad":
                t = timeit(lambda: ops.upconv2x2_dgrad(cat[..., cout:], wd, dx, x))
            else:
                part = torch.empty(ops.partial_rows(ops.PROD_UPDGRAD, B, hin, hin, cin), cin, device=DEV)
                t = timeit(lambda: ops.upconv2x2_dgrad(cat[..., cout:], wd, dx, x, dbias_part=part))
            res.append(t)
        ops.set_option(14, 0)
        print(f"{name:8s} {what:8s} | {res[0]:9.1f} | {res[1]:9.1f} | {res[2]:9.1f}")
for name, hin, cin, cout in [("conv31c", 140, 128, 256), ("conv32c", 138, 256, 256), ("conv42c", 66, 512, 512), ("conv52c", 30, 1024, 1024),
                             ("conv41e", 56, 1024, 512), ("conv32e", 102, 256, 256), ("conv21e", 200, 256, 128)]:
    hout = hin - 2
    x, y = rnd(B, hin, hin, cin).clamp_min(0), torch.empty(B, hout, hout, cout, dtype=torch.bfloat16, device=DEV)
    dy, dx = rnd(B, hout, hout, cout), torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=DEV)
    wf, wd = rnd(cout, 9 * cin) / 30, rnd(cin, 9 * cout) / 30
    bias = torch.zeros(cout, device=DEV)
    for what in ("fwd", "dgrad+b"):
        if (what == "fwd" and cout % 256) or (what != "fwd" and cin % 256):
            continue
        res = []
        for mode in (1, 2, 0):
            ops.set_option(14, mode)
            if what == "fwd":
                t = timeit(lambda: ops.conv3x3_fwd(x, wf, bias, y, relu=True))
            else:
                part = torch.empty(ops.partial_rows(ops.PROD_DGRAD, B, hin, hin, cin, cout, 0), cin, device=DEV)
                t = timeit(lambda: ops.conv3x3_dgrad(dy, wd, dx, x, dbias_part=part))
            res.append(t)
        ops.set_option(14, 0)
        print(f"{name:8s} {what:8s} | {res[0]:9.1f} | {res[1]:9.1f} | {res[2]:9.1f}")
