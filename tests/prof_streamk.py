"""Deep layers on the CTA-pair kernel: whole units round-robin vs stream-K (contiguous K-step ranges per pair):
python tests/prof_streamk.py [batch]"""
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


CASES = [("conv31c", 140, 128, 256), ("conv32c", 138, 256, 256), ("conv41c", 68, 256, 512), ("conv42c", 66, 512, 512),
         ("conv51c", 32, 512, 1024), ("conv52c", 30, 1024, 1024), ("conv41e", 56, 1024, 512), ("conv42e", 54, 512, 512),
         ("conv31e", 104, 512, 256), ("conv32e", 102, 256, 256), ("conv21e", 200, 256, 128)]
print(f"{'layer':8s} {'pass':8s} | {'units us':>9s} {'TF/s':>7s} | {'auto us':>9s} {'TF/s':>7s} | {'always us':>9s} {'TF/s':>7s} | max |diff| auto vs units")
tot = [0.0, 0.0, 0.0]
for name, hin, cin, cout in CASES:
    hout = hin - 2
    x, y = rnd(B, hin, hin, cin).clamp_min(0), torch.empty(B, hout, hout, cout, dtype=torch.bfloat16, device=DEV)
    dy, dx = rnd(B, hout, hout, cout), torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=DEV)
    wf, wd = rnd(cout, 9 * cin) / 30, rnd(cin, 9 * cout) / 30
    bias = torch.zeros(cout, device=DEV)
    fl = 2.0 * B * hout * hout * cin * cout * 9
    for what in ("fwd", "dgrad+b"):
        if (what == "fwd" and cout % 256) or (what != "fwd" and cin % 256):
            continue
        res = []
        for mode in (0, 1, 2):
            ops.set_option(13, mode)
            if what == "fwd":
                t = timeit(lambda: ops.conv3x3_fwd(x, wf, bias, y, relu=True))
                out = y.float()
            else:
                part = torch.empty(ops.partial_rows(ops.PROD_DGRAD, B, hin, hin, cin, cout, 0), cin, device=DEV)
                t = timeit(lambda: ops.conv3x3_dgrad(dy, wd, dx, x, dbias_part=part))
                out = dx.float()
            res.append((t, out))
        ops.set_option(13, 1)
        for i in range(3):
            tot[i] += res[i][0]
        print(f"{name:8s} {what:8s} | {res[0][0]:9.1f} {fl / res[0][0] / 1e6:7.1f} | {res[1][0]:9.1f} {fl / res[1][0] / 1e6:7.1f} | "
              f"{res[2][0]:9.1f} {fl / res[2][0] / 1e6:7.1f} | {float((res[0][1] - res[1][1]).abs().max()):.4f}")
for name, hin, cin, cout in [("upconv4", 28, 1024, 512), ("upconv3", 52, 512, 256), ("upconv2", 100, 256, 128), ("upconv1", 196, 128, 64)]:
    x = rnd(B, hin, hin, cin).clamp_min(0)
    cat = torch.zeros(B, 2 * hin, 2 * hin, 2 * cout, dtype=torch.bfloat16, device=DEV)
    wf, wd = rnd(4 * cout, cin) / 30, rnd(cin, 4 * cout) / 30
    bias = torch.zeros(cout, device=DEV)
    dx = torch.empty_like(x)
    fl = 2.0 * B * hin * hin * cin * cout * 4
    for what in ("fwd", "dgrad+b"):
        res = []
        for mode in (0, 1, 2):
            ops.set_option(13, mode)
            if what == "fwd":
                t = timeit(lambda: ops.upconv2x2_fwd(x, wf, bias, cat[..., cout:]))
            else:
                part = torch.empty(ops.partial_rows(ops.PROD_UPDGRAD, B, hin, hin, cin), cin, device=DEV)
                t = timeit(lambda: ops.upconv2x2_dgrad(cat[..., cout:], wd, dx, x, dbias_part=part))
            res.append(t)
        ops.set_option(13, 1)
        for i in range(3):
            tot[i] += res[i]
        print(f"{name:8s} {what:8s} | {res[0]:9.1f} {fl / res[0] / 1e6:7.1f} | {res[1]:9.1f} {fl / res[1] / 1e6:7.1f} | {res[2]:9.1f} {fl / res[2] / 1e6:7.1f} |")
print("totals (ms): units %.3f  auto %.3f  always %.3f" % tuple(t / 1e3 for t in tot))
