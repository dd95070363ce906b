"""256-channel layers (N = 256) on the haloed-patch kernel (igemm2, one CTA per tile) vs the tap-per-load kernel on CTA
pairs: python tests/prof_pair256.py [batch]"""
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


CASES = [("conv31c", 140, 128, 256), ("conv32c", 138, 256, 256), ("conv32e", 102, 256, 256), ("conv21e", 200, 256, 128),
         ("conv31e", 104, 512, 256), ("conv22c", 282, 128, 128)]
print(f"{'layer':8s} {'pass':6s} | {'igemm2 us':>9s} {'TF/s':>7s} | {'pairs us':>9s} {'TF/s':>7s} | equal")
for name, hin, cin, cout in CASES:
    hout = hin - 2
    x, y = rnd(B, hin, hin, cin).clamp_min(0), torch.empty(B, hout, hout, cout, dtype=torch.bfloat16, device=DEV)
    dy, dx = rnd(B, hout, hout, cout), torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=DEV)
    wf, wd = rnd(cout, 9 * cin) / 30, rnd(cin, 9 * cout) / 30
    bias = torch.zeros(cout, device=DEV)
    part = torch.empty(ops.partial_rows(ops.PROD_DGRAD, B, hin, hin, cin, cout, 0), cin, device=DEV)
    fl = 2.0 * B * hout * hout * cin * cout * 9
    for what in ("fwd", "dgrad", "dgrad+b"):
        res = []
        for mode in (1, 0):
            ops.set_option(2, mode)
            if what == "fwd":
                t = timeit(lambda: ops.conv3x3_fwd(x, wf, bias, y, relu=True))
                out = y.clone()
            elif what == "dgrad":
                t = timeit(lambda: ops.conv3x3_dgrad(dy, wd, dx, x))
                out = dx.clone()
            else:
                part = torch.empty(ops.partial_rows(ops.PROD_DGRAD, B, hin, hin, cin, cout, 0), cin, device=DEV)
                t = timeit(lambda: ops.conv3x3_dgrad(dy, wd, dx, x, dbias_part=part))
                out = dx.clone()
            res.append((t, out))
        ops.set_option(2, 1)
        print(f"{name:8s} {what:7s} | {res[0][0]:9.1f} {fl / res[0][0] / 1e6:7.1f} | {res[1][0]:9.1f} {fl / res[1][0] / 1e6:7.1f} | "
              f"{bool(torch.equal(res[0][1], res[1][1]))}")
