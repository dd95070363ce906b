"""Clock probes of the haloed-patch kernel's epilogue (thread 128 of every CTA): cycles waiting for the accumulator
(MMA-bound if large) vs cycles of TMEM load + arithmetic + staging per chunk: python tests/prof_epi_clocks.py [batch]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
# the clock probes live in the diagnostics build only (UB_DIAGNOSTICS); the product library ignores the buffer
os.environ.setdefault("UB_LIB_PATH", os.path.join(ROOT, "dl-unet_b200", "csrc", "libunet_b200_diag.so"))
import torch  # noqa: E402
from b200 import lib, ops  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
buf = torch.zeros(148 * 8, dtype=torch.int64, device=DEV)
ORDERS = (0, 1)   # ub_set_option(17): haloed-patch kernel on one CTA per tile / on CTA pairs
L = lib.load()


def rnd(*shape):
    return torch.randn(*shape, device=DEV).to(torch.bfloat16)


print(f"{'layer':8s} {'pass':8s} | chunks/CTA | wait acc | epilogue work | total cycles | wait % | work %  (mean over CTAs)")
for name, hin, cin, cout in [("conv12c", 570, 64, 64), ("conv21c", 284, 64, 128), ("conv22c", 282, 128, 128), ("conv11e", 392, 128, 64),
                             ("conv12e", 390, 64, 64), ("conv22e", 198, 128, 128)]:
    hout = hin - 2
    x, y = rnd(B, hin, hin, cin).clamp_min(0), torch.empty(B, hout, hout, cout, dtype=torch.bfloat16, device=DEV)
    dy, dx = rnd(B, hout, hout, cout), torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=DEV)
    wf, wd = rnd(cout, 9 * cin) / 30, rnd(cin, 9 * cout) / 30
    bias = torch.zeros(cout, device=DEV)
    for what in ("fwd", "dgrad+b"):
        if what == "fwd":
            fn = lambda: ops.conv3x3_fwd(x, wf, bias, y, relu=True)
        else:
            part = torch.empty(ops.partial_rows(ops.PROD_DGRAD, B, hin, hin, cin, cout, 0), cin, device=DEV)
            fn = lambda: ops.conv3x3_dgrad(dy, wd, dx, x, dbias_part=part)
        for order in ORDERS:
            ops.set_option(17, order)
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            buf.zero_()
            L.ub_set_prof_buffer(C.c_void_p(buf.data_ptr()))
            fn()
            torch.cuda.synchronize()
            L.ub_set_prof_buffer(C.c_void_p(0))
            v = buf.view(148, 8).double()
            n, wait, work, tot = v[:, 5].mean().item(), v[:, 0].mean().item(), v[:, 1].mean().item(), v[:, 6].mean().item()
            ph = [v[:, i].mean().item() / max(n, 1) for i in (2, 3, 4)]
            print(f"{name:8s} {what:8s} pairs {order} {us:7.1f} us | {n:9.1f} | {wait:9.0f} | {work:9.0f} | {tot:9.0f} | {100 * wait / tot:5.1f} | {100 * work / tot:5.1f} | work/chunk {work / max(n, 1):6.0f} cyc, total/chunk {tot / max(n, 1):6.0f}"
                  f" | per chunk: staging-buffer wait {ph[0]:5.0f}, mask + tmem wait {ph[1]:5.0f}, arithmetic + stores {ph[2]:5.0f}")
