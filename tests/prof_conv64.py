"""The 64 -> 64 channel level-1 convolution alone (for ncu): python tests/prof_conv64.py [igemm3: 0|1] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = "cuda:0"
ops.set_option(3, 2 if mode else 0)
hout = 568
x = torch.randn(B, hout + 2, hout + 2, 64, device=dev).to(torch.bfloat16)
y = torch.empty(B, hout, hout, 64, dtype=torch.bfloat16, device=dev)
dy = torch.randn(B, hout, hout, 64, device=dev).to(torch.bfloat16)
dx = torch.empty(B, hout + 2, hout + 2, 64, dtype=torch.bfloat16, device=dev)
wf = torch.randn(64, 9 * 64, device=dev).to(torch.bfloat16) / 30
bias = torch.zeros(64, device=dev)
for _ in range(2):
    ops.conv3x3_fwd(x, wf, bias, y)
    ops.conv3x3_dgrad(dy, wf, dx, x)
torch.cuda.synchronize()
print("ok")
if "--sweep" in sys.argv:
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
    for dbg in (0, 4, 4 + 8, 4 + 16, 4 + 8 + 16):
        ops.set_option(4, dbg)
        print(f"dbg={dbg}: fwd {timeit(lambda: ops.conv3x3_fwd(x, wf, bias, y)):.1f} us  dgrad {timeit(lambda: ops.conv3x3_dgrad(dy, wf, dx, x)):.1f} us", flush=True)
    ops.set_option(4, 0)

if "--phases" in sys.argv:
    # epilogue phase clocks of the haloed-patch kernel (thread 128 of every CTA), forward and dgrad
    from b200 import lib
    ops.set_option(3, 0)
    buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    for name, fn in (("fwd", lambda: ops.conv3x3_fwd(x, wf, bias, y)), ("dgrad", lambda: ops.conv3x3_dgrad(dy, wf, dx, x)),
                     ("dgrad, no mask", lambda: ops.conv3x3_dgrad(dy, wf, dx, None)),
                     ("dgrad, no L2 prefetch of the mask", lambda: (ops.set_option(6, 0), ops.conv3x3_dgrad(dy, wf, dx, x), ops.set_option(6, 1)))):
        buf.zero_()
        lib.load().ub_set_prof_buffer(buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.load().ub_set_prof_buffer(None)
        m = buf.view(148, 8).double().mean(0).tolist()
        n = max(m[5], 1.0)
        print(f"{name}: chunks/CTA {m[5]:.0f}  total {m[6] / n:.0f} clk/chunk = wait MMA {m[0] / n:.0f} + tmem+math+stage {m[1] / n:.0f} "
              f"+ fence/store-wait {m[2] / n:.0f} + barrier {m[3] / n:.0f} + store issue {m[4] / n:.0f} + rest "
              f"{(m[6] - sum(m[:5])) / n:.0f}", flush=True)
if "--c128" in sys.argv:
    # conv21c dgrad shape: 128 -> 64 channels on a 282 -> 284 map (stationary weights, two k-chunks, two patch slots)
    from b200 import lib
    dy2 = torch.randn(B, 282, 282, 128, device=dev).to(torch.bfloat16)
    dx2 = torch.empty(B, 284, 284, 64, dtype=torch.bfloat16, device=dev)
    wd2 = torch.randn(64, 9 * 128, device=dev).to(torch.bfloat16) / 30
    buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    fn = lambda: ops.conv3x3_dgrad(dy2, wd2, dx2, None)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"128->64 dgrad 284: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us")
    lib.load().ub_set_prof_buffer(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.load().ub_set_prof_buffer(None)
    m = buf.view(148, 8).double().mean(0).tolist()
    n = max(m[5], 1.0)
    print(f"chunks/CTA {m[5]:.0f}  total {m[6] / n:.0f} clk/chunk = wait MMA {m[0] / n:.0f} + tmem+math+stage {m[1] / n:.0f} + rest {(m[6] - m[0] - m[1]) / n:.0f}")
