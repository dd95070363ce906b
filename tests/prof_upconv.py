"""One up-convolution layer alone (for ncu / timing): python tests/prof_upconv.py [level 1-4] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402
from b200.geometry import Geometry  # noqa: E402

lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = "cuda:0"
g = Geometry(572)
cout = {1: 64, 2: 128, 3: 256, 4: 512}[lvl]
cin = 2 * cout
hin = g.up[lvl] // 2
x = torch.randn(B, hin, hin, cin, device=dev).to(torch.bfloat16)
cat = torch.zeros(B, 2 * hin, 2 * hin, 2 * cout, dtype=torch.bfloat16, device=dev)
dcat = torch.randn(B, 2 * hin, 2 * hin, 2 * cout, device=dev).to(torch.bfloat16)
dx = torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=dev)
wf = torch.randn(4 * cout, cin, device=dev).to(torch.bfloat16) / 30
wd = torch.randn(cin, 4 * cout, device=dev).to(torch.bfloat16) / 30
bias = torch.zeros(cout, device=dev)
dwp = ops.wgrad_workspace(ops.PACK_UP_FWD, x, cout)


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


fl = 2.0 * B * hin * hin * cin * cout * 4
for name, fn in (("fwd", lambda: ops.upconv2x2_fwd(x, wf, bias, cat[..., cout:])),
                 ("dgrad", lambda: ops.upconv2x2_dgrad(dcat[..., cout:], wd, dx, x)),
                 ("wgrad", lambda: ops.upconv2x2_wgrad(x, dcat[..., cout:], dwp))):
    t = timeit(fn)
    print(f"upconv{lvl} {name}: {t:.1f} us  {fl / t / 1e6:.0f} TF/s", flush=True)
# the same GEMM with a DENSE output (no pixel shuffle): what the store pattern costs
dense = torch.zeros(B, 2 * hin, 2 * hin, cout, dtype=torch.bfloat16, device=dev)
t = timeit(lambda: ops.upconv2x2_fwd(x, wf, bias, dense))
print(f"upconv{lvl} fwd into a dense {cout}-channel buffer: {t:.1f} us")
