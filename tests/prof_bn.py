"""Deep-layer forward / dgrad with the N tile forced to 256 / 128 / 64 (wave quantisation vs operand re-reads):
python tests/prof_bn.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
LAYERS = [("conv41c", 68, 256, 512), ("conv42c", 66, 512, 512), ("conv51c", 32, 512, 1024), ("conv52c", 30, 1024, 1024),
          ("conv41e", 56, 1024, 512), ("conv42e", 54, 512, 512), ("conv31e", 104, 512, 256), ("conv32e", 102, 256, 256),
          ("conv31c", 140, 128, 256), ("conv32c", 138, 256, 256)]


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


for name, hin, cin, cout in LAYERS:
    ho = hin - 2
    x = torch.randn(B, hin, hin, cin, device=DEV).to(torch.bfloat16)
    y = torch.empty(B, ho, ho, cout, dtype=torch.bfloat16, device=DEV)
    dy = torch.randn(B, ho, ho, cout, device=DEV).to(torch.bfloat16)
    dx = torch.empty(B, hin, hin, cin, dtype=torch.bfloat16, device=DEV)
    wf = torch.randn(cout, 9 * cin, device=DEV).to(torch.bfloat16) / 30
    wd = torch.randn(cin, 9 * cout, device=DEV).to(torch.bfloat16) / 30
    bias = torch.zeros(cout, device=DEV)
    fl = 2.0 * B * ho * ho * cin * cout * 9
    row = f"{name:8s}"
    for bn, pair in ((0, 1), (0, 0), (128, 0)):
        ops.set_tuning(bn=bn)
        ops.set_option(11, pair)
        tf = timeit(lambda: ops.conv3x3_fwd(x, wf, bias, y))
        td = timeit(lambda: ops.conv3x3_dgrad(dy, wd, dx, x))
        row += f" | bn{bn:<3d} pair{pair} fwd {tf:6.1f} us {fl / tf / 1e6:6.0f} TF  dgrad {td:6.1f} us {fl / td / 1e6:6.0f} TF"
    ops.set_tuning()
    ops.set_option(11, 1)
    print(row, flush=True)
