"""Weight-map kernels under ncu / CUDA events: python tests/prof_weightmap.py [batch] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import weightmap  # noqa: E402
from b200.synth import synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
_, labels, _ = synthetic_batch(B, 572, seed=5)
lab = labels[:, 0].contiguous().cuda()
out = torch.empty(B, 388, 388, device="cuda")
for _ in range(reps):
    weightmap.weighted_map(lab, check=False, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    weightmap.weighted_map(lab, check=False, out=out)
e1.record()
torch.cuda.synchronize()
print(f"weighted_map B={B}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
