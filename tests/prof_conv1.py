"""conv11c forward / weight-gradient kernels alone (for ncu): python tests/prof_conv1.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda:0"
img = torch.rand(n, 1, 572, 572, device=dev)
w1, b1 = torch.randn(64, 1, 3, 3, device=dev), torch.randn(64, device=dev)
y = torch.empty(n, 570, 570, 64, dtype=torch.bfloat16, device=dev)
dw, db = torch.zeros(64, 9, device=dev), torch.zeros(64, device=dev)
for _ in range(2):
    ops.conv1_fwd(img, w1, b1, y)
    ops.conv1_wgrad(img, y, dw, db)
torch.cuda.synchronize()
print("ok")
