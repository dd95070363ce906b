"""Every kernel family once on tiny shapes, for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tests/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
from network import Unet  # noqa: E402
from b200 import ops, synth, train as T  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Unet().to(dev)
opt = T.FusedSGD(net)
images, labels, weight = synth.synthetic_batch(1, 220, seed=3)       # 220 -> 36: crop and pad branches, all kernels
for kind in (T.weighted_bce_with_logits, T.weighted_softmax_ce):
    opt.zero_grad()
    loss = kind(net(images.to(dev)), labels.to(dev), weight.to(dev))
    loss.backward()
    opt.step()
mask = net.predict(images.to(dev), labels.shape[-1])
# every wgrad / conv kernel variant on a small layer
x = torch.randn(1, 20, 36, 128, device=dev).to(torch.bfloat16)
dy = torch.randn(1, 18, 34, 64, device=dev).to(torch.bfloat16)
for mode in (0, 1, 2, 4, 5):
    ops.set_option(1, mode)
    ops.conv3x3_wgrad(x, dy, ops.wgrad_workspace(ops.PACK_CONV_FWD, x, 64))
ops.set_option(1, 3)
for mode in (0, 2):
    ops.set_option(2, mode)
    y = torch.empty(1, 18, 34, 64, dtype=torch.bfloat16, device=dev)
    ops.conv3x3_fwd(x, torch.zeros(64, 9 * 128, dtype=torch.bfloat16, device=dev), torch.zeros(64, device=dev), y)
    dx = torch.empty(1, 20, 36, 128, dtype=torch.bfloat16, device=dev)
    ops.conv3x3_dgrad(dy, torch.zeros(128, 9 * 64, dtype=torch.bfloat16, device=dev), dx, x)
ops.set_option(2, 1)
torch.cuda.synchronize()
print("sanitize_smoke done: loss", float(loss), "mask fg", float(mask.float().mean()))
