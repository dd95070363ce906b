"""Every kernel family once on tiny shapes, for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tests/sanitize_smoke.py
(On the round-2 GPU pool compute-sanitizer is closed; the script still runs plain as a launch-error / dead-lock-guard check
of every kernel variant: each launch is checked through the C ABI's return code and ub_device_error.)"""
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
# round 2: haloed-patch kernel on CTA pairs with an odd tile count, fused column sums and fused pool; CTA-pair kernel with
# boxes that span several images (ragged batch: part of the last box lies beyond the tensor); up-convolution the same
ops.set_option(2, 2)
xs = torch.randn(3, 27, 35, 64, device=dev).clamp_min(0).to(torch.bfloat16)
dys = torch.randn(3, 25, 33, 128, device=dev).to(torch.bfloat16)
ops.conv3x3_dgrad(dys, torch.zeros(64, 9 * 128, dtype=torch.bfloat16, device=dev), torch.empty_like(xs), xs, torch.zeros(64, device=dev))
pooled = torch.empty(3, 12, 16, 128, dtype=torch.bfloat16, device=dev)
idx = torch.empty(3, 12, 16, 128, dtype=torch.uint8, device=dev)
ops.conv3x3_fwd_pool(xs, torch.zeros(128, 9 * 64, dtype=torch.bfloat16, device=dev), torch.zeros(128, device=dev), 128, None, pooled, idx, None, 0)
ops.set_option(2, 1)
ops.set_tuning(bn=256)
x5 = torch.randn(5, 30, 30, 256, device=dev).clamp_min(0).to(torch.bfloat16)
y5 = torch.empty(5, 28, 28, 256, dtype=torch.bfloat16, device=dev)
ops.conv3x3_fwd(x5, torch.zeros(256, 9 * 256, dtype=torch.bfloat16, device=dev), torch.zeros(256, device=dev), y5)
ops.conv3x3_dgrad(y5, torch.zeros(256, 9 * 256, dtype=torch.bfloat16, device=dev), torch.empty_like(x5), x5, torch.zeros(256, device=dev))
ops.set_tuning()
xu = torch.randn(6, 28, 28, 512, device=dev).clamp_min(0).to(torch.bfloat16)
cat = torch.zeros(6, 56, 56, 512, dtype=torch.bfloat16, device=dev)
ops.upconv2x2_fwd(xu, torch.zeros(4 * 256, 512, dtype=torch.bfloat16, device=dev), torch.zeros(256, device=dev), cat[..., 256:])
ops.upconv2x2_dgrad(cat[..., 256:], torch.zeros(512, 4 * 256, dtype=torch.bfloat16, device=dev), torch.empty_like(xu), xu, dbias=torch.zeros(512, device=dev))
torch.cuda.synchronize()
print("sanitize_smoke done: loss", float(loss), "mask fg", float(mask.float().mean()))
