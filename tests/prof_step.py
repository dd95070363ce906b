"""Minimal driver for ncu: N training steps of the batch-16 572x572 workload (no timing, no baseline).
usage: python tests/prof_step.py [steps] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from network import Unet  # noqa: E402
from b200 import synth, train as T  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Unet().to(dev)
opt = T.FusedSGD(net)
images, labels, weight = synth.synthetic_batch(batch, 572, seed=1)
x, y, w = images.to(dev), labels[:, 0].to(torch.uint8).to(dev), weight.to(dev)
step = T.TrainStep(net, opt, graph=False)      # the schedule bench.py times (there replayed as one CUDA graph), issued eagerly
for i in range(steps):
    loss = step(x, y, w)
torch.cuda.synchronize()
print("loss", loss.item())
