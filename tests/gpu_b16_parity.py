"""Parity numbers of the benchmark configuration (16 tiles of 572 x 572) against the fp32 oracle evaluated by stock
PyTorch on the GPU (TF32 off): per-tile and whole-batch logits error, loss, gradients, argmax agreement.
python tests/gpu_b16_parity.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dl-unet_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402
import unet_oracle as O  # noqa: E402
from network import Unet  # noqa: E402
from b200 import train as T  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = {k: (v.to(torch.bfloat16).float() if k.endswith("weight") and k not in ("conv11c.weight", "finalconv.weight") else v)
      for k, v in O.init_state_dict(0).items()}
images, labels, weight = O.synthetic_batch(B, 572, seed=1000)
images = images.to(torch.bfloat16).float()
sd_d = {k: v.to(DEV) for k, v in sd.items()}
oloss, ologits, ograds = O.train_step(sd_d, {}, images.to(DEV), labels.to(DEV), weight.to(DEV))
net = Unet()
net.load_state_dict(sd)
net.to(DEV)
step = T.TrainStep(net, T.FusedSGD(net), graph=False)
loss = step(images.to(DEV), labels.to(DEV), weight.to(DEV))
torch.cuda.synchronize()
lg = step.logits.double()
og = ologits.double()
d = (lg - og).abs()
per_tile = [(d[i].max() / og[i].abs().max()).item() for i in range(B)]
print("per-tile max-rel:", " ".join(f"{v:.2e}" for v in per_tile))
print(f"whole-batch max-rel {(d.max() / og.abs().max()).item():.3e}  rel-L2 {((lg - og).norm() / og.norm()).item():.3e}  "
      f"|ref| max {og.abs().max().item():.1f} rms {og.pow(2).mean().sqrt().item():.1f}")
print(f"loss ours {loss.item():.6f} oracle {oloss.item():.6f} rel {abs(loss.item() - oloss.item()) / abs(oloss.item()):.2e}")
agree = (lg.argmax(1) == og.argmax(1)).double()
print("argmax agreement per tile:", " ".join(f"{agree[i].mean().item():.5f}" for i in range(B)), f" all {agree.mean().item():.5f}")
worst = max(((p.grad.double() - ograds[k].double()).norm() / (ograds[k].double().norm() + 1e-300)).item() for k, p in net.named_parameters())
print(f"worst gradient rel-L2 {worst:.3e}")
