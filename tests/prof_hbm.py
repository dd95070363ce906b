"""The one-launch bookkeeping kernels of a batch-16 training step, each alone with a cold L2 (CUDA events around every
call, a 512 MB write between calls): python tests/prof_hbm.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
from b200 import ops  # noqa: E402
from b200.train import FusedSGD, TrainStep  # noqa: E402
from network import Unet  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
net = Unet().to(DEV)
opt = FusedSGD(net, lr=1e-4, momentum=0.99)
step = TrainStep(net, opt, graph=False)
img = torch.rand(B, 1, 572, 572, device=DEV)
lab = (torch.rand(B, 388, 388, device=DEV) > 0.5).long()
wm = torch.rand(B, 388, 388, device=DEV) + 0.5
for _ in range(2):
    step(img, lab, wm)
torch.cuda.synchronize()
eng = net.engine
st = eng.store
flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, reps=8):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


table = eng._slabs["table"]
slab_bytes = eng._slabs["arena"].numel() * 4
ws = eng.workspace(B, 572, torch.device(DEV), True)
a = ws.act
cases = [
    ("unpack_all", lambda: table.unpack_all(False), slab_bytes + 4 * st.total),
    ("pack_all", lambda: st._job_table().pack_all(), 8 * st.total),
    ("sgd", lambda: ops.sgd_momentum(st.p_arena, st.g_arena, st.m_arena, 0.0, 0.99, False), 20 * st.total),
    ("conv1_fwd", lambda: ops.conv1_fwd(img, st.p("conv11c.weight"), st.p("conv11c.bias"), a["e11"]), img.numel() * 4 + a["e11"].numel() * 2),
]
for name, fn, nbytes in cases:
    med, best = timeit(fn)
    print(f"{name:12s} median {med:7.1f} us  best {best:7.1f} us  {nbytes / 1e6:7.1f} MB  {nbytes / med / 1e3:7.1f} GB/s (median)")

# per job of the unpack table: where the one-launch kernel spends its time
print("unpack jobs (alone, cold L2):")
tot = 0.0
for i, e in enumerate(table.keep):
    nbytes = e[4].numel() * 4.0 * (e[8] + 1)
    med, best = timeit(lambda: table.unpack_layer(i), reps=4)
    tot += med
    print(f"  job {i:3d} kind {e[5]} cout {e[6]:5d} cin {e[7]:6d} slabs {e[8]:5d}  {med:7.1f} us  {nbytes / 1e6:7.2f} MB  {nbytes / med / 1e3:7.1f} GB/s")
print(f"  sum of the jobs alone: {tot:.1f} us")
