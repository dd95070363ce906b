"""Data-parallel correctness on >= 2 GPUs (run under torchrun):
   DP gradients (sum all-reduce of per-rank gradients, loss normalised by the GLOBAL batch) must equal
   the single-GPU gradients of the concatenated batch; weights stay identical across ranks after steps.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_dp_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from network import Unet  # noqa: E402
from b200 import synth, train as T  # noqa: E402
from b200.dist import GradAllReducer, partition  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
S, per = 252, 2
G = per * world
images, labels, weight = synth.synthetic_batch(G, S, seed=11)      # the same global batch on every rank
lo, hi = partition(G, rank, world)

torch.manual_seed(0)
net = Unet().to(dev)
opt = T.FusedSGD(net)
# prime the arenas, then attach the reducer (small buckets so that several all-reduces overlap backward)
opt.zero_grad()
T.weighted_bce_with_logits(net(images[lo:hi].to(dev)), labels[lo:hi].to(dev), weight[lo:hi].to(dev), global_batch=G).backward()
if os.environ.get("UB_DP_ALLREDUCE", "symm") == "nccl":
    red = GradAllReducer(net, bucket_mb=8.0)          # bucketed NCCL all-reduce on a side stream
else:
    from b200.symm import SymmAllReducer
    red = SymmAllReducer(net, use_multicast=os.environ.get("UB_DP_ALLREDUCE", "symm") != "peer")   # own NVLS / peer kernels
    if rank == 0:
        print("symmetric-memory all-reduce, multicast:", red.multicast, flush=True)
torch.manual_seed(0)
ref_sd = Unet().state_dict()
net.load_state_dict(ref_sd)
net.engine.store.first_step = True
red.broadcast_parameters(0)

opt.zero_grad()
loss = T.weighted_bce_with_logits(net(images[lo:hi].to(dev)), labels[lo:hi].to(dev), weight[lo:hi].to(dev), global_batch=G)
loss.backward()
torch.cuda.synchronize()
dp_grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
loss_sum = loss.detach().clone()
dist.all_reduce(loss_sum)

# single-GPU reference on the concatenated batch (every rank computes it; no reducer)
net1 = Unet().to(dev)
net1.load_state_dict(ref_sd)
l1 = T.weighted_bce_with_logits(net1(images.to(dev)), labels.to(dev), weight.to(dev))
l1.backward()
torch.cuda.synchronize()
worst = 0.0
for k, p in net1.named_parameters():
    a, b = dp_grads[k].double(), p.grad.double()
    worst = max(worst, ((a - b).norm() / (b.norm() + 1e-300)).item())
ok = worst < 2e-3 and abs(loss_sum.item() - l1.item()) <= 1e-4 * abs(l1.item())
# two optimizer steps: weights must stay bit-identical across ranks
for _ in range(2):
    opt.step()
    opt.zero_grad()
    T.weighted_bce_with_logits(net(images[lo:hi].to(dev)), labels[lo:hi].to(dev), weight[lo:hi].to(dev), global_batch=G).backward()
opt.step()
torch.cuda.synchronize()
p_arena = net.engine.store.p_arena
mx = p_arena.clone()
mn = p_arena.clone()
dist.all_reduce(mx, op=dist.ReduceOp.MAX)
dist.all_reduce(mn, op=dist.ReduceOp.MIN)
same = bool(torch.equal(mx, mn))
print(f"rank {rank}: worst grad rel_l2 DP vs single = {worst:.3e}; loss dp-sum {loss_sum.item():.6f} vs single {l1.item():.6f}; "
      f"weights identical across ranks after 3 steps: {same}; {'PASS' if ok and same else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok and same else 1)
