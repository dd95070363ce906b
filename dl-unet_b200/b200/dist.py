"""Data-parallel training across the GPUs of one box: one process per GPU (torchrun), tiles
partitioned by rank, ONE exchange step -- a sum all-reduce of the flat gradient arena over
NCCL / NVLink, issued bucket by bucket on a side stream while backward is still producing the
later buckets (the arena is laid out in backward-completion order, so buckets fill front to back).

The reference has no multi-device code (main_main.py:157-167); this is SURVEY.md section 8 row (e).
The loss must be normalised by the GLOBAL batch (b200.train.*(global_batch=...)) so that the summed
gradients equal the single-GPU gradients of the concatenated batch.
"""
import os

import torch
import torch.distributed as dist


def partition(n_items, rank, world):
    """Contiguous, balanced split of n_items independent units (tiles) over ranks."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def bucket_bounds(total, ready_ends, bucket_elems):
    """Pure helper (unit-tested on CPU): given the arena offsets at which successive layers complete,
    return the [start, end) slices to all-reduce, each at least bucket_elems long except the last."""
    out, start = [], 0
    for e in ready_ends:
        if e - start >= bucket_elems or e == total:
            if e > start:
                out.append((start, e))
            start = e
    if start < total:
        out.append((start, total))
    return out


class GradAllReducer:
    def __init__(self, unet, bucket_mb=32.0, group=None, reserve_sms=None, reserve_kernels=None):
        """reserve_sms / reserve_kernels: after a bucket's all-reduce is enqueued, the next `reserve_kernels` compute
        launches size their persistent grids for (SM count - reserve_sms) SMs, so the NCCL kernel starts at once on
        free SMs instead of both kernels waiting for each other's CTAs (ops.reserve_sms).  Defaults from the
        environment: UB_DP_RESERVE_SMS (0 = off), UB_DP_RESERVE_KERNELS (2)."""
        self.engine = unet.engine if hasattr(unet, "engine") else unet
        self.group = group
        self.world = dist.get_world_size(group)
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.reserve_sms = int(os.environ.get("UB_DP_RESERVE_SMS", "0")) if reserve_sms is None else reserve_sms
        self.reserve_kernels = int(os.environ.get("UB_DP_RESERVE_KERNELS", "2")) if reserve_kernels is None else reserve_kernels
        self.comm_stream = None
        self.start = 0
        self.works = []
        self.attach()

    def attach(self):
        self.engine.grad_ready_hook = self._on_ready
        self.engine.grad_done_hook = self._on_done

    def detach(self):
        """Single-device schedule again (bench.py's compute-only leg of the scaling breakdown)."""
        self.engine.grad_ready_hook = None
        self.engine.grad_done_hook = None

    def broadcast_parameters(self, src=0):
        st = self.engine.store
        dist.broadcast(st.p_arena, src, group=self.group)
        dist.broadcast(st.m_arena, src, group=self.group)  # momentum must agree too
        st.dirty = True

    def _flush(self, end):
        st = self.engine.store
        if end <= self.start:
            return
        if hasattr(self.engine, "flush_pending_unpack"):
            self.engine.flush_pending_unpack()     # slabs / partial rows of the bucket's layers -> gradients, one launch
        if not st.g_arena.is_cuda:  # host-side logic tests (gloo): same bucketing, no streams
            self.works.append(dist.all_reduce(st.g_arena[self.start:end], group=self.group, async_op=True))
            self.start = end
            return
        if self.comm_stream is None:
            # high priority: at a kernel boundary the pending NCCL blocks are placed before the next compute kernel's
            self.comm_stream = torch.cuda.Stream(device=st.g_arena.device, priority=-1)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self.works.append(dist.all_reduce(st.g_arena[self.start:end], group=self.group, async_op=True))
        self.start = end
        if self.reserve_sms > 0:
            from . import ops
            ops.reserve_sms(self.reserve_sms, self.reserve_kernels)

    def _on_ready(self, s, e):
        if e - self.start >= self.bucket_elems:
            self._flush(e)

    def _on_done(self):
        self._flush(self.engine.store.total)
        if self.reserve_sms > 0 and self.engine.store.g_arena.is_cuda:
            from . import ops
            ops.reserve_sms(0, 0)
        for w in self.works:
            w.wait()  # the compute stream waits for the reduced buckets
        if self.comm_stream is not None:
            # rejoin the side stream (a no-op for ordering -- the waits above cover it -- but a CUDA-graph capture of
            # the step requires every forked stream to join back)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.works = []
        self.start = 0


def make_reducer(unet, group=None, mode=None, bucket_mb=None):
    """The gradient exchange of a data-parallel run.  mode (default: environment UB_DP_ALLREDUCE, else "symm"):
      "symm"  our own all-reduce over torch symmetric memory on the compute stream (b200.symm.SymmAllReducer):
              NVLS multicast (multimem.ld_reduce / multimem.st) where the fabric offers it, peer loads / stores else;
      "peer"  the same without multicast;
      "nccl"  bucketed NCCL all-reduce on a side stream (GradAllReducer) -- also the fallback when symmetric memory
              cannot be set up (no peer access, CPU / gloo groups).
    The engine's arenas must exist (one forward, or unet.engine.store.adopt(device))."""
    mode = mode or os.environ.get("UB_DP_ALLREDUCE", "symm")
    st = (unet.engine if hasattr(unet, "engine") else unet).store
    if mode != "nccl" and st.g_arena is not None and st.g_arena.is_cuda:
        try:
            from .symm import SymmAllReducer
            return SymmAllReducer(unet, group=group, use_multicast=(mode != "peer"))
        except Exception as e:  # noqa: BLE001 -- any failure of the symmetric-memory set-up: use NCCL
            import sys
            print(f"symmetric-memory all-reduce unavailable ({e!r}); using the NCCL all-reducer", file=sys.stderr)
    mb = bucket_mb if bucket_mb is not None else float(os.environ.get("UB_DP_BUCKET_MB", "32"))
    return GradAllReducer(unet, bucket_mb=mb, group=group)
