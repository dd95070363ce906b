"""Data-parallel gradient exchange through NVLink / NVSwitch symmetric memory (csrc/allreduce.cuh, ub_allreduce_sym):
the gradient arena of every rank is allocated as torch symmetric memory, so every rank maps every peer's arena and --
on NVSwitch -- one multicast address for all of them; after backward each rank reduces 1/N of the arena with
multimem.ld_reduce (the switch adds the N copies) and broadcasts it with multimem.st.  Everything runs on the compute
stream: the data-parallel step is the single-GPU kernel schedule plus four small launches, deterministic, and captures
into one CUDA graph.  SURVEY.md section 8 row e-train; the bucketed NCCL all-reduce (b200.dist.GradAllReducer) is the
alternative for process groups without peer access."""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib as _lib
from .ops import STATS, check

FLAG_WORDS = 2 * 32 * 16           # produced / done x buckets x ranks (csrc/allreduce.cuh)


class SymmAllReducer:
    def __init__(self, unet, group=None, use_multicast=True):
        import torch.distributed._symmetric_memory as symm_mem
        self.engine = unet.engine if hasattr(unet, "engine") else unet
        st = self.engine.store
        if st.g_arena is None or not st.g_arena.is_cuda:
            raise RuntimeError("SymmAllReducer needs the engine's CUDA arenas: run one forward first")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        dev = st.g_arena.device
        # the gradient arena moves into symmetric memory (same layout; views and job tables are rebuilt on the next step)
        arena = symm_mem.empty(st.total, dtype=torch.float32, device=dev)
        flags = symm_mem.empty(FLAG_WORDS, dtype=torch.int32, device=dev)
        arena.copy_(st.g_arena)
        flags.zero_()
        torch.cuda.synchronize(dev)
        self.h_arena = symm_mem.rendezvous(arena, self.group)
        self.h_flags = symm_mem.rendezvous(flags, self.group)
        self.arena, self.flags = arena, flags
        mc = int(getattr(self.h_arena, "multicast_ptr", 0) or 0) if use_multicast else 0   # 0: no multicast object (no NVSwitch)
        self.multicast = mc != 0
        self._peer = (C.c_void_p * self.world)(*[int(p) for p in self.h_arena.buffer_ptrs])
        self._flag = (C.c_void_p * self.world)(*[int(p) for p in self.h_flags.buffer_ptrs])
        self._mc = C.c_void_p(mc)
        self.ctrs = torch.zeros(2, dtype=torch.int32, device=dev)      # step counter, finished-CTA counter
        for p in unet.parameters():
            p.grad = None
        st.g_arena = arena
        st.arena_id += 1                                               # workspaces rebuild the tables that point into it
        dist.barrier(self.group)                                       # every rank's flag pad is zero before anyone signals
        self.attach()

    def attach(self):
        self.engine.grad_ready_hook = None      # single-device schedule: one unpack launch at the end of backward ...
        self.engine.grad_done_hook = self._on_done   # ... then the exchange

    def detach(self):
        self.engine.grad_ready_hook = None
        self.engine.grad_done_hook = None

    def broadcast_parameters(self, src=0):
        st = self.engine.store
        dist.broadcast(st.p_arena, src, group=self.group)
        dist.broadcast(st.m_arena, src, group=self.group)
        st.dirty = True

    def _on_done(self):
        st = self.engine.store
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = _lib.load().ub_allreduce_sym(self._peer, self._flag, self._mc, C.c_void_p(self.ctrs.data_ptr()),
                                          C.c_void_p(self.ctrs.data_ptr() + 4), self.rank, self.world, 0, st.total, 0, 1,
                                          1 | 2 | 4 | 8, stream)
        check(rc, "allreduce_sym")
        STATS.launches += 4
