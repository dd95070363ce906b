"""Loss and optimizer pieces of the training hot loop (trainer.py:30, 54-78) on the CUDA kernels:
a fused weighted-loss autograd node and an optim.SGD subclass whose step() is one kernel over
the engine's flat arenas."""
import torch

from . import ops


class _WeightedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label, weight, kind, global_batch, workspace):
        b, _, hp, wp = logits.shape
        l = label.shape[-1]
        nb = global_batch if global_batch else b
        scale = 1.0 / ((2.0 if kind == ops.LOSS_BCE else 1.0) * nb * l * l)
        logits = logits.contiguous()
        pad_border = hp != l
        dlogits = (torch.zeros_like(logits) if pad_border else torch.empty_like(logits))
        out = torch.empty(1, dtype=torch.float32, device=logits.device)
        ops.loss_fwd_bwd(kind, logits, label, weight, dlogits, out, workspace, scale)
        ctx.save_for_backward(dlogits)
        loss = out[0]
        if nb != b:  # the kernel normalised by the local batch; report the share of the global mean
            loss = loss * (b / nb)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None, None, None, None


_loss_ws = {}


def _prep(logits, labels, weight):
    if not logits.is_cuda:
        raise RuntimeError("the fused loss has no CPU implementation")
    dev = logits.device
    if labels.dim() == 4:
        labels = labels[:, 0]
    labels = labels.to(dev)
    if labels.dtype not in (torch.uint8, torch.int64):
        labels = labels.to(torch.int64)
    labels = labels.contiguous()
    weight = weight.to(device=dev, dtype=torch.float32).contiguous()
    if weight.shape != labels.shape:
        raise RuntimeError(f"weight map {tuple(weight.shape)} must match the label {tuple(labels.shape)} "
                           "(per-pixel broadcast; SURVEY.md 0.1-3)")
    if (logits.shape[-1] - labels.shape[-1]) % 2:
        raise RuntimeError("logits / label size difference must be even for the centre crop (trainer.py:60)")
    ws = _loss_ws.get(dev)
    if ws is None:
        ws = _loss_ws[dev] = ops.loss_workspace(dev)
    return labels, weight, ws


def weighted_bce_with_logits(logits, labels, weight, global_batch=None):
    """trainer.py:60-75 in one kernel: centre-crop of the logits to the label size, one-hot target,
    nn.BCEWithLogitsLoss(weight) with the per-pixel weight, mean over B*2*L*L.  logits are the
    network's uncropped [B,2,Hp,Wp] output; labels [B,1,L,L] or [B,L,L]; weight [B,L,L]."""
    labels, weight, ws = _prep(logits, labels, weight)
    return _WeightedLossFn.apply(logits, labels, weight, ops.LOSS_BCE, global_batch, ws)


def weighted_softmax_ce(logits, labels, weight, global_batch=None):
    """Pixel-wise weight-map softmax cross-entropy (BASELINE config 5), mean over B*L*L."""
    labels, weight, ws = _prep(logits, labels, weight)
    return _WeightedLossFn.apply(logits, labels, weight, ops.LOSS_SOFTMAX_CE, global_batch, ws)


class FusedSGD(torch.optim.SGD):
    """optim.SGD(lr, momentum) (trainer.py:30) whose step() is a single fused kernel over the flat
    parameter / gradient / momentum arenas of a `network.Unet`.  The learning rate is read from
    param_groups at every step, so ReduceLROnPlateau (trainer.py:31,132) keeps working."""

    def __init__(self, unet, lr=1e-4, momentum=0.99):
        super().__init__(unet.parameters(), lr=lr, momentum=momentum)
        self._unet = unet

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("FusedSGD does not take a closure")
        eng = self._unet.engine
        st = eng.store
        if st.g_arena is None:
            raise RuntimeError("FusedSGD.step() before any backward through the B200 U-Net")
        grp = self.param_groups[0]
        if grp.get("weight_decay", 0) or grp.get("nesterov", False) or grp.get("dampening", 0):
            raise RuntimeError("FusedSGD implements plain momentum SGD only (trainer.py:30)")
        # parameters and gradients must still be the arena views
        for name, p in st.params().items():
            if p.grad is None:
                raise RuntimeError(f"{name}.grad is None: call backward() first")
            if p.grad.data_ptr() != st.grad_view(name).data_ptr():
                st.grad_view(name).copy_(p.grad)
        # parameters, gradients and momentum share one arena layout -> one elementwise launch
        ops.sgd_momentum(st.p_arena, st.g_arena, st.m_arena, float(grp["lr"]), float(grp["momentum"]), st.first_step)
        st.first_step = False
        st.dirty = True


class TrainStep:
    """The whole hot-loop body of trainer.py:54-78 -- zero_grad, forward, crop + one-hot + weighted loss, backward,
    SGD-momentum step -- issued as one explicit kernel schedule without autograd bookkeeping, and (graph=True)
    captured into a CUDA graph and replayed, so a step costs one host call regardless of its ~100 kernel launches.
    The reference trains with batch 2 (main_main.py:136), where host launch overhead, not the GPU, bounds the step.

        step = TrainStep(unet, optimizer, graph=True)
        loss = step(images, labels, weight_maps)        # device tensors; returns a 0-dim device tensor

    The graph is keyed on the input shapes and on (lr, momentum): a ReduceLROnPlateau change re-captures."""

    def __init__(self, unet, optimizer, kind="bce", graph=True, global_batch=None, graph_with_allreduce=False):
        """graph_with_allreduce: also capture when a GradAllReducer is attached (the NCCL all-reduces of the gradient
        buckets and their side stream become part of the graph).  Experimental and off: measured on 2 x B200 it is no
        faster than the eager schedule at 64 tiles per GPU (2076 vs 2113 tiles/s) and the process group then hangs at
        shutdown."""
        self.graph_with_allreduce = graph_with_allreduce
        self.unet, self.opt = unet, optimizer
        self.kind = ops.LOSS_BCE if kind == "bce" else ops.LOSS_SOFTMAX_CE
        self.use_graph = graph
        self.global_batch = global_batch
        self.graphs = {}
        self._loss_ws = None
        self.logits = None       # the step's uncropped logits [B,2,Hp,Wp] (a static buffer when replaying a graph)

    def _eager(self, images, labels, weight, out_loss, dlogits_cache):
        eng = self.unet.engine
        st = eng.store
        logits, ws = eng.forward(images, train=True)
        self.logits = logits
        b, _, hp, wp = logits.shape
        l = labels.shape[-1]
        nb = self.global_batch or b
        scale = 1.0 / ((2.0 if self.kind == ops.LOSS_BCE else 1.0) * nb * l * l)
        key = tuple(logits.shape)
        dlogits = dlogits_cache.get(key)
        if dlogits is None:
            dlogits = dlogits_cache[key] = torch.zeros_like(logits)   # cropped-away border (if any) stays zero
        if self._loss_ws is None:
            self._loss_ws = ops.loss_workspace(logits.device)
        ops.loss_fwd_bwd(self.kind, logits, labels, weight, dlogits, out_loss, self._loss_ws, scale)
        for p in self.unet.parameters():   # zero_grad(): backward() refills the arena from scratch
            p.grad = None
        eng.backward(ws, dlogits)
        eng.assign_grads()
        grp = self.opt.param_groups[0]
        ops.sgd_momentum(st.p_arena, st.g_arena, st.m_arena, float(grp["lr"]), float(grp["momentum"]), st.first_step)
        st.first_step = False
        st.dirty = True

    def __call__(self, images, labels, weight):
        if labels.dim() == 4:
            labels = labels[:, 0]
        if labels.dtype not in (torch.uint8, torch.int64):
            labels = labels.to(torch.int64)
        labels, weight = labels.contiguous(), weight.contiguous().float()
        st = self.unet.engine.store
        hooked = self.unet.engine.grad_ready_hook is not None and not self.graph_with_allreduce
        if not self.use_graph or st.p_arena is None or st.first_step or hooked:
            out = torch.zeros(1, device=images.device)
            self._eager(images, labels, weight, out, self.__dict__.setdefault("_dl", {}))
            return out[0]
        grp = self.opt.param_groups[0]
        key = (tuple(images.shape), tuple(labels.shape), labels.dtype, float(grp["lr"]), float(grp["momentum"]))
        g = self.graphs.get(key)
        if g is None:
            g = self.graphs[key] = self._capture(images, labels, weight)
        g["img"].copy_(images, non_blocking=True)
        g["lab"].copy_(labels, non_blocking=True)
        g["w"].copy_(weight, non_blocking=True)
        g["graph"].replay()
        # the replay ran pack_all (from the PRE-step masters) and then the SGD kernel, all without Python: the bf16
        # operand caches are now one step behind the fp32 masters, so the next forward outside the graph (validation,
        # predict) must re-derive them
        st.dirty = True
        self.logits = g["logits"]
        return g["loss"][0]

    def _capture(self, images, labels, weight):
        dev = images.device
        img, lab, w = images.clone(), labels.clone(), weight.clone()
        loss = torch.zeros(1, device=dev)
        dl = {}
        # warm-up on a side stream (allocates the workspace, job tables, dlogits), then capture the same schedule.
        # NOTE: the warm-up performs a real optimizer step; it is undone by restoring parameters and momentum.
        st = self.unet.engine.store
        p0, m0 = st.p_arena.clone(), st.m_arena.clone()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._eager(img, lab, w, loss, dl)
        torch.cuda.current_stream().wait_stream(s)
        st.p_arena.copy_(p0)
        st.m_arena.copy_(m0)
        st.dirty = True
        graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may query events while this thread captures
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            self._eager(img, lab, w, loss, dl)
        # the capture itself executed nothing; parameters are still p0
        # everything the captured kernels address by raw pointer must outlive the graph: the dlogits buffer in `dl`
        # was allocated by the warm-up (ordinary pool) -- dropping it here would let torch.cuda.empty_cache()
        # (trainer.py calls it every epoch) unmap memory the replays still write to
        return {"graph": graph, "img": img, "lab": lab, "w": w, "loss": loss, "logits": self.logits, "dlogits": dl}
