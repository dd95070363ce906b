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
