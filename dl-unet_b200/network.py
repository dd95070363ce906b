"""Drop-in for the reference's `network.py`: same `Unet` nn.Module surface -- no constructor
arguments, attributes conv11c ... finalconv, the same 46-entry fp32 state_dict, the same
initialisation (including its 2/sqrt(N) quirk), `crop_and_concat(A, B)` and `forward(t)` --
but forward/backward run on hand-written sm_100a kernels through the C ABI in
include/unet_b200.h instead of ATen.  There is no CPU implementation: a CPU input raises.

Reference: network.py:8-192.
"""
import torch
import torch.nn as nn

from b200 import engine as _engine
from b200.geometry import LAYERS, crop_offset


def _init_std(name, cin):
    # network.py:70-105.  `2 / (N) ** 0.5` is 2/sqrt(N) by operator precedence; conv11c uses sqrt(2).
    if name == "conv11c":
        return 2 ** 0.5
    if name in ("conv41e", "conv31e", "conv21e", "conv11e"):
        return 2 / (cin // 2 * 3.0 ** 2 + cin // 2 * 2.0 ** 2) ** 0.5
    return 2 / (cin * 3.0 ** 2) ** 0.5


class Unet(nn.Module):
    """U-Net 2D (Ronneberger et al. 2015), valid convolutions, 64-1024 channels."""

    def __init__(self):
        super().__init__()
        # Layer objects exist for their parameters / state_dict keys only (network.py:23-58);
        # they are created in the reference's order so a given torch seed yields the same weights.
        for (name, kind, cin, cout) in LAYERS:
            if kind == "up":
                layer = nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2)
            elif kind == "head":
                layer = nn.Conv2d(cin, cout, kernel_size=1)
            else:
                layer = nn.Conv2d(cin, cout, kernel_size=3)
            setattr(self, name, layer)
        for (name, kind, cin, cout) in LAYERS:  # network.py:70-105
            layer = getattr(self, name)
            layer.weight = nn.Parameter(torch.empty_like(layer.weight).normal_(mean=0, std=_init_std(name, cin)))
        object.__setattr__(self, "_b200", _engine.Engine(self))

    @property
    def engine(self):
        return self._b200

    def crop_and_concat(self, A, B):
        """network.py:108-127 for NCHW tensors, on the GPU kernels' terms: c = int((HA-HB)*0.5);
        c < 0 zero-pads A, c > 0 centre-crops it; channel order [A, B].  Inside forward() this is
        never materialised (producers write into the concat buffer); this method exists for API
        parity and performs the same placement with a device copy."""
        c = crop_offset(A.shape[2], B.shape[2])
        ha = A.shape[2] - 2 * c
        if ha != B.shape[2] or A.shape[3] - 2 * c != B.shape[3]:
            raise RuntimeError(
                f"Sizes of tensors must match except in dimension 1. Expected size {ha} but got size {B.shape[2]} "
                "for tensor number 1 in the list.")
        out = torch.zeros(A.shape[0], A.shape[1] + B.shape[1], B.shape[2], B.shape[3], dtype=A.dtype, device=A.device)
        if c >= 0:
            out[:, :A.shape[1]] = A[:, :, c:c + ha, c:c + ha]
        else:
            out[:, :A.shape[1], -c:-c + A.shape[2], -c:-c + A.shape[3]] = A
        out[:, A.shape[1]:] = B
        return out

    def forward(self, t):
        """t: fp32 [B,1,S,S] on a CUDA device -> fp32 logits [B,2,S-184,S-184] (network.py:129-192)."""
        return _engine.unet_forward(self._b200, t)

    @torch.no_grad()
    def predict(self, t, label_size=None):
        """Inference fast path (tester.py:27-30): forward + centre-crop + argmax fused into the head
        kernel; returns an int64 mask [B, L, L]."""
        if not t.is_cuda:
            raise RuntimeError("the B200 U-Net path has no CPU implementation")
        eng = self._b200
        from b200.geometry import Geometry
        out = Geometry(t.shape[-1]).out
        am = torch.empty(t.shape[0], out, out, dtype=torch.uint8, device=t.device)
        eng.forward(t, train=False, argmax_out=am, want_logits=False)
        if label_size is not None and label_size != out:
            pad = int((out - label_size) / 2)
            am = am[:, pad:label_size + pad, pad:label_size + pad]
        return am.long()
