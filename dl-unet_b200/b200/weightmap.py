"""Loss weight maps on the device: drop-in GPU versions of the reference's
`functions.weighted_map` (functions.py:7-78) and `functions.class_balance` (functions.py:82-117).

The reference builds these per sample on the CPU with OpenCV (connectedComponents, one distanceTransform per
object, a sort over objects) and uploads the result every step (trainer.py:68-73); here the labels go to the
device once and a handful of integer kernels (csrc/weightmap.cuh) produce the whole batch's maps, bit-compatible
with the reference including its quirks: w0 = 20 (not the paper's 10), sigma^2 = 25, and w_c truncated to an
integer when the label tensor is an integer tensor (functions.py:37-40)."""
import ctypes as C

import torch

from . import lib as _lib
from .ops import STATS, _chk, _need_cuda, _ptr, _stream

_WS = {}


def _workspace(b, h, w, device):
    key = (b, h, w, str(device))
    ws = _WS.get(key)
    if ws is None:
        nbytes = _lib.load().ub_weight_map_workspace_bytes(b, h, w)
        ws = _WS[key] = (torch.empty(nbytes, dtype=torch.uint8, device=device),
                         torch.zeros(b, dtype=torch.int32, device=device))
    return ws


def _prep(gt_batch):
    _need_cuda(gt_batch)
    if gt_batch.dim() != 3:
        raise RuntimeError(f"expected a [B,H,W] label batch, got {tuple(gt_batch.shape)}")
    integer = not gt_batch.dtype.is_floating_point
    if gt_batch.dtype not in (torch.uint8, torch.int64):
        gt_batch = gt_batch.to(torch.int64)          # float / bool / int32 labels: values are 0 / 1
    return gt_batch.contiguous(), integer


def _raise_like_reference(status, what):
    bad = status.nonzero()
    if bad.numel():
        raise IndexError(f"{what}: sample {int(bad[0])} has a single class / no object "
                         "(index 1 is out of bounds, as in the reference: functions.py:40,60)")


def weighted_map(gt_batch, w0=20.0, sigma2=25.0, check=True, out=None):
    """gt_batch: CUDA [B,H,W] binary labels (int64 / uint8 / float).  Returns fp32 [B,H,W] on the same device.
    check=True reads the per-sample status back (one sync) and raises IndexError where the reference does."""
    gt, integer = _prep(gt_batch)
    b, h, w = gt.shape
    if out is None:
        out = torch.empty(b, h, w, dtype=torch.float32, device=gt.device)
    ws, status = _workspace(b, h, w, gt.device)
    _chk(_lib.load().ub_weight_map(_ptr(gt), gt.element_size(), b, h, w, C.c_double(w0), C.c_double(sigma2),
                                   int(integer), _ptr(out), _ptr(ws), _ptr(status), _stream()), "weight_map")
    STATS.launches += 5   # count, init, merge, flatten, columns, rows: six kernels behind one call
    if check:
        _raise_like_reference(status, "weighted_map")
    return out


def class_balance(gt_batch, check=True, out=None):
    gt, _ = _prep(gt_batch)
    b, h, w = gt.shape
    if out is None:
        out = torch.empty(b, h, w, dtype=torch.float32, device=gt.device)
    ws, status = _workspace(b, h, w, gt.device)
    _chk(_lib.load().ub_class_balance(_ptr(gt), gt.element_size(), b, h, w, _ptr(out), _ptr(ws), _ptr(status), _stream()),
         "class_balance")
    STATS.launches += 1
    if check:
        _raise_like_reference(status, "class_balance")
    return out


def mask_counts(pred, label):
    """Device-side reduction behind the reference's IoU / Pixel_error (functions.py:172-213): returns an int64 CUDA
    tensor [|p & l|, |p | l|, sum |p - l|] without copying the masks to the host."""
    _need_cuda(pred, label)
    if pred.numel() != label.numel():
        raise RuntimeError(f"pred {tuple(pred.shape)} and label {tuple(label.shape)} differ in size")
    conv = lambda t: t.contiguous() if t.dtype in (torch.uint8, torch.int64) else t.to(torch.int64).contiguous()
    pred, label = conv(pred), conv(label)
    counts = torch.zeros(3, dtype=torch.int64, device=pred.device)
    _chk(_lib.load().ub_mask_metrics(_ptr(pred), pred.element_size(), _ptr(label), label.element_size(), pred.numel(),
                                     _ptr(counts), _stream()), "mask_metrics")
    return counts


def evaluation_metrics(pred, label):
    """functions.py:150-168 on CUDA masks: 2x1 float64 array [IoU; pixel error] (one 24-byte read-back)."""
    import numpy as np
    c = mask_counts(pred, label).cpu().numpy().astype(np.float64)
    out = np.empty([2, 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        out[0] = c[0] / c[1]                       # nan for two empty masks, as numpy's 0/0 in the reference
    out[1] = c[2] / pred.numel()
    return out
