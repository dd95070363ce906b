"""Drop-in for the reference's `functions.py` (same names, arguments, return types, and -- on
purpose -- the same numerical quirks).  Producers and consumers either side of
the network: the weight maps that feed the fused loss run as device kernels (SURVEY.md 8 row f1,
no host implementation here), the metrics reduce CUDA masks on the device (row f3) and keep the
reference's two-line numpy formulas for CPU masks.

Reference: functions.py:7-213.
"""
import numpy as np
import torch

from b200.geometry import input_size_compute as _input_size_compute

from b200 import weightmap as _wm


def _on_device(gt_batch, fn):
    """The weight maps are device kernels (csrc/weightmap.cuh); there is no host implementation.  A CPU label
    batch -- what the reference's loaders yield (trainer.py:69,72) -- is uploaded to the current CUDA device, processed
    there and returned on the CPU, so the call keeps the reference's signature and placement.
    Size limit of the device kernels: H < 2048, W <= 2048 and H*W < 2^21 - 1 (RuntimeError otherwise)."""
    if gt_batch.is_cuda:
        return fn(gt_batch)
    if not torch.cuda.is_available():
        raise RuntimeError("functions.weighted_map / class_balance run on the B200 kernels: no CUDA device, no CPU path")
    return fn(gt_batch.cuda()).cpu()


def class_balance(gt_batch):
    """functions.py:82-117: per-sample class-frequency weights, w[gt == v] = count(1) / count(v) with `counts[1]` the
    SECOND unique value (IndexError on a single-class label, as upstream).  Device kernels: b200.weightmap."""
    return _on_device(gt_batch, _wm.class_balance)


def weighted_map(gt_batch):
    """functions.py:7-78: Ronneberger's w = w_c + w0 exp(-(d1+d2)^2 / (2 sigma^2)) on background pixels, with the
    file's constants (w0 = 20, sigma^2 = 25), 4-connected components and exact L2 distances to the two nearest
    objects; w_c is truncated to an integer for integer labels, like the reference (functions.py:37-40).
    Device kernels: b200.weightmap."""
    return _on_device(gt_batch, _wm.weighted_map)


def input_size_compute(image):
    """functions.py:121-146: (original_size, input_size, output_size) of the smallest valid U-Net
    input whose output covers the image."""
    return _input_size_compute(image.shape[-1])


def Pixel_error(pred, label):
    """functions.py:172-189."""
    p, l = pred.cpu().numpy(), label.cpu().numpy()
    return np.sum(abs(p - l)) / p.size


def IoU(pred, label):
    """functions.py:193-213."""
    p, l = pred.cpu().numpy(), label.cpu().numpy()
    return np.sum(np.logical_and(p, l)) / np.sum(np.logical_or(p, l))


def evaluation_metrics(pred, label):
    """functions.py:150-168: 2x1 array [IoU; pixel error].  CUDA masks are reduced on the device."""
    if pred.is_cuda and label.is_cuda:
        from b200 import weightmap as _wm
        return _wm.evaluation_metrics(pred, label)
    out = np.empty([2, 1])
    out[0] = IoU(pred, label)
    out[1] = Pixel_error(pred, label)
    return out
