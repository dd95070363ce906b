"""Drop-in for the reference's `functions.py` (same names, arguments, return types, and -- on
purpose -- the same numerical quirks).  These are the CPU producers and
consumers either side of the GPU hot path: weight maps feed the fused loss, the metrics read the
masks back.  They are not kernels; GPU versions are the "next" rows f1 / f3 of SURVEY.md section 8.

Reference: functions.py:7-213.
"""
import numpy as np
import torch

from b200.geometry import input_size_compute as _input_size_compute

try:  # OpenCV does the connected components / distance transform in the reference
    import cv2 as _cv
except Exception:  # pragma: no cover
    _cv = None


def class_balance(gt_batch):
    """functions.py:82-117: per-sample class-frequency weights, w[gt == v] = count(1) / count(v)
    with `counts[1]` the SECOND unique value (IndexError on a single-class label, as upstream).
    A CUDA label batch is processed on the device (b200.weightmap)."""
    if gt_batch.is_cuda:
        from b200 import weightmap as _wm
        return _wm.class_balance(gt_batch)
    w_batch = torch.empty(gt_batch.shape, dtype=torch.float32)
    for b in range(gt_batch.shape[0]):
        gt = gt_batch[b]
        uval, counts = torch.unique(gt, return_counts=True)
        w = torch.ones(gt.shape, dtype=torch.float32)
        for pos in range(len(uval)):
            w[gt == uval[pos]] = counts[1].float() / counts[pos].float()
        w_batch[b] = w
    return w_batch


def weighted_map(gt_batch):
    """functions.py:7-78: Ronneberger's w = w_c + w0 exp(-(d1+d2)^2 / (2 sigma^2)) on background
    pixels, with the file's constants (w0 = 20, sigma^2 = 25), 4-connected components and exact L2
    distance transforms.  w_c is held in the label's dtype, so an int64 label truncates the class
    ratio exactly like the reference does (functions.py:37-40)."""
    if gt_batch.is_cuda:   # device path: csrc/weightmap.cuh, same values (and quirks) without OpenCV or the host
        from b200 import weightmap as _wm
        return _wm.weighted_map(gt_batch)
    if _cv is None:
        raise RuntimeError("weighted_map needs OpenCV (cv2), as the reference does")
    w0, sig2 = 20, 25
    w_batch = torch.empty(gt_batch.shape, dtype=torch.float32)
    for b in range(gt_batch.shape[0]):
        gt = gt_batch[b]
        uval, counts = torch.unique(gt.float(), return_counts=True)
        w_c = torch.empty_like(gt)
        for pos in range(len(uval)):
            w_c[gt == uval[pos]] = counts[1].float() / counts[pos].float()
        img = gt.reshape(gt.shape[0], -1).numpy()
        n_obj, objects = _cv.connectedComponents(img.astype(np.uint8), connectivity=4)
        maps = np.zeros([n_obj - 1, img.shape[0], img.shape[1]])
        for i in range(n_obj - 1):
            sep = (objects == i + 1).astype(np.float64)
            # uint8(obj) - 1 wraps to 255 outside the object and is 0 inside it: distance to the object
            maps[i] = _cv.distanceTransform(sep.astype(np.uint8) - 1, _cv.DIST_L2, maskSize=0)
        maps = np.sort(maps, 0)
        d1 = maps[0]
        d2 = maps[1] if maps.shape[0] > 1 else 0
        w_d = np.multiply(w0 * np.exp(-np.square(d1 + d2) / (2 * sig2)), -1 * (gt - 1))
        w_batch[b] = (w_c + w_d).clone().detach()
    return w_batch


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
