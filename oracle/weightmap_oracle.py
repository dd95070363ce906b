"""CPU oracle for the loss weight maps -- TEST INFRASTRUCTURE ONLY (see unet_oracle.py's header).

A numpy / scipy restatement of the reference's `weighted_map` (functions.py:7-78) and `class_balance`
(functions.py:82-117) that does not use OpenCV: components via scipy.ndimage.label (4-connectivity), exact
squared Euclidean distances via the feature transform, then the reference's arithmetic step by step (float32
distances as cv.distanceTransform returns them, float64 formula, one rounding to float32).

Parity pin: oracle/validate_against_reference.py runs the UNMODIFIED reference functions on seeded labels and
stores their outputs in tests/golden/weightmap_*.npz; tests/test_oracle_golden.py checks this file against them
bit for bit."""
import numpy as np
from scipy import ndimage


def _w_c(gt):
    """functions.py:34-40: counts[1] / counts[class], stored in the label's own dtype."""
    uval, counts = np.unique(gt.astype(np.float32), return_counts=True)
    w_c = np.empty(gt.shape, dtype=gt.dtype)
    for pos in range(len(uval)):
        ratio = np.float32(counts[1]) / np.float32(counts[pos])          # IndexError on a single class, as upstream
        w_c[gt == uval[pos]] = ratio if gt.dtype.kind == "f" else np.trunc(ratio)
    return w_c


def weighted_map(gt_batch, w0=20, sig2=25):
    """gt_batch: numpy [B,H,W] of {0,1} (integer or float dtype) -> float32 [B,H,W]."""
    out = np.empty(gt_batch.shape, dtype=np.float32)
    cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])                  # connectivity=4 (functions.py:46)
    for b in range(gt_batch.shape[0]):
        gt = gt_batch[b]
        w_c = _w_c(gt)
        objects, n = ndimage.label(gt.astype(np.uint8), structure=cross)
        maps = np.zeros((n,) + gt.shape)
        yy, xx = np.indices(gt.shape)
        for i in range(n):                                               # functions.py:54-56, one transform per object
            iy, ix = ndimage.distance_transform_edt(objects != i + 1, return_distances=False, return_indices=True)
            dsq = (yy - iy) ** 2 + (xx - ix) ** 2                         # exact integers
            maps[i] = np.sqrt(dsq.astype(np.float32))                    # cv.DIST_L2 precise: float32 sqrt of the exact d^2
        maps = np.sort(maps, 0)
        d1 = maps[0]                                                     # IndexError without objects, as upstream
        d2 = maps[1] if n > 1 else 0
        w_d = w0 * np.exp(-np.square(d1 + d2) / (2 * sig2)) * (-1 * (gt.astype(np.float64) - 1))
        out[b] = (w_c.astype(np.float64) + w_d).astype(np.float32)
    return out


def class_balance(gt_batch):
    out = np.empty(gt_batch.shape, dtype=np.float32)
    for b in range(gt_batch.shape[0]):
        gt = gt_batch[b]
        uval, counts = np.unique(gt, return_counts=True)
        w = np.ones(gt.shape, dtype=np.float32)
        for pos in range(len(uval)):
            w[gt == uval[pos]] = np.float32(counts[1]) / np.float32(counts[pos])
        out[b] = w
    return out


def synthetic_labels(batch, size, seed, kind="blobs"):
    """Seeded binary label batches for the parity tests: 'blobs' (touching / separate ellipses), 'noise'
    (salt-and-pepper: hundreds of tiny objects), 'single' (exactly one object)."""
    rng = np.random.RandomState(seed)
    lab = np.zeros((batch, size, size), dtype=np.int64)
    yy, xx = np.indices((size, size))
    for b in range(batch):
        if kind == "noise":
            lab[b] = (rng.rand(size, size) < 0.15).astype(np.int64)
        elif kind == "single":
            cy, cx = size * (0.3 + 0.4 * rng.rand(2))
            lab[b] = (((yy - cy) / (size / 5)) ** 2 + ((xx - cx) / (size / 7)) ** 2 <= 1).astype(np.int64)
        else:
            for _ in range(max(3, size // 12)):
                cy, cx = rng.rand(2) * size
                ry, rx = 2 + rng.rand(2) * size / 8
                lab[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = 1
    return lab
