"""Device-side training-sample pipeline: the reference's `ImageDataset.__getitem__`
(data.py:93-137) for a whole batch at once -- weighted random crop, reflect pad, 30-degree-step spline rotation,
elastic deformation, label threshold, min-max normalisation -- on images that live on the GPU.

The reference does this per sample in the DataLoader's main thread (num_workers = 0; about one second per 388-pixel
crop with scipy); `DeviceAugmenter` keeps the dataset on the device and produces a batch with one library call
(csrc/augment.cuh).  The random integers (crop position, rotation) are drawn on the host exactly as data.py does,
from numpy's global RNG; the two displacement noise fields come from torch's device RNG unless given."""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .geometry import input_size_compute
from .ops import STATS, _chk, _need_cuda, _ptr, _stream


class CropSampler:
    """Host side of the pipeline (no device needed): the crop anchors and their weighted distribution per dataset
    frame (data.py:66-82) and the random integers of one __getitem__ call (data.py:97-104, 115)."""

    def __init__(self, targets, crop=388, skip=10):
        """targets: uint8 array [N, H, W] of 0 / 255 masks (data.py:63's gt_bin)."""
        targets = np.asarray(targets)
        self.n, self.h, self.w = targets.shape
        self.crop, self.skip = int(crop), int(skip)
        # data.py:66-82: crop anchors every `skip` pixels, weighted towards windows that are ~50 % foreground
        self.pairs = [(i, j) for i in range(0, self.h - self.crop, self.skip) for j in range(0, self.w - self.crop, self.skip)]
        self.distribution = [self._crop_distribution(t) for t in targets]

    def _crop_distribution(self, gt_bin):
        from scipy.stats import norm
        # window means through an integral image (the reference takes np.mean of every window)
        ii = np.pad(gt_bin.astype(np.float64), ((1, 0), (1, 0))).cumsum(0).cumsum(1)
        c = self.crop
        p = []
        for (i, j) in self.pairs:
            x = (ii[i + c, j + c] - ii[i, j + c] - ii[i + c, j] + ii[i, j]) / (c * c) / 255
            p.append(0 if (x < 0.1 or x > 0.9) else 10 * norm.pdf(x, loc=0.5, scale=0.05))
        p = np.asarray(p, dtype=np.float64)
        return np.ones(len(p)) / len(p) if p.sum() == 0 else p / p.sum()

    def draw(self, indices):
        """The random integers of data.py:97-104,115 for each dataset index, from numpy's global RNG in the
        reference's order.  -> int32 [B,4] = (image index, x, y, rotation degrees)."""
        out = np.empty((len(indices), 4), dtype=np.int32)
        for k, idx in enumerate(indices):
            crop_id = np.random.choice(range(len(self.pairs)), 1, p=self.distribution[idx])[0]
            x, y = self.pairs[crop_id]
            x += np.random.randint(-self.skip / 2, (self.skip / 2) + 1)
            y += np.random.randint(-self.skip / 2, (self.skip / 2) + 1)
            x = min(max(0, x), self.h - self.crop)
            y = min(max(0, y), self.w - self.crop)
            out[k] = (idx, x, y, np.random.choice(np.arange(0, 360, 30)))
        return out


class DeviceAugmenter(CropSampler):
    def __init__(self, images, targets, crop=388, alpha=3, sigma=10, skip=10, device="cuda"):
        """images / targets: uint8 arrays or tensors [N, H, W] (targets 0 / 255, as data.py:63 produces them)."""
        images, targets = torch.as_tensor(np.asarray(images)), torch.as_tensor(np.asarray(targets))
        if images.dtype != torch.uint8 or targets.dtype != torch.uint8 or images.shape != targets.shape or images.dim() != 3:
            raise RuntimeError("DeviceAugmenter expects uint8 [N,H,W] image and target stacks of equal shape")
        self.images, self.targets = images.to(device).contiguous(), targets.to(device).contiguous()
        _need_cuda(self.images)
        CropSampler.__init__(self, targets.numpy(), crop=crop, skip=skip)
        self.alpha, self.sigma = float(alpha), float(sigma)
        _, self.input_size, _ = input_size_compute(self.crop)
        self._ws = {}

    def __call__(self, indices=None, draws=None, noise=None):
        """-> (images fp32 [B,1,S,S], labels int64 [B,1,crop,crop]) on the device.  `draws`: int [B,4] from draw()
        (default: drawn now); `noise`: fp64 [B,2,S,S] uniform fields (default: torch's device RNG)."""
        if draws is None:
            draws = self.draw(indices)
        draws = np.ascontiguousarray(draws, dtype=np.int32)
        b, S = draws.shape[0], self.input_size
        dev = self.images.device
        if (draws[:, 3] % 30).any() or (draws[:, 0] < 0).any() or (draws[:, 0] >= self.n).any() or (draws[:, 1:3] < 0).any() \
                or (draws[:, 1] > self.h - self.crop).any() or (draws[:, 2] > self.w - self.crop).any():
            raise RuntimeError("augmentation draws out of range (image index, crop window or rotation not a multiple of 30)")
        d_draws = torch.from_numpy(draws).to(dev)
        if noise is None:
            noise = torch.rand(b, 2, S, S, dtype=torch.float64, device=dev)
        else:
            noise = torch.as_tensor(noise, dtype=torch.float64).to(dev).contiguous()
            if tuple(noise.shape) != (b, 2, S, S):
                raise RuntimeError(f"noise must be [B,2,{S},{S}], got {tuple(noise.shape)}")
        lib = _lib.load()
        ws = self._ws.get(b)
        if ws is None:
            ws = self._ws[b] = torch.empty(lib.ub_augment_workspace_bytes(b, self.crop, S, C.c_double(self.sigma)),
                                           dtype=torch.uint8, device=dev)
        img = torch.empty(b, 1, S, S, dtype=torch.float32, device=dev)
        lab = torch.empty(b, 1, self.crop, self.crop, dtype=torch.int64, device=dev)
        _chk(lib.ub_augment_batch(_ptr(self.images), _ptr(self.targets), self.n, self.h, self.w, _ptr(d_draws), _ptr(noise),
                                  b, self.crop, S, C.c_double(self.alpha), C.c_double(self.sigma), _ptr(img), _ptr(lab),
                                  _ptr(ws), _stream()), "augment_batch")
        STATS.launches += 9   # ten kernels behind one call
        return img, lab
