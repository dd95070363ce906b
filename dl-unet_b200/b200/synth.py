"""Synthetic workloads of the shapes BASELINE.json names (there is no network for the real
datasets): uniform [0,1] tiles, blob labels, and a border-emphasis weight map
w = 1 + w0 * exp(-d^2 / (2 sigma^2)) on background pixels (w0 = 10, sigma = 5, BASELINE config 5)."""
import torch

from .geometry import Geometry


def synthetic_batch(batch, size=572, seed=0, w0=10.0, sigma=5.0):
    """-> images fp32 [B,1,S,S], labels int64 [B,1,L,L], weight fp32 [B,L,L] (CPU tensors)."""
    out = Geometry(size).out
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(batch, 1, size, size, generator=g)
    yy, xx = torch.meshgrid(torch.arange(out), torch.arange(out), indexing="ij")
    labels = torch.zeros(batch, 1, out, out, dtype=torch.int64)
    dist = torch.full((batch, out, out), 1e9)
    for b in range(batch):
        for _ in range(max(2, out // 24)):
            cy, cx = (torch.rand(2, generator=g) * out).tolist()
            ry, rx = (4 + torch.rand(2, generator=g) * max(4.0, out / 6)).tolist()
            d = torch.sqrt(((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2)
            labels[b, 0][d <= 1] = 1
            dist[b] = torch.minimum(dist[b], (d - 1).clamp_min(0) * min(ry, rx))
    weight = 1.0 + w0 * torch.exp(-(dist ** 2) / (2 * sigma ** 2)) * (1 - labels[:, 0]).float()
    return images, labels, weight.float()
