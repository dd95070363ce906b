"""Overlap-tile planning for whole-image inference (Ronneberger's strategy; the reference only has the
whole-image mirror padding of data.py:249-277).  Pure index arithmetic + torch tensor plumbing; the
per-tile forward is the kernel path.  NOTE: the reference network zero-pads its
skip connections, so it is not tile-equivariant -- a stitched mosaic equals the reference run on the
same TILES, not the reference run on the whole image."""
import torch
import torch.nn.functional as F

from .geometry import Geometry


def plan(height, width, tile_in=572):
    """-> (tile_out, halo, ny, nx, padded_h, padded_w): tiles of tile_in input pixels produce tile_out
    output pixels; the image is reflect-padded by `halo` and bottom/right-filled up to ny*tile_out."""
    tile_out = Geometry(tile_in).out
    halo = (tile_in - tile_out) // 2
    ny = -(-height // tile_out)
    nx = -(-width // tile_out)
    return tile_out, halo, ny, nx, ny * tile_out + 2 * halo, nx * tile_out + 2 * halo


def cut_tiles(image, tile_in=572):
    """image [H, W] (any device) -> tiles [ny*nx, 1, tile_in, tile_in] (row-major tile order)."""
    h, w = image.shape
    tile_out, halo, ny, nx, ph, pw = plan(h, w, tile_in)
    img = image[None, None]
    img = F.pad(img, (halo, halo, halo, halo), mode="reflect")                     # mirror, no edge repeat
    img = F.pad(img, (0, pw - img.shape[-1], 0, ph - img.shape[-2]), mode="replicate")
    t = img[0, 0].unfold(0, tile_in, tile_out).unfold(1, tile_in, tile_out)        # [ny, nx, tile_in, tile_in]
    return t.reshape(ny * nx, 1, tile_in, tile_in).contiguous()


def stitch(masks, height, width, tile_in=572):
    """masks [ny*nx, tile_out, tile_out] -> [height, width]."""
    tile_out, halo, ny, nx, _, _ = plan(height, width, tile_in)
    m = masks.reshape(ny, nx, tile_out, tile_out).permute(0, 2, 1, 3).reshape(ny * tile_out, nx * tile_out)
    return m[:height, :width]


def mirror_transform(image, dtype=torch.float32):
    """data.py:249-312 (`mirror_transform` / `mirror_transform_tensor`) on the device: a CUDA image [..., n, n] is
    mirrored outwards up to the network input size of `input_size_compute` -> [B, 1, input_size, input_size].
    Same index rule as the reference's slice assignments (edge not repeated on the low side, repeated on the high
    side); like them it needs an even size difference.  The reference returns float64; the network input is fp32."""
    from . import lib as _lib
    from .geometry import input_size_compute
    from .ops import _chk, _need_cuda, _ptr, _stream
    _need_cuda(image)
    n = image.shape[-1]
    if image.shape[-2] != n:
        raise RuntimeError(f"expected square images, got {tuple(image.shape)}")
    _, size, _ = input_size_compute(n)
    if (size - n) % 2:
        raise ValueError(f"could not broadcast: input size {size} minus image size {n} is odd (data.py:270-277)")
    src = image.reshape(-1, n, n).contiguous().float()
    out = torch.empty(src.shape[0], 1, size, size, dtype=torch.float32, device=image.device)
    _chk(_lib.load().ub_mirror_pad(_ptr(src), _ptr(out), src.shape[0], n, size, _stream()), "mirror_pad")
    return out.to(dtype)


def segment_image(unet, image, tile_in=572, batch=16, ranks=(0, 1)):
    """Overlap-tile inference of one whole image [H, W] (CUDA, fp32): reflect-pad, cut into tiles, run forward +
    centre-crop + argmax (tester.py:27-30) on this rank's share of the tiles (rank r of n takes tiles r, r+n, ...;
    independent units, no collective) and return (tile_indices, uint8 masks [k, tile_out, tile_out]).  With
    ranks=(0, 1) the masks are stitched and the [H, W] uint8 mask is returned instead."""
    rank, world = ranks
    h, w = image.shape
    tile_out, _, ny, nx, _, _ = plan(h, w, tile_in)
    mine = list(range(rank, ny * nx, world))
    tiles = cut_tiles(image, tile_in)[mine]
    out = torch.empty(len(mine), tile_out, tile_out, dtype=torch.uint8, device=image.device)
    for b0 in range(0, len(mine), batch):
        t = tiles[b0:b0 + batch]
        unet.engine.forward(t, train=False, argmax_out=out[b0:b0 + t.shape[0]], want_logits=False)
    if world == 1:
        return stitch(out, h, w, tile_in)
    return mine, out
