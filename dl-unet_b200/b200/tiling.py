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
