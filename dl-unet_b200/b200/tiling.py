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


def _src_index(i, n, halo):
    """Source row / column of window coordinate i (image frame): reflect without edge repeat inside the halo,
    repeat the last reflected line beyond it -- the index rule of cut_tiles' two F.pad calls and of ub_cut_tiles."""
    i = min(i, n - 1 + halo)
    i = -i if i < 0 else i
    return 2 * (n - 1) - i if i > n - 1 else i


def needed_rows(height, tile_ids, nx, tile_in, tile_out):
    """(row0, rows): the band of image rows the windows of `tile_ids` read (so a rank uploads only its band)."""
    halo = (tile_in - tile_out) // 2
    lo, hi = height, -1
    for ty in sorted({t // nx for t in tile_ids}):
        for i in (ty * tile_out - halo, ty * tile_out - halo + tile_in - 1):
            r = _src_index(i, height, halo)
            lo, hi = min(lo, r), max(hi, r)
        if ty * tile_out - halo < 0:
            lo = 0                                            # the reflection passes through row 0
        if ty * tile_out - halo + tile_in - 1 > height - 1:
            hi = height - 1                                   # ... or through the last row
    return lo, hi - lo + 1


def cut_tiles_device(band, height, width, row0, tile_ids_dev, n_tiles, nx, tile_in, tile_out, out):
    """ub_cut_tiles: gather the input windows of the listed tiles straight from (a band of) the image."""
    from . import lib as _lib
    from .ops import _chk, _need_cuda, _ptr, _stream
    _need_cuda(band, tile_ids_dev, out)
    _chk(_lib.load().ub_cut_tiles(_ptr(band), height, width, row0, band.shape[0], _ptr(tile_ids_dev), n_tiles, nx, tile_in,
                                  tile_out, _ptr(out), _stream()), "cut_tiles")
    return out


def stitch_device(masks, tile_ids_dev, n_tiles, nx, tile_out, out):
    """ub_stitch_tiles: scatter uint8 mask tiles into the [H, W] mask (clipped at the image border)."""
    from . import lib as _lib
    from .ops import _chk, _need_cuda, _ptr, _stream
    _need_cuda(masks, tile_ids_dev, out)
    _chk(_lib.load().ub_stitch_tiles(_ptr(masks), _ptr(tile_ids_dev), n_tiles, nx, tile_out, _ptr(out), out.shape[0],
                                     out.shape[1], _stream()), "stitch_tiles")
    return out


class TileSegmenter:
    """Overlap-tile inference of H x W images, tiles sharded over the ranks of a process group with NO collective on
    the data path (SURVEY.md 8 e-infer): rank r owns a contiguous range of tiles, uploads only the band of image rows
    those tiles read, gathers their input windows with one kernel (no padded image, no tile stack of the whole image),
    runs forward + centre-crop + argmax (tester.py:27-30, argmax fused into the head epilogue) in batches, and the
    uint8 mask tiles are gathered on rank 0 and scattered into the stitched mask by one kernel.
    `segment_many` pipelines a stream of host images: upload of image i+1 and download of mask i-1 overlap the
    forward of image i (two slots, three streams)."""

    def __init__(self, unet, height, width, tile_in=572, batch=16, rank=0, world=1, group=None, device=None):
        from .dist import partition
        self.unet, self.H, self.W, self.tile_in, self.batch = unet, height, width, tile_in, batch
        self.rank, self.world, self.group = rank, world, group
        self.dev = device or next(unet.parameters()).device
        self.tile_out, self.halo, self.ny, self.nx, _, _ = plan(height, width, tile_in)
        n = self.ny * self.nx
        self.shares = [list(range(*partition(n, r, world))) for r in range(world)]
        self.mine = self.shares[rank]
        self.cap = max(len(s) for s in self.shares)
        self.row0, self.rows = needed_rows(height, self.mine, self.nx, tile_in, self.tile_out) if self.mine else (0, 1)
        i32 = dict(dtype=torch.int32, device=self.dev)
        self.ids = [torch.tensor(s if s else [0], **i32) for s in self.shares]
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.band = [torch.empty(self.rows, width, dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.tiles = torch.empty(min(batch, max(1, len(self.mine))), 1, tile_in, tile_in, dtype=torch.float32, device=self.dev)
        self.masks = torch.empty(self.cap, self.tile_out, self.tile_out, **u8)
        if rank == 0:
            self.parts = [torch.empty_like(self.masks) for _ in range(world)] if world > 1 else None
            self.full = [torch.empty(height, width, **u8) for _ in range(2)]
            self.host = [torch.empty(height, width, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.s_in = torch.cuda.Stream(device=self.dev)
        self.s_out = torch.cuda.Stream(device=self.dev)

    def band_of(self, image):
        """The rows of a full [H, W] image (host or device) this rank needs."""
        return image[self.row0:self.row0 + self.rows]

    def segment_band(self, band):
        """band: device fp32 [rows, W] = image rows [row0, row0 + rows).  Fills and returns this rank's uint8 mask
        tiles [cap, tile_out, tile_out] (the first len(mine) are valid)."""
        k = len(self.mine)
        for b0 in range(0, k, self.batch):
            nb = min(self.batch, k - b0)
            t = self.tiles[:nb]
            cut_tiles_device(band, self.H, self.W, self.row0, self.ids[self.rank][b0:b0 + nb], nb, self.nx, self.tile_in,
                             self.tile_out, t)
            self.unet.engine.forward(t, train=False, argmax_out=self.masks[b0:b0 + nb], want_logits=False)
        return self.masks

    def collect(self, slot=0):
        """Gather every rank's mask tiles on rank 0 and stitch them into full[slot] (device uint8 [H, W])."""
        import torch.distributed as dist
        if self.world > 1:
            dist.gather(self.masks, self.parts if self.rank == 0 else None, dst=0, group=self.group)
        if self.rank != 0:
            return None
        parts = self.parts if self.world > 1 else [self.masks]
        for r, share in enumerate(self.shares):
            if share:
                stitch_device(parts[r], self.ids[r], len(share), self.nx, self.tile_out, self.full[slot])
        return self.full[slot]

    def __call__(self, image):
        """One image (host or device, [H, W] fp32) -> stitched uint8 mask on rank 0 (device), None elsewhere."""
        band = self.band[0]
        band.copy_(self.band_of(image), non_blocking=True)
        self.segment_band(band)
        return self.collect(0)

    def segment_many(self, host_images):
        """Generator over pinned host images [H, W] fp32: yields (index, host uint8 mask [H, W]) on rank 0 and
        (index, None) elsewhere.  A yielded mask lives in a pinned staging buffer that is reused two images later."""
        cur = torch.cuda.current_stream(self.dev)
        uploaded, computed, downloaded = [None, None], [None, None], [None, None]
        images = iter(host_images)

        def upload(i, img):
            s = i % 2
            with torch.cuda.stream(self.s_in):
                if computed[s] is not None:
                    self.s_in.wait_event(computed[s])             # the forward that read this slot has finished
                self.band[s].copy_(self.band_of(img), non_blocking=True)
                uploaded[s] = torch.cuda.Event()
                uploaded[s].record(self.s_in)

        nxt = next(images, None)
        i = 0
        if nxt is not None:
            upload(0, nxt)
        pending = None
        while nxt is not None:
            s = i % 2
            cur.wait_event(uploaded[s])
            following = next(images, None)
            if following is not None:
                upload(i + 1, following)
            self.segment_band(self.band[s])
            if self.rank == 0 and downloaded[s] is not None:
                cur.wait_event(downloaded[s])                     # full[s] was copied out
            full = self.collect(s)
            computed[s] = torch.cuda.Event()
            computed[s].record(cur)
            if self.rank == 0:
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(computed[s])
                    self.host[s].copy_(full, non_blocking=True)
                    downloaded[s] = torch.cuda.Event()
                    downloaded[s].record(self.s_out)
            if pending is not None:
                pi, ps = pending
                if self.rank == 0:
                    downloaded[ps].synchronize()
                yield pi, (self.host[ps] if self.rank == 0 else None)
            pending = (i, s)
            nxt = following
            i += 1
        if pending is not None:
            pi, ps = pending
            if self.rank == 0:
                downloaded[ps].synchronize()
            else:
                computed[ps].synchronize()
            yield pi, (self.host[ps] if self.rank == 0 else None)


def segment_image(unet, image, tile_in=572, batch=16, ranks=(0, 1)):
    """Overlap-tile inference of one whole image [H, W] (CUDA, fp32): this rank's share of the tiles (a contiguous
    range; independent units, no collective) goes through gather -> forward + centre-crop + argmax (tester.py:27-30).
    Returns (tile_indices, uint8 masks [k, tile_out, tile_out]); with ranks=(0, 1) the [H, W] uint8 mask instead."""
    rank, world = ranks
    h, w = image.shape
    seg = TileSegmenter(unet, h, w, tile_in, batch, rank, world, device=image.device)
    if world == 1:
        return seg(image.float()).clone()
    band = seg.band[0]
    band.copy_(seg.band_of(image.float()))
    masks = seg.segment_band(band)
    return seg.mine, masks[:len(seg.mine)].clone()
