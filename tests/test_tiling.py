"""CPU: overlap-tile planner (b200/tiling.py) -- pure index arithmetic around the kernel path."""
import torch

from b200 import tiling


def test_plan_4096():
    tile_out, halo, ny, nx, ph, pw = tiling.plan(4096, 4096)
    assert (tile_out, halo, ny, nx) == (388, 92, 11, 11)          # BASELINE config 4: 121 tiles
    assert ph == pw == 11 * 388 + 184


def test_cut_and_stitch_round_trip():
    for (h, w) in ((800, 900), (388, 388), (389, 1000)):
        img = torch.rand(h, w)
        t = tiling.cut_tiles(img)
        tile_out, halo, ny, nx, _, _ = tiling.plan(h, w)
        assert t.shape == (ny * nx, 1, 572, 572)
        centre = t[:, 0, halo:halo + tile_out, halo:halo + tile_out]
        assert torch.equal(tiling.stitch(centre, h, w), img)        # tile centres tile the image exactly


def test_halo_is_mirror_without_edge_repeat():
    img = torch.arange(400 * 400, dtype=torch.float32).reshape(400, 400)
    t = tiling.cut_tiles(img)
    halo = 92
    # first tile: the row above the image's first row is the image's SECOND row (reflect, data.py:249-277)
    assert torch.equal(t[0, 0, halo - 1, halo:halo + 388], img[1, :388])
    assert torch.equal(t[0, 0, halo:halo + 388, halo - 3], img[:388, 3])


def test_src_index_and_band_match_the_pad_and_unfold_route():
    """The index rule ub_cut_tiles implements (tiling._src_index) against the F.pad(reflect) + F.pad(replicate) +
    unfold reference, and needed_rows: every source row of a rank's windows lies inside its band."""
    from b200.dist import partition
    for (h, w, tile_in) in ((300, 420, 252), (389, 1000, 572), (97, 130, 188)):
        img = torch.rand(h, w)
        ref = tiling.cut_tiles(img, tile_in)
        tile_out, halo, ny, nx, _, _ = tiling.plan(h, w, tile_in)
        for t in (0, nx - 1, ny * nx - 1, (ny // 2) * nx + nx // 2):
            ty, tx = divmod(t, nx)
            ys = torch.tensor([tiling._src_index(ty * tile_out + p - halo, h, halo) for p in range(tile_in)])
            xs = torch.tensor([tiling._src_index(tx * tile_out + p - halo, w, halo) for p in range(tile_in)])
            assert torch.equal(img[ys][:, xs], ref[t, 0])
        for world in (1, 2, 3):
            for r in range(world):
                ids = list(range(*partition(ny * nx, r, world)))
                if not ids:
                    continue
                row0, rows = tiling.needed_rows(h, ids, nx, tile_in, tile_out)
                used = {tiling._src_index(ty * tile_out + p - halo, h, halo) for ty in {t // nx for t in ids} for p in range(tile_in)}
                assert row0 <= min(used) and max(used) < row0 + rows <= h
