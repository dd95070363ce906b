"""GPU parity of the device-side loss weight maps (csrc/weightmap.cuh through
ub_weight_map / ub_class_balance) against the CPU oracle (oracle/weightmap_oracle.py) and the golden outputs of
the reference's own functions.weighted_map / class_balance (functions.py:7-117).

Integer work (components, squared distances) is exact, so GPU == oracle bit for bit except where CUDA's and
numpy's double exp() differ in the last place AND that survives the rounding to fp32 (allowed: 1 float ulp on
< 0.01 % of the pixels).  Against the reference golden the bound is the oracle's own 2.5e-6 (OpenCV's sqrt)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close_to_oracle(w, ref):
    w, ref = w.cpu().numpy(), np.asarray(ref)
    diff = np.abs(w - ref)
    ulp = np.spacing(np.abs(ref).astype(np.float32))
    assert float((diff / ulp).max()) <= 1.0, float((diff / ulp).max())
    assert (w != ref).mean() < 1e-4, (w != ref).mean()


@pytest.mark.parametrize("kind,size,b,seed", [("blobs", 96, 2, 3), ("noise", 48, 1, 4), ("single", 64, 1, 5),
                                                ("noise", 128, 2, 11), ("blobs", 200, 3, 12), ("blobs", 37, 5, 13)])
@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8, torch.float32])
def test_weighted_map_vs_oracle(kind, size, b, seed, dtype):
    import weightmap_oracle as W
    from b200 import weightmap as G
    lab = W.synthetic_labels(b, size, seed, kind)
    ref = W.weighted_map(lab.astype(np.float32 if dtype == torch.float32 else np.int64))
    w = G.weighted_map(torch.from_numpy(lab).to(dtype).to(DEV))
    assert w.dtype == torch.float32 and w.shape == (b, size, size)
    close_to_oracle(w, ref)
    close_to_oracle(G.class_balance(torch.from_numpy(lab).to(dtype).to(DEV)), W.class_balance(lab))


@pytest.mark.parametrize("name", ["weightmap_blobs_96", "weightmap_noise_48", "weightmap_single_64", "weightmap_blobs_388"])
def test_weighted_map_vs_reference_golden(golden_dir, name):
    import weightmap_oracle as W
    from b200 import weightmap as G
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    lab = torch.from_numpy(W.synthetic_labels(int(g["batch"]), int(g["size"]), int(g["seed"]), str(g["kind"])))
    wi = G.weighted_map(lab.to(DEV)).cpu().numpy()
    wf = G.weighted_map(lab.float().to(DEV)).cpu().numpy()
    assert float(np.abs(wi - g["w_i64"]).max()) <= 2.5e-6 and float(np.abs(wf - g["w_f32"]).max()) <= 2.5e-6
    assert np.array_equal(G.class_balance(lab.to(DEV)).cpu().numpy(), g["cb"])


def test_component_labelling_hard_shapes():
    """Spirals, combs and checkerboards: long equivalence chains for the union-find, diagonal-only contacts that
    must NOT merge under 4-connectivity."""
    import weightmap_oracle as W
    from b200 import weightmap as G
    n = 64
    yy, xx = np.indices((n, n))
    spiral = np.zeros((n, n), dtype=np.int64)
    for k in range(0, n // 2, 4):                        # nested square rings, each opened once -> one long snake
        spiral[k, k:n - k] = spiral[n - 1 - k, k:n - k] = spiral[k:n - k, k] = spiral[k:n - k, n - 1 - k] = 1
        spiral[k + 1:k + 4, k] = 0
        spiral[k + 2, k:k + 4] = 1 if k + 4 < n // 2 else 0
    comb = ((xx % 4 == 0) | (yy == 0)).astype(np.int64)
    comb[:, 32:] = ((yy[:, 32:] % 4 == 0) & (xx[:, 32:] > 34)).astype(np.int64)
    checker = ((yy + xx) % 2 == 0).astype(np.int64)       # every pixel its own object under 4-connectivity
    checker[20:40, 20:40] = 0
    lab = np.stack([spiral, comb, checker])
    ref = W.weighted_map(lab)
    close_to_oracle(G.weighted_map(torch.from_numpy(lab).to(DEV)), ref)


def test_full_size_batch_and_properties():
    """BASELINE config 2 shape: 16 label tiles of 388x388 in one call, against the oracle, plus size-independent
    properties: foreground pixels carry exactly w_c = 1, the map is invariant under a flip of the label, and a
    background pixel far from every object carries w_c only."""
    import weightmap_oracle as W
    from b200 import weightmap as G
    lab = W.synthetic_labels(16, 388, 21, "blobs")
    t = torch.from_numpy(lab).to(DEV)
    w = G.weighted_map(t)
    close_to_oracle(w[:3], W.weighted_map(lab[:3]))
    assert bool((w[t == 1] == 1.0).all())
    wf = G.weighted_map(t.flip(2))
    assert torch.equal(wf.flip(2), w)
    wt = G.weighted_map(t.transpose(1, 2).contiguous())
    assert torch.equal(wt.transpose(1, 2), w)


def test_reference_error_behaviour_and_no_cpu_path():
    from b200 import weightmap as G
    with pytest.raises(IndexError):
        G.weighted_map(torch.zeros(2, 16, 16, dtype=torch.int64, device=DEV))
    with pytest.raises(IndexError):
        G.class_balance(torch.ones(1, 16, 16, dtype=torch.int64, device=DEV))
    with pytest.raises(RuntimeError):
        G.weighted_map(torch.zeros(1, 16, 16, dtype=torch.int64))            # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        G.weighted_map(torch.zeros(1, 2048, 2048, dtype=torch.uint8, device=DEV))   # beyond the packed-candidate limit
    mixed = torch.zeros(2, 16, 16, dtype=torch.int64, device=DEV)
    mixed[0, 3:6, 3:6] = 1
    with pytest.raises(IndexError):
        G.weighted_map(mixed)                                                # second sample has no object
    assert G.weighted_map(mixed, check=False).shape == (2, 16, 16)


def test_mask_metrics_match_reference_formulas():
    """functions.py:172-213: IoU = sum(p & l) / sum(p | l), pixel error = sum |p - l| / size."""
    import functions as Fn
    torch.manual_seed(3)
    pred = (torch.rand(388, 388, device=DEV) > 0.4).long()
    lab = (torch.rand(388, 388, device=DEV) > 0.5).long()
    m = Fn.evaluation_metrics(pred, lab)                         # CUDA masks -> device reduction
    p, l = pred.cpu().numpy(), lab.cpu().numpy()
    assert m.shape == (2, 1)
    assert m[0, 0] == np.sum(np.logical_and(p, l)) / np.sum(np.logical_or(p, l))
    assert m[1, 0] == np.sum(abs(p - l)) / p.size
    m8 = Fn.evaluation_metrics(pred.to(torch.uint8), lab)
    assert np.array_equal(m, m8)


@pytest.mark.parametrize("n", [388, 512, 250])
def test_mirror_transform_matches_reference_slices(n):
    """data.py:249-312: the reference's nine slice assignments, restated here literally (numpy) as the oracle."""
    from b200 import tiling
    from b200.geometry import input_size_compute
    torch.manual_seed(n)
    image = torch.rand(n, n)
    _, size, _ = input_size_compute(n)
    pad = (size - n) // 2
    img = image.numpy().astype(np.float64)
    ref = np.zeros((size, size))
    ref[pad:pad + n, pad:pad + n] = img
    ref[:pad, pad:pad + n] = img[pad:0:-1, 0:n]
    ref[:pad, :pad] = ref[0:pad, pad + pad:pad:-1]
    ref[:pad, pad + n:] = ref[0:pad, n + pad - 1:n - 1:-1]
    ref[n + pad:, pad:pad + n] = img[n:n - pad - 1:-1, 0:n]
    ref[n + pad:, :pad] = ref[n + pad:, pad + pad:pad:-1]
    ref[n + pad:, n + pad:] = ref[n + pad:, n + pad - 1:n - 1:-1]
    ref[pad:n + pad, :pad] = img[:, pad:0:-1]
    ref[pad:n + pad, n + pad:] = img[:, n - 1:n - 1 - pad:-1]
    out = tiling.mirror_transform(image.to(DEV))
    assert out.shape == (1, 1, size, size)
    assert np.array_equal(out[0, 0].cpu().numpy(), ref.astype(np.float32))       # a pure gather: bit-exact


def test_segment_image_equals_per_tile_forward():
    """Whole-image overlap-tile inference = the per-tile predict() masks put side by side, and sharding the tiles
    over two ranks gives the same masks (independent units, no collective)."""
    from network import Unet
    from b200 import tiling
    torch.manual_seed(1)
    net = Unet().to(DEV)
    image = torch.rand(300, 420, device=DEV)
    mask = tiling.segment_image(net, image, tile_in=252, batch=3)
    tile_out, halo, ny, nx, _, _ = tiling.plan(300, 420, 252)
    assert mask.shape == (300, 420) and mask.dtype == torch.uint8 and (ny, nx) == (5, 7)
    tiles = tiling.cut_tiles(image, 252)
    ref = torch.cat([net.predict(tiles[i:i + 1]) for i in range(ny * nx)]).to(torch.uint8)
    assert torch.equal(mask, tiling.stitch(ref, 300, 420, 252))
    parts = [tiling.segment_image(net, image, tile_in=252, batch=4, ranks=(r, 2)) for r in range(2)]
    full = torch.empty_like(ref)
    for idx, m in parts:
        full[idx] = m
    assert torch.equal(full, ref)


@pytest.mark.parametrize("shape", [(300, 420, 252), (389, 700, 572), (97, 130, 188)])
def test_cut_and_stitch_kernels_equal_the_torch_route(shape):
    """ub_cut_tiles / ub_stitch_tiles against b200.tiling.cut_tiles (F.pad reflect + replicate + unfold) / stitch:
    bit-equal, also when a rank holds only the band of rows its tiles read."""
    from b200 import tiling
    from b200.dist import partition
    h, w, tile_in = shape
    torch.manual_seed(h)
    image = torch.rand(h, w, device=DEV)
    ref = tiling.cut_tiles(image, tile_in)
    tile_out, halo, ny, nx, _, _ = tiling.plan(h, w, tile_in)
    n = ny * nx
    for world in (1, 3):
        got = torch.full_like(ref, float("nan"))
        for r in range(world):
            ids = list(range(*partition(n, r, world)))
            if not ids:
                continue
            row0, rows = tiling.needed_rows(h, ids, nx, tile_in, tile_out)
            band = image[row0:row0 + rows].contiguous()
            ids_dev = torch.tensor(ids, dtype=torch.int32, device=DEV)
            out = torch.empty(len(ids), 1, tile_in, tile_in, device=DEV)
            tiling.cut_tiles_device(band, h, w, row0, ids_dev, len(ids), nx, tile_in, tile_out, out)
            got[ids] = out
        assert torch.equal(got, ref)
    masks = (torch.rand(n, tile_out, tile_out, device=DEV) > 0.5).to(torch.uint8)
    full = torch.full((h, w), 9, dtype=torch.uint8, device=DEV)
    tiling.stitch_device(masks, torch.arange(n, dtype=torch.int32, device=DEV), n, nx, tile_out, full)
    assert torch.equal(full, tiling.stitch(masks, h, w, tile_in))


def test_tile_segmenter_pipeline_equals_segment_image():
    """TileSegmenter.segment_many (pinned host images in, pinned host masks out, upload / forward / download
    overlapped over two slots) yields for every image the mask segment_image computes."""
    from network import Unet
    from b200 import tiling
    torch.manual_seed(3)
    net = Unet().to(DEV)
    imgs = [torch.rand(300, 420).pin_memory() for _ in range(5)]
    want = [tiling.segment_image(net, im.to(DEV), tile_in=252, batch=4).cpu() for im in imgs]
    seg = tiling.TileSegmenter(net, 300, 420, tile_in=252, batch=4)
    n = 0
    for i, mask in seg.segment_many(imgs):
        assert i == n and torch.equal(mask, want[i])
        n += 1
    assert n == len(imgs)


def test_drop_in_functions_cpu_labels_go_through_the_device(golden_dir):
    """dl-unet_b200/functions.py keeps the reference's call placement (trainer.py:69,72 pass CPU labels): a CPU label
    batch is uploaded, processed by the kernels and returned on the CPU -- equal to the reference's golden outputs."""
    import functions
    import weightmap_oracle as W
    g = np.load(os.path.join(golden_dir, "class_balance.npz"))
    out = functions.class_balance(torch.from_numpy(g["lab"]))
    assert not out.is_cuda and np.array_equal(out.numpy(), g["w"])
    g = np.load(os.path.join(golden_dir, "weightmap_blobs_96.npz"))
    lab = torch.from_numpy(W.synthetic_labels(int(g["batch"]), int(g["size"]), int(g["seed"]), str(g["kind"])))
    w = functions.weighted_map(lab)
    assert not w.is_cuda and float(np.abs(w.numpy() - g["w_i64"]).max()) <= 2.5e-6
    assert functions.weighted_map(lab.to(DEV)).is_cuda
