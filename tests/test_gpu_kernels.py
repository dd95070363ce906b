"""GPU parity tests, per kernel, through the C ABI (b200.ops -> libunet_b200.so), against the matching
torch.nn.functional op on IDENTICAL (bf16-rounded) inputs.  Integer / index / copy results are
bit-exact; floating point within the tolerance written at each assert: a bf16-output kernel is
allowed the bf16 rounding of its own output (2^-9 relative per element, ~1.7e-3 rel-L2) on top of
fp32 accumulation-order noise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_OUT_TOL = 3e-3   # rel-L2 of a bf16-rounded output vs the fp32 reference
FP32_TOL = 2e-5       # rel-L2 of an fp32 output (accumulation order only)


@pytest.fixture(scope="module")
def ops():
    from b200 import ops as o
    return o


def bf(t):
    return t.to(torch.bfloat16)


def nhwc(t):
    return bf(t.permute(0, 2, 3, 1)).contiguous()


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


DEV = "cuda:0"
CONV_SHAPES = [(1, 20, 20, 64, 64), (2, 30, 30, 128, 128), (1, 18, 34, 64, 256), (1, 12, 12, 256, 512),
               (3, 10, 10, 128, 64), (1, 9, 9, 1024, 1024), (2, 70, 70, 64, 64), (1, 5, 3, 64, 64)]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_fwd_dgrad_wgrad(ops, shape):
    n, h, w, cin, cout = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
    b = torch.randn(cout, device=DEV)
    xq = bf(x).float().requires_grad_(True)
    wq = bf(wt).float().requires_grad_(True)
    ref = F.relu(F.conv2d(xq, wq, b))
    y = torch.full((n, h - 2, w - 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_fwd(nhwc(x), ops.pack_weight(wt, ops.PACK_CONV_FWD), b, y, relu=True)
    assert rel_l2(nchw(y), ref) < BF16_OUT_TOL
    # backward of the un-activated conv with a given upstream gradient
    dy = torch.randn(n, cout, h - 2, w - 2, device=DEV)
    dyq = bf(dy).float()
    gx, gw = torch.autograd.grad(F.conv2d(xq, wq), (xq, wq), dyq)
    dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_dgrad(nhwc(dy), ops.pack_weight(wt, ops.PACK_CONV_DGRAD), dx, None)
    assert rel_l2(nchw(dx), gx) < BF16_OUT_TOL
    ops.conv3x3_dgrad(nhwc(dy), ops.pack_weight(wt, ops.PACK_CONV_DGRAD), dx, nhwc(x))   # fused ReLU backward
    assert rel_l2(nchw(dx), gx * (xq > 0)) < BF16_OUT_TOL
    # split-K slabs are stored, never accumulated: a NaN-filled workspace must come back fully overwritten, and two
    # runs must agree bit for bit (no atomics)
    dwp = ops.wgrad_workspace(ops.PACK_CONV_FWD, nhwc(x), cout, fill=float("nan"))
    ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dwp)
    dw = torch.empty_like(wt)
    ops.unpack_grad(dwp, dw, ops.PACK_CONV_FWD)
    assert rel_l2(dw, gw) < FP32_TOL
    dwp2 = ops.wgrad_workspace(ops.PACK_CONV_FWD, nhwc(x), cout, fill=float("nan"))
    ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dwp2)
    assert torch.equal(dwp, dwp2)
    dw2 = dw.clone()
    ops.unpack_grad(dwp2, dw2, ops.PACK_CONV_FWD, accumulate=True)
    assert torch.equal(dw2, dw + dw)


# shapes that reach every weight-gradient kernel and its edge cases: the persistent tap-per-unit kernel with several
# units per CTA and both accumulators in use (items x splits > 148), a last M tile of one 64-row block (Cout = 320),
# N tiles of 64 / 128 / 256, maps that no chunk shape tiles exactly; the tap-row kernel (Cin = 256, wide maps); the
# haloed-patch kernel with two co blocks; the shifted-dY kernel (Cout = 64) with Cin = 128
WGRAD_SHAPES = [(4, 30, 30, 512, 512), (2, 58, 58, 1024, 512), (3, 22, 38, 256, 256), (2, 20, 20, 192, 320),
                (2, 68, 68, 256, 512), (1, 36, 164, 256, 128), (1, 40, 170, 256, 256), (2, 50, 66, 128, 256),
                (2, 90, 90, 128, 64), (16, 30, 30, 1024, 1024)]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_conv3x3_wgrad_shapes(ops, shape):
    n, h, w, cin, cout = shape
    torch.manual_seed(sum(shape))
    x = bf(torch.randn(n, cin, h, w, device=DEV).clamp_min(0))
    dy = bf(torch.randn(n, cout, h - 2, w - 2, device=DEV))
    wt = torch.zeros(cout, cin, 3, 3, device=DEV, requires_grad=True)
    (gw,) = torch.autograd.grad(F.conv2d(x.float(), wt), wt, dy.float())
    dwp = ops.wgrad_workspace(ops.PACK_CONV_FWD, nhwc(x), cout, fill=float("nan"))
    ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dwp)
    dw = torch.empty(cout, cin, 3, 3, device=DEV)
    ops.unpack_grad(dwp, dw, ops.PACK_CONV_FWD)
    assert rel_l2(dw, gw) < FP32_TOL
    dwp2 = ops.wgrad_workspace(ops.PACK_CONV_FWD, nhwc(x), cout, fill=float("nan"))
    ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dwp2)
    assert torch.equal(dwp, dwp2)


@pytest.mark.parametrize("shape", [(1, 6, 6, 64), (2, 45, 37, 64), (1, 70, 70, 128), (3, 20, 26, 64)])
def test_conv3x3_fwd_fused_head(ops, shape):
    """conv12e + finalconv in one launch (network.py:187-190): logits come from the fp32 accumulators, so they match
    the fp32 reference to fp32 tolerance -- tighter than conv -> bf16 store -> head; the stored activation still is
    the bf16-rounded conv output; y = None (inference) gives the same logits and the tester.py:30 argmax."""
    n, h, w, cin = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
    wt = torch.randn(64, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
    b = torch.randn(64, device=DEV)
    hw = torch.randn(2, 64, 1, 1, device=DEV) / 8
    hb = torch.randn(2, device=DEV)
    act = F.relu(F.conv2d(bf(x).float(), bf(wt).float(), b))
    ref = F.conv2d(act, hw, hb)
    wp = ops.pack_weight(wt, ops.PACK_CONV_FWD)
    y = torch.full((n, h - 2, w - 2, 64), float("nan"), dtype=torch.bfloat16, device=DEV)
    logits = torch.full((n, 2, h - 2, w - 2), float("nan"), device=DEV)
    ops.conv3x3_fwd_head(nhwc(x), wp, b, y, hw, hb, logits, None)
    assert rel_l2(nchw(y), act) < BF16_OUT_TOL
    assert rel_l2(logits, ref) < FP32_TOL
    y2 = torch.empty_like(y)
    ops.conv3x3_fwd(nhwc(x), wp, b, y2, relu=True)
    # same activation as the plain conv (which may run on the other conv kernel: accumulation order differs)
    assert rel_l2(y.float(), y2.float()) < 1e-3
    lg2 = torch.full_like(logits, float("nan"))
    am = torch.full((n, h - 2, w - 2), 7, dtype=torch.uint8, device=DEV)
    ops.conv3x3_fwd_head(nhwc(x), wp, b, None, hw, hb, lg2, am)        # inference: no activation store
    assert torch.equal(lg2, logits)
    assert torch.equal(am.long(), (logits[:, 1] > logits[:, 0]).long())
    am2 = torch.full_like(am, 7)
    ops.conv3x3_fwd_head(nhwc(x), wp, b, None, hw, hb, None, am2)      # argmax only
    assert torch.equal(am2, am)


@pytest.mark.parametrize("case", [(2, 18, 18, 64, 64, 12), (1, 34, 26, 64, 128, 20), (1, 70, 70, 128, 128, 8), (2, 23, 31, 64, 64, 9),
                                  (1, 12, 12, 256, 256, 4)])
def test_conv3x3_fwd_fused_pool(ops, case):
    """conv{l}2c + 2x2 max pool + argmax codes + skip copy in one launch == conv kernel followed by the pool kernel,
    bit for bit (same activation bits, same tie / NaN rules, same crop_and_concat placement); odd output sizes pool in
    floor mode; y = None never stores the full-resolution activation."""
    n, h, w, cin, cout, hu = case
    torch.manual_seed(sum(case))
    x = nhwc(torch.randn(n, cin, h, w, device=DEV).clamp_min(0))
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
    b = torch.randn(cout, device=DEV)
    wp = ops.pack_weight(wt, ops.PACK_CONV_FWD)
    ho, wo = h - 2, w - 2
    hp, wq = ho // 2, wo // 2
    coff = int((hp - hu) * 0.5)
    square = hp == wq and ho % 2 == 0 and wo % 2 == 0
    act = F.relu(F.conv2d(nchw(x), bf(wt).float(), b))
    outs = []
    for store in (True, False):
        y = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV) if store else None
        pooled = torch.full((n, hp, wq, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
        idx = torch.full((n, hp, wq, cout), 9, dtype=torch.uint8, device=DEV)
        cat = torch.zeros(n, hu, hu, 2 * cout, dtype=torch.bfloat16, device=DEV)
        ops.conv3x3_fwd_pool(x, wp, b, cout, y, pooled, idx, cat[..., :cout] if square else None, coff)
        outs.append((y, pooled, idx, cat))
    y, pooled, idx, cat = outs[0]
    assert rel_l2(nchw(y), act) < BF16_OUT_TOL                            # the conv itself
    # reference for the pooling part: the stand-alone pool kernel (parity-tested on its own) on the SAME activation
    p_ref = torch.empty_like(pooled)
    i_ref = torch.empty_like(idx)
    cat_ref = torch.zeros_like(cat)
    if square:
        ops.pool2x2_fwd(y, p_ref, i_ref, cat_ref[..., :cout], coff)
    else:                                                                  # the pool kernel wants even, dense input
        ops.pool2x2_fwd(y[:, :2 * hp, :2 * wq].contiguous(), p_ref, i_ref, None, 0)
    assert torch.equal(pooled, p_ref) and torch.equal(idx, i_ref) and torch.equal(cat, cat_ref)
    # ... and ATen itself on the stored activation: same values, same argmax (first maximum in window order)
    a_val, a_idx = F.max_pool2d(nchw(y).float(), 2, 2, return_indices=True)
    assert torch.equal(nchw(pooled).float(), a_val)
    code = idx.permute(0, 3, 1, 2).long()
    py = torch.arange(hp, device=DEV).view(1, 1, hp, 1)
    px = torch.arange(wq, device=DEV).view(1, 1, 1, wq)
    assert torch.equal((2 * py + (code >> 1)) * wo + 2 * px + (code & 1), a_idx)
    # without the activation store: identical pooled outputs
    assert torch.equal(outs[1][1], pooled) and torch.equal(outs[1][2], idx) and torch.equal(outs[1][3], cat)


@pytest.mark.parametrize("shape", [(1, 20, 20, 64), (2, 35, 51, 128), (3, 18, 19, 64), (1, 66, 34, 128), (4, 11, 83, 64), (2, 102, 100, 64)])
def test_wgrad_64_output_channels_tap_packing(ops, shape):
    """wgrad4.cuh (Cout = 64: conv12c, conv11e, conv12e): nine taps in ten accumulator slots (option 19, default: an
    N = 192 step with a one-pixel-shifted dY copy + an N = 128 step with a one-row-shifted copy) and in twelve (three
    N = 128 steps), both against fp32 autograd; ragged widths / heights exercise the extra chunk row and column the
    shifted copies need, NaN-filled slabs must come back fully overwritten, two runs agree bit for bit."""
    n, h, w, cin = shape
    cout = 64
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
    wt = torch.randn(cout, cin, 3, 3, device=DEV)
    dy = torch.randn(n, cout, h - 2, w - 2, device=DEV)
    xq, wq = bf(x).float(), bf(wt).float().requires_grad_(True)
    (gw,) = torch.autograd.grad(F.conv2d(xq, wq), wq, bf(dy).float())
    res = {}
    try:
        for pack in (1, 0):
            ops.set_option(19, pack)
            dwp = ops.wgrad_workspace(ops.PACK_CONV_FWD, nhwc(x), cout, fill=float("nan"))
            ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dwp)
            dw = torch.empty_like(wt)
            ops.unpack_grad(dwp, dw, ops.PACK_CONV_FWD)
            assert rel_l2(dw, gw) < FP32_TOL, pack
            dwp2 = ops.wgrad_workspace(ops.PACK_CONV_FWD, nhwc(x), cout, fill=float("nan"))
            ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dwp2)
            assert torch.equal(dwp, dwp2)
            res[pack] = dw
    finally:
        ops.set_option(19, 1)
    assert rel_l2(res[1], res[0]) < FP32_TOL


@pytest.mark.parametrize("shape", [(2, 20, 20, 64, 64, 0), (1, 70, 70, 64, 128, 64), (2, 33, 41, 128, 128, 0), (1, 12, 12, 256, 512, 0),
                                   (1, 104, 104, 256, 256, 0), (2, 30, 30, 512, 1024, 0), (2, 58, 58, 256, 512, 256),
                                   (3, 17, 23, 1024, 512, 0)])
def test_conv3x3_dgrad_fused_bias(ops, shape):
    """dgrad + column sums of dx over the channels >= c0 (the bias gradient of the layer that produced the input;
    c0 > 0: only the up-conv half of a concat buffer), fused into the epilogue or -- for shapes that run on the
    other kernel -- one extra pass: same numbers either way."""
    n, h, w, cout, cin, c0 = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
    dy = nhwc(torch.randn(n, cout, h - 2, w - 2, device=DEV))
    wd = ops.pack_weight(wt, ops.PACK_CONV_DGRAD)
    try:
        for pairs in (1, 0):        # 0: 256-column layers stay on the haloed-patch kernel (its 4-chunk column-sum path)
            ops.set_option(11, pairs)
            for mask in (nhwc(x), None):
                ref = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
                ops.conv3x3_dgrad(dy, wd, ref, mask)
                dx = torch.full_like(ref, float("nan"))
                db = torch.zeros(cin - c0, device=DEV)
                ops.conv3x3_dgrad(dy, wd, dx, mask, db, c0)
                assert torch.equal(dx, ref)
                assert rel_l2(db, ref[..., c0:].float().sum((0, 1, 2))) < FP32_TOL
                db2 = torch.zeros(cin - c0, device=DEV)
                ops.conv3x3_dgrad(dy, wd, dx, mask, db2, c0)
                assert torch.equal(db, db2)                        # fixed summation order
    finally:
        ops.set_option(11, 1)


def test_conv3x3_strided_views_write_in_place(ops):
    n, h, w, cin, cout = 2, 22, 22, 64, 128
    torch.manual_seed(1)
    big = bf(torch.randn(n, h + 6, w + 6, 192, device=DEV))
    xv = big[:, 3:3 + h, 3:3 + w, 64:128]
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / 24
    b = torch.randn(cout, device=DEV)
    ref = F.relu(F.conv2d(nchw(xv), bf(wt).float(), b))
    out = torch.zeros(n, h - 2, w - 2, 384, dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_fwd(xv, ops.pack_weight(wt, ops.PACK_CONV_FWD), b, out[..., 128:256], relu=True)
    assert rel_l2(nchw(out[..., 128:256]), ref) < BF16_OUT_TOL
    assert bool((out[..., :128] == 0).all()) and bool((out[..., 256:] == 0).all())   # neighbours untouched


@pytest.mark.parametrize("shape", [(1, 8, 8, 128, 64), (2, 14, 14, 1024, 512), (1, 25, 25, 256, 128), (1, 52, 52, 512, 256),
                                   (6, 28, 28, 1024, 512), (2, 98, 98, 128, 64), (9, 28, 28, 512, 256), (16, 52, 52, 256, 128)])
def test_upconv2x2(ops, shape):
    n, h, w, cin, cout = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
    wt = torch.randn(cin, cout, 2, 2, device=DEV) / cin ** 0.5
    b = torch.randn(cout, device=DEV)
    xq = bf(x).float().requires_grad_(True)
    wq = bf(wt).float().requires_grad_(True)
    ref = F.conv_transpose2d(xq, wq, b, stride=2)
    cat = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device=DEV)
    ops.upconv2x2_fwd(nhwc(x), ops.pack_weight(wt, ops.PACK_UP_FWD), b, cat[..., cout:])
    assert rel_l2(nchw(cat[..., cout:]), ref) < BF16_OUT_TOL
    assert bool((cat[..., :cout] == 0).all())                       # the skip half of the concat buffer is untouched
    dcat = bf(torch.randn(n, 2 * h, 2 * w, 2 * cout, device=DEV))
    dyv = dcat[..., cout:]
    gx, gw = torch.autograd.grad(ref, (xq, wq), nchw(dyv))
    dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.upconv2x2_dgrad(dyv, ops.pack_weight(wt, ops.PACK_UP_DGRAD), dx, nhwc(x))
    assert rel_l2(nchw(dx), gx * (xq > 0)) < BF16_OUT_TOL
    # ... with the column sums of dx (bias gradient of the conv below), fused on the CTA-pair kernel: same dx, same sums
    dx2 = torch.full_like(dx, float("nan"))
    db = torch.zeros(cin, device=DEV)
    ops.upconv2x2_dgrad(dyv, ops.pack_weight(wt, ops.PACK_UP_DGRAD), dx2, nhwc(x), dbias=db)
    assert torch.equal(dx2, dx)
    assert rel_l2(db, dx.float().sum((0, 1, 2))) < FP32_TOL
    db2 = torch.zeros(cin, device=DEV)
    ops.upconv2x2_dgrad(dyv, ops.pack_weight(wt, ops.PACK_UP_DGRAD), dx2, nhwc(x), dbias=db2)
    assert torch.equal(db, db2)                                     # fixed summation order
    dwp = ops.wgrad_workspace(ops.PACK_UP_FWD, nhwc(x), cout, fill=float("nan"))
    ops.upconv2x2_wgrad(nhwc(x), dyv, dwp)
    dw = torch.empty_like(wt)
    ops.unpack_grad(dwp, dw, ops.PACK_UP_FWD)
    assert rel_l2(dw, gw) < FP32_TOL
    dwp2 = ops.wgrad_workspace(ops.PACK_UP_FWD, nhwc(x), cout, fill=float("nan"))
    ops.upconv2x2_wgrad(nhwc(x), dyv, dwp2)
    assert torch.equal(dwp, dwp2)                                   # deterministic split-K


def test_pack_modes_exact(ops):
    wt = torch.randn(128, 64, 3, 3, device=DEV)
    assert torch.equal(ops.pack_weight(wt, 0), bf(wt.permute(0, 2, 3, 1).reshape(128, 576)))
    assert torch.equal(ops.pack_weight(wt, 1), bf(torch.flip(wt, (2, 3)).permute(1, 2, 3, 0).reshape(64, 9 * 128)))
    wu = torch.randn(128, 64, 2, 2, device=DEV)
    assert torch.equal(ops.pack_weight(wu, 2), bf(wu.permute(2, 3, 1, 0).reshape(256, 128)))
    assert torch.equal(ops.pack_weight(wu, 3), bf(wu.permute(0, 2, 3, 1).reshape(128, 256)))


def test_conv1(ops):
    torch.manual_seed(4)
    n, h, w = 2, 40, 44
    img = torch.rand(n, 1, h, w, device=DEV)
    w1 = torch.randn(64, 1, 3, 3, device=DEV).requires_grad_(True)
    b1 = torch.randn(64, device=DEV).requires_grad_(True)
    y = torch.empty(n, h - 2, w - 2, 64, dtype=torch.bfloat16, device=DEV)
    ops.conv1_fwd(img, w1.detach(), b1.detach(), y)
    pre = F.conv2d(img, w1, b1)
    assert rel_l2(nchw(y), F.relu(pre)) < BF16_OUT_TOL
    dy = bf(torch.randn(n, h - 2, w - 2, 64, device=DEV))
    gw, gb = torch.autograd.grad(pre, (w1, b1), nchw(dy))
    # CUDA-core comparator kernel (option 8 = 0): fp32 image, fp32 products
    ops.set_option(8, 0)
    try:
        dw, db = torch.zeros(64, 9, device=DEV), torch.zeros(64, device=DEV)
        ops.conv1_wgrad(img, dy, dw, db)
        assert rel_l2(dw.view(64, 1, 3, 3), gw) < FP32_TOL and rel_l2(db, gb) < FP32_TOL
    finally:
        ops.set_option(8, 1)


@pytest.mark.parametrize("shape", [(2, 40, 44), (3, 100, 200), (1, 67, 131), (16, 190, 190), (1, 3, 3)])
def test_conv1_wgrad_tensor_pipe(ops, shape):
    """conv11c weight + bias gradient as a K-major-B tcgen05 GEMM (wgrad.cuh, CONV1): exact products for a
    bf16-representable image (what the parity tests and the benchmark feed), image rounded to bf16 otherwise;
    deterministic partial rows."""
    n, h, w = shape
    torch.manual_seed(sum(shape))
    img = torch.rand(n, 1, h, w, device=DEV)
    dy = bf(torch.randn(n, h - 2, w - 2, 64, device=DEV))
    for image, tol in ((bf(img).float(), FP32_TOL), (img, 4e-3)):
        w1 = torch.zeros(64, 1, 3, 3, device=DEV, requires_grad=True)
        b1 = torch.zeros(64, device=DEV, requires_grad=True)
        gw, gb = torch.autograd.grad(F.conv2d(image, w1, b1), (w1, b1), nchw(dy).float())
        dw, db = torch.zeros(64, 9, device=DEV), torch.zeros(64, device=DEV)
        ops.conv1_wgrad(image, dy, dw, db)
        assert rel_l2(dw.view(64, 1, 3, 3), gw) < tol and rel_l2(db, gb) < FP32_TOL
    rows = ops.partial_rows(ops.PROD_CONV1, n, h, w)
    parts = [torch.full((rows, ops.CONV1_STRIDE), float("nan"), device=DEV) for _ in range(2)]
    for part in parts:
        ops.conv1_wgrad(img, dy, part=part)
    assert torch.equal(parts[0], parts[1]) and not torch.isnan(parts[0]).any()


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 12), (1, 8, 8, 128, 8), (2, 20, 20, 64, 8), (1, 24, 24, 256, 20), (1, 11, 11, 64, 7)])
def test_pool_indices_and_skip_write_bit_exact(ops, case):
    n, h, w, c, hu = case
    torch.manual_seed(sum(case))
    x = bf(torch.randn(n, h, w, c, device=DEV).clamp_min(0))        # post-ReLU: all-zero windows and ties abound
    hp = h // 2
    coff = int((hp - hu) * 0.5)
    y = torch.empty(n, hp, hp, c, dtype=torch.bfloat16, device=DEV)
    idx = torch.empty(n, hp, hp, c, dtype=torch.uint8, device=DEV)
    cat = torch.zeros(n, hu, hu, 2 * c, dtype=torch.bfloat16, device=DEV)
    ops.pool2x2_fwd(x, y, idx, cat[..., :c], coff)
    rv, ri = F.max_pool2d(nchw(x), 2, 2, return_indices=True)
    assert torch.equal(nchw(y), rv)                                  # values bit-exact
    oy = torch.arange(hp, device=DEV).view(1, 1, hp, 1)
    ox = torch.arange(hp, device=DEV).view(1, 1, 1, hp)
    code = idx.permute(0, 3, 1, 2).long()
    assert torch.equal((2 * oy + code // 2) * w + (2 * ox + code % 2), ri)   # ATen flat indices bit-exact
    assert torch.equal(nchw(cat[..., :c]), F.pad(rv, (-coff,) * 4))          # crop_and_concat placement bit-exact
    assert bool((cat[..., c:] == 0).all())
    # backward: sum of the two producers, ReLU mask, scatter
    dyp = bf(torch.randn(n, hp, hp, c, device=DEV))
    dcat = bf(torch.randn(n, hu, hu, 2 * c, device=DEV))
    dx = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device=DEV)
    dbias = torch.zeros(c, device=DEV)
    ops.pool2x2_bwd(dyp, dcat[..., :c], coff, idx, y, dx, dbias)
    # fused bias gradient = column sums of the scattered gradient, accumulated before its bf16 rounding
    gsum = (dyp.float() + F.pad(dcat[..., :c].float().permute(0, 3, 1, 2), (coff,) * 4).permute(0, 2, 3, 1)) * (y.float() > 0)
    assert rel_l2(dbias, gsum.sum((0, 1, 2))) < FP32_TOL
    assert rel_l2(dbias, dx.float().sum((0, 1, 2))) < BF16_OUT_TOL
    xr = nchw(x).requires_grad_(True)
    g = bf(nchw(dyp) + F.pad(nchw(dcat[..., :c]), (coff,) * 4)).float()
    F.max_pool2d(F.relu(xr), 2, 2).backward(g)
    assert torch.equal(nchw(dx), xr.grad)


def test_pool_nan_propagates(ops):
    x = torch.zeros(1, 2, 2, 64, dtype=torch.bfloat16, device=DEV)
    x[0, 1, 0, 3] = float("nan")
    x[0, 0, 1, 3] = 5.0
    y = torch.empty(1, 1, 1, 64, dtype=torch.bfloat16, device=DEV)
    idx = torch.empty(1, 1, 1, 64, dtype=torch.uint8, device=DEV)
    ops.pool2x2_fwd(x, y, idx, None, 0)
    rv, ri = F.max_pool2d(nchw(x), 2, 2, return_indices=True)
    assert torch.isnan(y[0, 0, 0, 3]) and torch.isnan(rv[0, 3, 0, 0])
    assert int(idx[0, 0, 0, 3]) == 2 and int(ri[0, 3, 0, 0]) == 2


def test_head_fwd_bwd(ops):
    torch.manual_seed(5)
    n, h, w = 2, 36, 36
    x = bf(torch.randn(n, h, w, 64, device=DEV).clamp_min(0))
    wf = (torch.randn(2, 64, 1, 1, device=DEV) / 8).requires_grad_(True)
    bfin = torch.randn(2, device=DEV).requires_grad_(True)
    logits = torch.empty(n, 2, h, w, device=DEV)
    am = torch.empty(n, h, w, dtype=torch.uint8, device=DEV)
    ops.head_fwd(x, wf.detach(), bfin.detach(), logits, am)
    xr = nchw(x).requires_grad_(True)
    ref = F.conv2d(F.relu(xr), wf, bfin)
    assert rel_l2(logits, ref) < 1e-4                                 # fp32 output
    assert torch.equal(am.long(), (logits[:, 1] > logits[:, 0]).long())   # argmax rule: ties -> class 0
    dl = torch.randn(n, 2, h, w, device=DEV)
    gx, gw, gb = torch.autograd.grad(ref, (xr, wf, bfin), dl)
    dx = torch.empty_like(x)
    dw, db = torch.zeros(2, 64, device=DEV), torch.zeros(2, device=DEV)
    dbx = torch.zeros(64, device=DEV)
    ops.head_bwd(x, wf.detach(), dl, dx, dw, db, dbx)
    assert rel_l2(dbx, dx.float().sum((0, 1, 2))) < FP32_TOL        # fused bias gradient of conv12e
    assert rel_l2(nchw(dx), gx) < BF16_OUT_TOL
    assert rel_l2(dw.view(2, 64, 1, 1), gw) < 1e-3 and rel_l2(db, gb) < FP32_TOL


@pytest.mark.parametrize("case", [(1, 36, 36), (3, 40, 36), (2, 388, 388), (2, 30, 30), (5, 44, 44)])
def test_losses_fwd_bwd(ops, case):
    b, hp, l = case
    torch.manual_seed(6)
    for scale in (10.0, 150.0):                                      # N(0,10^2) and the net's real logit scale
        z = (torch.randn(b, 2, hp, hp, device=DEV) * scale).requires_grad_(True)
        lab = (torch.rand(b, l, l, device=DEV) > 0.6).long()
        wm = 1 + 10 * torch.rand(b, l, l, device=DEV)
        pad = (hp - l) // 2
        zc = z[:, :, pad:pad + l, pad:pad + l]
        ws = ops.loss_workspace(DEV)
        for kind in (ops.LOSS_BCE, ops.LOSS_SOFTMAX_CE):
            if kind == ops.LOSS_BCE:       # trainer.py:63-75 with the per-pixel weight
                ref = F.binary_cross_entropy_with_logits(zc, torch.stack([1 - lab, lab], 1).float(), weight=wm.unsqueeze(1))
                gs = 1.0 / (2 * b * l * l)
            else:
                ref = (F.cross_entropy(zc, lab, reduction="none") * wm).mean()
                gs = 1.0 / (b * l * l)
            (gz,) = torch.autograd.grad(ref, z)
            for lt in (lab, lab.to(torch.uint8)):
                out = torch.zeros(1, device=DEV)
                dz = torch.zeros_like(z)
                ops.loss_fwd_bwd(kind, z.detach(), lt, wm, dz, out, ws, gs)
                assert abs(out.item() - ref.item()) <= 2e-5 * abs(ref.item())      # tolerance: fp32 sum order + fast exp/log
                assert rel_l2(dz, gz) < 1e-5


def test_colsum_sgd(ops):
    torch.manual_seed(7)
    g = bf(torch.randn(3, 10, 12, 256, device=DEV))
    out = torch.zeros(128, device=DEV)
    ops.colsum(g[..., 128:], out)
    assert rel_l2(out, g[..., 128:].float().sum((0, 1, 2))) < 1e-6
    nel = 1000003
    p = torch.randn(nel, device=DEV)
    pr = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([pr], lr=1e-4, momentum=0.99)
    buf = torch.zeros(nel, device=DEV)
    gr = torch.randn(nel, device=DEV)
    for step in range(3):
        pr.grad = gr * (step + 1)
        opt.step()
        ops.sgd_momentum(p, gr * (step + 1), buf, 1e-4, 0.99, step == 0)
    assert rel_l2(p, pr.detach()) < 1e-7
    assert rel_l2(buf, opt.state[pr]["momentum_buffer"]) < 1e-7


def test_bad_arguments_raise(ops):
    x = torch.zeros(1, 8, 8, 48, dtype=torch.bfloat16, device=DEV)    # channels not a multiple of 64
    y = torch.zeros(1, 6, 6, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        ops.conv3x3_fwd(x, torch.zeros(64, 9 * 48, dtype=torch.bfloat16, device=DEV), None, y)
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        ops.conv3x3_fwd(x, torch.zeros(64, 576, dtype=torch.bfloat16, device=DEV), None,
                        torch.zeros(1, 5, 6, 64, dtype=torch.bfloat16, device=DEV))


def test_guard_bands_untouched(ops):
    """compute-sanitizer is closed on this pool; instead every tensor-pipe output is carved out of a larger
    sentinel-filled buffer (guard rows before/after, guard channels either side) and the guards are checked:
    TMA stores must clip tile tails exactly and never touch neighbouring channels / pixels."""
    torch.manual_seed(11)
    SENT = 12345.0

    def carve(n, h, w, c):
        big = torch.full((n, h + 4, w + 4, c + 128), SENT, dtype=torch.bfloat16, device=DEV)
        return big, big[:, 2:2 + h, 2:2 + w, 64:64 + c]

    def guards_ok(big, n, h, w, c):
        inner = torch.zeros_like(big, dtype=torch.bool)
        inner[:, 2:2 + h, 2:2 + w, 64:64 + c] = True
        return bool((big[~inner] == SENT).all())

    for mode in (0, 2):
        ops.set_option(2, mode)
        for (n, h, w, cin, cout) in [(2, 21, 37, 64, 64), (1, 19, 23, 128, 128), (1, 30, 11, 64, 128)]:
            x = bf(torch.randn(n, h, w, cin, device=DEV))
            wf = bf(torch.randn(cout, 9 * cin, device=DEV) / 30)
            wd = bf(torch.randn(cin, 9 * cout, device=DEV) / 30)
            big, y = carve(n, h - 2, w - 2, cout)
            ops.conv3x3_fwd(x, wf, torch.zeros(cout, device=DEV), y, relu=True)
            assert guards_ok(big, n, h - 2, w - 2, cout) and not bool((y == SENT).any())
            dy = bf(torch.randn(n, h - 2, w - 2, cout, device=DEV))
            big, dx = carve(n, h, w, cin)
            ops.conv3x3_dgrad(dy, wd, dx, x)
            assert guards_ok(big, n, h, w, cin) and not bool((dx == SENT).any())
    ops.set_option(2, 1)
    # up-conv writes only its channel half, every pixel of it
    n, h, w, cin, cout = 1, 13, 9, 128, 64
    x = bf(torch.randn(n, h, w, cin, device=DEV))
    big, cat = carve(n, 2 * h, 2 * w, 2 * cout)
    ops.upconv2x2_fwd(x, bf(torch.randn(4 * cout, cin, device=DEV) / 12), torch.zeros(cout, device=DEV), cat[..., cout:])
    assert guards_ok(big, n, 2 * h, 2 * w, 2 * cout)
    assert bool((cat[..., :cout] == SENT).all()) and not bool((cat[..., cout:] == SENT).any())


@pytest.mark.parametrize("n_slabs", [1, 3, 9, 20, 75])
def test_unpack_range_sums_slabs_in_fixed_order(ops, n_slabs):
    """ops.JobTable.unpack_* (unpack_all_kernel): the per-layer slabs the wgrad kernels leave are summed and transposed
    to torch layout, one layer or all layers per launch, with 1 / 2 / 4 / 8 slab lanes per tile depending on the slab
    count.  Against the plain gather kernel (ub_unpack_grad) on random slabs; bit-identical from run to run."""
    torch.manual_seed(n_slabs)
    layers = [(0, 128, 64), (2, 64, 128), (0, 64, 192)]            # (kind, cout, cin)
    entries, refs = [], []
    for kind, cout, cin in layers:
        taps = 9 if kind == 0 else 4
        shape = (cout, cin, 3, 3) if kind == 0 else (cin, cout, 2, 2)
        slabs = torch.randn(n_slabs, cout * taps * cin, device=DEV)
        dw = torch.full(shape, float("nan"), device=DEV)
        w = torch.zeros(shape, device=DEV)
        fa = torch.empty(cout * taps * cin, dtype=torch.bfloat16, device=DEV)
        fb = torch.empty_like(fa)
        entries.append((w, fa, fb, slabs, dw, kind, cout, cin, n_slabs))
        ref = torch.empty(shape, device=DEV)
        ops.unpack_grad(slabs, ref, kind)
        refs.append(ref)
    # a vector job (kind 3): 130 elements from column 5 of partial rows 200 elements apart (the bias-gradient producers)
    part = torch.randn(4 * n_slabs + 1, 200, device=DEV)
    vec = torch.full((130,), float("nan"), device=DEV)
    entries.append((vec, vec, vec, part.view(-1)[5:], vec, 3, 130, 200, part.shape[0]))
    vref = torch.empty(130, device=DEV)
    ops.reduce_partials(part, part.shape[0], 200, 130, vref, accumulate=False, col=5)
    torch.testing.assert_close(vref, part[:, 5:135].double().sum(0).float(), rtol=1e-5, atol=1e-5)
    refs.append(vref)
    table = ops.JobTable(entries, DEV)
    table.unpack_layer(1)
    assert torch.isnan(entries[0][4]).all() and torch.isnan(entries[2][4]).all()       # only layer 1 was written
    torch.testing.assert_close(entries[1][4], refs[1], rtol=1e-5, atol=1e-5)
    table.unpack_all()
    first = [e[4].clone() for e in entries]
    for got, ref in zip(first, refs):
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
    table.unpack_all()
    assert all(torch.equal(e[4], f) for e, f in zip(entries, first))                    # deterministic
    table.unpack_all(accumulate=True)
    assert all(torch.equal(e[4], f + f) for e, f in zip(entries, first))


@pytest.mark.parametrize("shape", [(1, 12, 12, 256, 256), (2, 20, 21, 128, 256), (3, 13, 17, 256, 512), (1, 30, 30, 512, 1024),
                                   (5, 9, 40, 64, 256), (12, 40, 37, 256, 512), (16, 30, 30, 1024, 1024)])
def test_igemm_pair_kernel(ops, shape):
    """igemm_pair.cuh: the tap-per-load conv kernel on CTA pairs (tcgen05 cta_group::2, M = 256 pixels x N = 256 per
    instruction, each CTA loads half of the weight slab).  N tiles of 256 are forced so that small shapes reach it; odd
    tile counts leave one CTA of the last pair without work (zero-filled loads, clipped stores).  Same K order as the
    one-CTA kernel, so the two must agree bit for bit.  Stream-K (option 13 = 2: always; the two large shapes have more
    units than pairs, so units are really split) changes the summation order inside split units only: same tolerance
    against the fp32 reference, bit-identical from run to run."""
    n, h, w, cin, cout = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
    b = torch.randn(cout, device=DEV)
    xq, wq = bf(x).float().requires_grad_(True), bf(wt).float()
    dy = torch.randn(n, cout, h - 2, w - 2, device=DEV)
    ref = F.relu(F.conv2d(xq, wq, b))
    (gx,) = torch.autograd.grad(F.conv2d(xq, wq), xq, bf(dy).float())
    outs = {}
    ops.set_tuning(bn=256)
    try:
        for cfg in ((1, 0), (0, 0), (1, 2), (1, 2, "again")):
            ops.set_option(11, cfg[0])
            ops.set_option(13, cfg[1])
            y = torch.full((n, h - 2, w - 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            ops.conv3x3_fwd(nhwc(x), ops.pack_weight(wt, ops.PACK_CONV_FWD), b, y, relu=True)
            dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
            db = torch.zeros(cin, device=DEV)
            if cin % 256 == 0:      # the dgrad's N is Cin; the pair kernel sums the columns of dx in its epilogue
                ops.conv3x3_dgrad(nhwc(dy), ops.pack_weight(wt, ops.PACK_CONV_DGRAD), dx, nhwc(x), db)
            outs[cfg] = (y, dx, db)
    finally:
        ops.set_option(11, 1)
        ops.set_option(13, 1)
        ops.set_tuning()
    for cfg in ((1, 0), (1, 2)):
        y, dx, db = outs[cfg]
        assert rel_l2(nchw(y), ref) < BF16_OUT_TOL
        if cin % 256 == 0:
            assert rel_l2(nchw(dx), gx * (xq > 0)) < BF16_OUT_TOL
            assert rel_l2(db, dx.float().sum((0, 1, 2))) < FP32_TOL
    with_dgrad = cin % 256 == 0            # else dx was never written
    assert torch.equal(outs[(1, 0)][0], outs[(0, 0)][0])                                   # pairs == one CTA, bit for bit
    assert torch.equal(outs[(1, 2)][0], outs[(1, 2, "again")][0])                          # stream-K is deterministic
    if with_dgrad:
        assert torch.equal(outs[(1, 0)][1], outs[(0, 0)][1])
        assert all(torch.equal(a, b) for a, b in zip(outs[(1, 2)][1:], outs[(1, 2, "again")][1:]))
    if cin % 256 == 0:
        assert rel_l2(outs[(0, 0)][2], outs[(0, 0)][1].float().sum((0, 1, 2))) < FP32_TOL


@pytest.mark.parametrize("shape", [(16, 30, 30, 512, 256), (5, 30, 30, 256, 256), (16, 56, 56, 256, 256), (7, 54, 54, 256, 512),
                                   (16, 140, 140, 256, 256), (3, 32, 32, 256, 256)])
def test_igemm_pair_image_spanning_boxes(ops, shape):
    """CTA-pair kernel, 3x3 convolutions: TMA boxes that also span several images (option 18, default on: 28 x 28 maps
    fill the 128 MMA rows as 4 x 4 pixels x 8 images instead of 28 x 4 pixels) against boxes inside one image: same K
    order per output element, so forward and dgrad (ReLU mask) agree bit for bit; batch sizes that are no multiple of
    the image extent leave part of the last box beyond the tensor (zero-filled loads, clipped stores); the fused
    column sums group other rows per CTA (fp32 tolerance)."""
    n, h, w, cin, cout = shape
    torch.manual_seed(sum(shape))
    x = nhwc(torch.randn(n, cin, h, w, device=DEV).clamp_min(0))
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
    b = torch.randn(cout, device=DEV)
    dy = nhwc(torch.randn(n, cout, h - 2, w - 2, device=DEV))
    wf, wd = ops.pack_weight(wt, ops.PACK_CONV_FWD), ops.pack_weight(wt, ops.PACK_CONV_DGRAD)
    res = {}
    ops.set_tuning(bn=256)
    try:
        for boxes3d in (1, 0):
            ops.set_option(18, boxes3d)
            y = torch.full((n, h - 2, w - 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            ops.conv3x3_fwd(x, wf, b, y, relu=True)
            dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
            db = torch.zeros(cin, device=DEV)
            ops.conv3x3_dgrad(dy, wd, dx, x, db)
            res[boxes3d] = (y, dx, db)
    finally:
        ops.set_option(18, 1)
        ops.set_tuning()
    for t in res[1]:
        assert not bool(torch.isnan(t.float()).any())
    assert torch.equal(res[1][0], res[0][0])
    assert torch.equal(res[1][1], res[0][1])
    assert rel_l2(res[1][2], res[0][2]) < FP32_TOL
    assert rel_l2(nchw(res[1][0]), F.relu(F.conv2d(nchw(x), bf(wt).float(), b))) < BF16_OUT_TOL
    assert rel_l2(res[1][2], res[1][1].float().sum((0, 1, 2))) < FP32_TOL


@pytest.mark.parametrize("shape", [(1, 20, 20, 64, 64), (3, 27, 35, 64, 128), (2, 70, 38, 128, 64), (1, 40, 104, 128, 128),
                                   (5, 19, 21, 256, 128), (1, 100, 100, 64, 64)])
def test_igemm2_pair_equals_single(ops, shape):
    """The haloed-patch kernel on CTA pairs (option 17: tcgen05 cta_group::2, two adjacent pixel tiles per M = 256 MMA,
    half of every weight slab per CTA) against one CTA per tile: same K order, so forward (plain, fused pool, fused head)
    and dgrad (ReLU mask + fused column sums) agree bit for bit; odd tile counts leave the last pair's second CTA on a
    tile beyond the tensor (zero-filled loads, clipped stores, no pooled / logit writes)."""
    n, h, w, cin, cout = shape
    torch.manual_seed(sum(shape))
    x = nhwc(torch.randn(n, cin, h, w, device=DEV).clamp_min(0))
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
    b = torch.randn(cout, device=DEV)
    dy = nhwc(torch.randn(n, cout, h - 2, w - 2, device=DEV))
    wf, wd = ops.pack_weight(wt, ops.PACK_CONV_FWD), ops.pack_weight(wt, ops.PACK_CONV_DGRAD)
    hw_, hb_ = torch.randn(2, 64, device=DEV), torch.randn(2, device=DEV)
    ho, wo = h - 2, w - 2
    res = {}
    ops.set_option(2, 2)          # haloed-patch kernel wherever legal (small maps too)
    try:
        for pair in (1, 0):
            ops.set_option(17, pair)
            y = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            ops.conv3x3_fwd(x, wf, b, y, relu=True)
            dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
            db = torch.zeros(cin, device=DEV)
            ops.conv3x3_dgrad(dy, wd, dx, x, db)
            pooled = torch.full((n, ho // 2, wo // 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            idx = torch.full((n, ho // 2, wo // 2, cout), 9, dtype=torch.uint8, device=DEV)
            ops.conv3x3_fwd_pool(x, wf, b, cout, None, pooled, idx, None, 0)
            out = [y, dx, db, pooled, idx]
            if cout == 64:
                logits = torch.full((n, 2, ho, wo), float("nan"), device=DEV)
                ops.conv3x3_fwd_head(x, wf, b, None, hw_, hb_, logits=logits)
                out.append(logits)
            res[pair] = out
    finally:
        ops.set_option(17, 1)
        ops.set_option(2, 1)
    for i, (a, c) in enumerate(zip(res[1], res[0])):
        assert not bool(torch.isnan(a.float()).any())
        if i == 2:      # the column sums are grouped per CTA: same values, another fp32 summation order
            assert rel_l2(a, c) < FP32_TOL
        else:
            assert torch.equal(a, c)
    ref = F.relu(F.conv2d(nchw(x), bf(wt).float(), b))
    assert rel_l2(nchw(res[1][0]), ref) < BF16_OUT_TOL
