"""Tensor-level wrappers over the C ABI: each function takes torch CUDA tensors, builds the
ub_tensor4 views and launches on torch's current stream.  Activations are NHWC bf16 tensors of
shape [N, H, W, C] (possibly strided views with channel stride 1).  No function here computes
anything in PyTorch -- torch is used for device memory and streams only."""
import ctypes as C

import torch

from . import lib as _lib
from .lib import Tensor4, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Stats:
    """Launch accounting for bench.py: `launches` counts every kernel this library launched; with `timing` on, each
    launch is bracketed by CUDA events on the launching stream and recorded as (family, algorithmic FLOPs, algorithmic
    bytes, start, end).  Families: "igemm:<what>" / "wgrad:<what>" (tensor pipe), "hbm:<kernel>" (HBM-bound)."""

    def __init__(self):
        self.launches = 0
        self.timing = False
        self.records = []

    def reset(self):
        self.launches = 0
        self.records = []

    def summarize(self, prefix=None):
        """-> {family: (flops, bytes, total_ms, n_launches)}; call after a synchronize.  prefix: merge every family that
        starts with it into one entry."""
        out = {}
        for fam, fl, by, e0, e1 in self.records:
            if prefix is not None:
                if not fam.startswith(prefix):
                    continue
                fam = prefix
            t = e0.elapsed_time(e1)
            f, b, ms, n = out.get(fam, (0.0, 0.0, 0.0, 0))
            out[fam] = (f + fl, b + by, ms + t, n + 1)
        return out


STATS = _Stats()


class _timed:
    def __init__(self, family, work, flops=None):
        """work: algorithmic bytes for "hbm:*" families, algorithmic FLOPs otherwise; flops / work may both be given for
        kernels reported against either roof (up-convolutions): _timed(fam, bytes, flops)."""
        self.family = family
        if flops is not None:
            self.flops, self.bytes = flops, work
        elif family.startswith("hbm:"):
            self.flops, self.bytes = 0.0, work
        else:
            self.flops, self.bytes = work, 0.0

    def __enter__(self):
        if STATS.timing:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if STATS.timing:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            STATS.records.append((self.family, self.flops, self.bytes, self.e0, e1))
        return False


_RESERVE = {"left": 0}


def reserve_sms(n_free, kernels):
    """Size the persistent grids of the next `kernels` launches for (SM count - n_free) SMs: a collective that was just
    enqueued on another stream finds free SMs at once instead of waiting for CTAs to drain (and our CTAs do not queue
    behind it).  n_free = 0 restores full-size grids."""
    lib = _lib.load()
    if n_free <= 0 or kernels <= 0:
        lib.ub_set_sm_limit(0)
        _RESERVE["left"] = 0
        return
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    lib.ub_set_sm_limit(max(1, sms - int(n_free)))
    _RESERVE["left"] = int(kernels)


def _chk(code, what=""):
    """check() + launch accounting (every successful call below is exactly one kernel launch)."""
    check(code, what)
    STATS.launches += 1
    if _RESERVE["left"] > 0:
        _RESERVE["left"] -= 1
        if _RESERVE["left"] == 0:
            _lib.load().ub_set_sm_limit(0)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("unet_b200 ops need CUDA tensors: there is no CPU path")


def t4(t):
    """[N,H,W,C] bf16 view -> ub_tensor4."""
    if t.dtype != torch.bfloat16 or t.dim() != 4 or t.stride(3) != 1:
        raise RuntimeError(f"expected an NHWC bf16 view with unit channel stride, got {t.dtype} {tuple(t.shape)} {t.stride()}")
    n, h, w, c = t.shape
    sn = t.stride(0) if n > 1 else max(t.stride(0), h * t.stride(1))
    return Tensor4(t.data_ptr(), n, h, w, c, sn, t.stride(1), t.stride(2))


def _t4p(t):
    return C.byref(t4(t)) if t is not None else None


def set_tuning(bn=0, bw=0, bh=0, splits=0):
    _lib.load().ub_set_tuning(bn, bw, bh, splits)


OPT_STREAM_K = 13
_options = {}          # what set_option was last called with (the library has no getter)


def set_option(key, value):
    _lib.load().ub_set_option(int(key), int(value))
    _options[int(key)] = int(value)


def get_option(key, default):
    return _options.get(int(key), default)


def device_error():
    code = C.c_uint(0)
    rc = _lib.load().ub_device_error(C.byref(code))
    return rc, code.value


# ---------------------------------------------------------------- weights
PACK_CONV_FWD, PACK_CONV_DGRAD, PACK_UP_FWD, PACK_UP_DGRAD = 0, 1, 2, 3


def pack_weight(w, mode, out=None):
    """fp32 master in torch layout -> bf16 K-major GEMM operand (see ub_pack_weight)."""
    _need_cuda(w)
    w = w.detach()
    if not w.is_contiguous() or w.dtype != torch.float32:
        raise RuntimeError("pack_weight expects a contiguous fp32 weight")
    if mode in (PACK_CONV_FWD, PACK_CONV_DGRAD):
        cout, cin = w.shape[0], w.shape[1]
        shape = (cout, 9 * cin) if mode == PACK_CONV_FWD else (cin, 9 * cout)
    else:
        cin, cout = w.shape[0], w.shape[1]
        shape = (4 * cout, cin) if mode == PACK_UP_FWD else (cin, 4 * cout)
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    _chk(_lib.load().ub_pack_weight(_ptr(w), _ptr(out), mode, cout, cin, _stream()), "pack_weight")
    return out


def wgrad_slabs(kind, n, h_x, w_x, cin, cout):
    """Number of split-K slabs the weight-gradient kernel of this shape writes (kind: PACK_CONV_FWD for a 3x3 conv
    whose INPUT is h_x x w_x, PACK_UP_FWD for an up-convolution).  A pure host query."""
    r = _lib.load().ub_wgrad_slabs(int(kind), int(n), int(h_x), int(w_x), int(cin), int(cout))
    if r < 0:
        check(r, "wgrad_slabs")
    return r


def wgrad_workspace(kind, x, cout, fill=None):
    """Slab workspace [n_slabs, numel] for conv3x3_wgrad (kind PACK_CONV_FWD) / upconv2x2_wgrad (PACK_UP_FWD) of the NHWC
    input x; the kernels overwrite every element (fill: test pattern to prove it)."""
    n, h, w, cin = x.shape
    taps = 9 if kind == PACK_CONV_FWD else 4
    ws = torch.empty(wgrad_slabs(kind, n, h, w, cin, cout), cout * taps * cin, dtype=torch.float32, device=x.device)
    if fill is not None:
        ws.fill_(fill)
    return ws


def unpack_grad(slabs, dw, mode, accumulate=False):
    """slabs: fp32 [n_slabs, packed layout] as written by conv3x3_wgrad / upconv2x2_wgrad -> dw (torch layout),
    summed in slab order."""
    if mode == PACK_CONV_FWD:
        cout, cin = dw.shape[0], dw.shape[1]
    else:
        cin, cout = dw.shape[0], dw.shape[1]
    slabs = slabs.reshape(-1, dw.numel())
    _chk(_lib.load().ub_unpack_grad(_ptr(slabs), slabs.shape[0], dw.numel(), _ptr(dw), mode, cout, cin, int(accumulate),
                                    _stream()), "unpack_grad")
    return dw


class JobTable:
    """Device-side table of all conv / up-conv layers for the one-launch pack / unpack kernels."""

    def __init__(self, entries, device):
        """entries: list of (w fp32 torch-layout, dst_fwd bf16, dst_dgrad bf16, grad slabs fp32, dw fp32, kind, cout, cin,
        n_slabs)."""
        n = len(entries)
        lib = _lib.load()
        nbytes = lib.ub_job_table_bytes(n)
        host = torch.empty(nbytes, dtype=torch.uint8)
        arr = lambda col: (C.c_void_p * n)(*[e[col].data_ptr() for e in entries])
        iarr = lambda col: (C.c_int * n)(*[int(e[col]) for e in entries])
        ubegin = (C.c_int * (n + 1))()
        tiles = lib.ub_job_table_fill(C.c_void_p(host.data_ptr()), n, arr(0), arr(1), arr(2), arr(3), arr(4), iarr(5), iarr(6),
                                      iarr(7), iarr(8), ubegin)
        if tiles < 0:
            check(tiles, "job_table_fill")
        self.table = host.to(device)
        self.n = n
        self.total = tiles
        self.utile_begin = list(ubegin)
        self.keep = entries  # keep the tensors alive

    def _numel(self):
        return sum(e[0].numel() for e in self.keep if e[5] != 3)

    def pack_all(self):
        with _timed("hbm:pack_all", self._numel() * 8.0):                  # fp32 master in, two bf16 operands out
            _chk(_lib.load().ub_pack_all(_ptr(self.table), self.n, self.total, _stream()), "pack_all")

    def unpack_all(self, accumulate=False):
        """Every layer: sum the gradient slabs in fixed order and write the torch-layout gradients."""
        with _timed("hbm:unpack_all", sum(e[4].numel() * 4.0 * (e[8] + 1) for e in self.keep)):   # slabs in, gradient out
            _chk(_lib.load().ub_unpack_range(_ptr(self.table), self.n, 0, self.utile_begin[self.n], int(accumulate), _stream()),
                 "unpack_all")

    def unpack_layer(self, i, accumulate=False, j=None):
        """Jobs i .. j-1 (default: job i alone)."""
        j = i + 1 if j is None else j
        with _timed("hbm:unpack_layer", sum(e[4].numel() * 4.0 * (e[8] + 1) for e in self.keep[i:j])):
            _chk(_lib.load().ub_unpack_range(_ptr(self.table), self.n, self.utile_begin[i], self.utile_begin[j],
                                             int(accumulate), _stream()), "unpack_layer")


# ---------------------------------------------------------------- tensor-pipe ops
def conv3x3_fwd(x, w_fwd, bias, y, relu=True):
    _need_cuda(x, w_fwd, y)
    n, h, w, co = y.shape
    with _timed("igemm:conv_fwd", 2.0 * n * h * w * co * 9 * x.shape[3]):
        _chk(_lib.load().ub_conv3x3_fwd(_t4p(x), _ptr(w_fwd), _ptr(bias), _t4p(y), int(relu), _stream()), "conv3x3_fwd")
    return y


def conv3x3_fwd_head(x, w_fwd, bias, y, head_w, head_b, logits=None, argmax=None):
    """conv12e (bias + ReLU) with finalconv evaluated in the epilogue on the fp32 accumulators
    (network.py:187-190).  y = None: the 64-channel activation is not stored (inference)."""
    _need_cuda(x, w_fwd, head_w, head_b, y, logits, argmax)
    n, h, w, ci = x.shape
    with _timed("igemm:conv_fwd_head", 2.0 * n * (h - 2) * (w - 2) * 64 * 9 * ci):
        _chk(_lib.load().ub_conv3x3_fwd_head(_t4p(x), _ptr(w_fwd), _ptr(bias), _t4p(y), _ptr(head_w), _ptr(head_b),
                                             _ptr(logits), _ptr(argmax), _stream()), "conv3x3_fwd_head")


def conv3x3_fwd_pool(x, w_fwd, bias, cout, y, pooled, idx, cat=None, c_off=0):
    """3x3 conv (bias + ReLU) with the 2x2 max pool, its argmax codes and the skip copy produced by the same launch
    (network.py:132-133).  y = None: the full-resolution activation is never stored."""
    _need_cuda(x, w_fwd, pooled, idx, y, cat)
    n, h, w, ci = x.shape
    with _timed("igemm:conv_fwd_pool", 2.0 * n * (h - 2) * (w - 2) * cout * 9 * ci):
        _chk(_lib.load().ub_conv3x3_fwd_pool(_t4p(x), _ptr(w_fwd), _ptr(bias), int(cout), _t4p(y), _ptr(pooled), _ptr(idx),
                                             _t4p(cat), int(c_off), _stream()), "conv3x3_fwd_pool")


def conv3x3_dgrad(dy, w_dgrad, dx, relu_src=None, dbias=None, dbias_c0=0, dbias_part=None):
    """dbias_part: [partial_rows(PROD_DGRAD, n, h, w, C, Cdy, c0), C - c0] also receives partial rows of the column sums of
    dx over the channels >= c0 -- the bias gradient of the layer that produced this layer's input (fused into the
    epilogue where supported); dbias (fp32 [C - c0]): the same, reduced here and ADDED to it."""
    _need_cuda(dy, w_dgrad, dx)
    n, h, w, co = dy.shape
    own = dbias is not None and dbias_part is None
    length = dx.shape[3] - dbias_c0
    if own:
        dbias_part = _part(partial_rows(PROD_DGRAD, n, dx.shape[1], dx.shape[2], dx.shape[3], co, dbias_c0), length, dx)
    with _timed("igemm:conv_dgrad", 2.0 * n * h * w * co * 9 * dx.shape[3]):
        if dbias_part is None:
            _chk(_lib.load().ub_conv3x3_dgrad(_t4p(dy), _ptr(w_dgrad), _t4p(dx), _t4p(relu_src), _stream()), "conv3x3_dgrad")
        else:
            _chk(_lib.load().ub_conv3x3_dgrad_bias(_t4p(dy), _ptr(w_dgrad), _t4p(dx), _t4p(relu_src), _ptr(dbias_part),
                                                   int(dbias_c0), _stream()), "conv3x3_dgrad_bias")
    if own:
        reduce_partials(dbias_part, dbias_part.shape[0], length, length, dbias, True)
    return dx


def conv3x3_wgrad(x, dy, dw_packed):
    """dw_packed: fp32 slab workspace of wgrad_slabs(...) * Cout * 9 * Cin elements, fully overwritten (no zeroing,
    no atomics: the slabs are summed by unpack_grad / JobTable.unpack_*)."""
    _need_cuda(x, dy, dw_packed)
    n, h, w, co = dy.shape
    with _timed("wgrad:conv", 2.0 * n * h * w * co * 9 * x.shape[3]):
        _chk(_lib.load().ub_conv3x3_wgrad(_t4p(x), _t4p(dy), _ptr(dw_packed), _stream()), "conv3x3_wgrad")
    return dw_packed


def upconv2x2_fwd(x, w_fwd, bias, y):
    _need_cuda(x, w_fwd, y)
    n, h, w, ci = x.shape
    # bytes: x read, y written (bf16) -- the up-convolutions of levels 1-2 are HBM-bound
    with _timed("igemm:up_fwd", n * h * w * 2.0 * (ci + 4 * y.shape[3]), 2.0 * n * h * w * ci * 4 * y.shape[3]):
        _chk(_lib.load().ub_upconv2x2_fwd(_t4p(x), _ptr(w_fwd), _ptr(bias), _t4p(y), _stream()), "upconv2x2_fwd")
    return y


def upconv2x2_dgrad(dy, w_dgrad, dx, relu_src=None, dbias=None, dbias_part=None):
    """dbias_part: [partial_rows(PROD_UPDGRAD, n, h, w, C), C] also receives partial rows of the column sums of dx -- the
    bias gradient of the conv that produced the up-convolution's input; dbias (fp32 [C]): reduced here and ADDED to it."""
    _need_cuda(dy, w_dgrad, dx)
    n, h, w, ci = dx.shape
    own = dbias is not None and dbias_part is None
    if own:
        dbias_part = _part(partial_rows(PROD_UPDGRAD, n, h, w, ci), ci, dx)
    with _timed("igemm:up_dgrad", n * h * w * 2.0 * (4 * dy.shape[3] + ci * (2 if relu_src is not None else 1)),
                2.0 * n * h * w * ci * 4 * dy.shape[3]):
        if dbias_part is None:
            _chk(_lib.load().ub_upconv2x2_dgrad(_t4p(dy), _ptr(w_dgrad), _t4p(dx), _t4p(relu_src), _stream()), "upconv2x2_dgrad")
        else:
            _chk(_lib.load().ub_upconv2x2_dgrad_bias(_t4p(dy), _ptr(w_dgrad), _t4p(dx), _t4p(relu_src), _ptr(dbias_part),
                                                     _stream()), "upconv2x2_dgrad_bias")
    if own:
        reduce_partials(dbias_part, dbias_part.shape[0], ci, ci, dbias, True)
    return dx


def upconv2x2_wgrad(x, dy, dw_packed):
    _need_cuda(x, dy, dw_packed)
    n, h, w, ci = x.shape
    with _timed("wgrad:up", n * h * w * 2.0 * (ci + 4 * dy.shape[3]), 2.0 * n * h * w * ci * 4 * dy.shape[3]):
        _chk(_lib.load().ub_upconv2x2_wgrad(_t4p(x), _t4p(dy), _ptr(dw_packed), _stream()), "upconv2x2_wgrad")
    return dw_packed


# ---------------------------------------------------------------- partial sums (bias gradients)
# Kernels that reduce over pixels into a short vector store one row of partial sums per CTA (no atomics); the rows are
# summed in order by the unpack job table (engine) or by reduce_partials (the `dbias=` / `out=` forms below, which keep
# the old "accumulate into a [C] tensor" semantics for tests and tools).
PROD_POOL, PROD_COLSUM, PROD_DGRAD, PROD_HEAD, PROD_CONV1, PROD_UPDGRAD = 0, 1, 2, 3, 4, 5
HEAD_STRIDE, CONV1_STRIDE = 194, 640


def partial_rows(producer, n, h, w, c=0, c_dy=0, c0=0):
    r = _lib.load().ub_partial_rows(int(producer), int(n), int(h), int(w), int(c), int(c_dy), int(c0))
    if r < 0:
        check(r, "partial_rows")
    return r


def reduce_partials(part, rows, stride, length, dst, accumulate=True, col=0):
    """dst[e] (+)= sum over the first `rows` rows of part[:, col + e], in row order."""
    base = C.c_void_p(part.data_ptr() + 4 * col)
    _chk(_lib.load().ub_reduce_partials(base, int(rows), int(stride), int(length), _ptr(dst), int(accumulate), _stream()),
         "reduce_partials")
    return dst


def _part(rows, stride, like):
    return torch.empty(rows, stride, dtype=torch.float32, device=like.device)


# ---------------------------------------------------------------- HBM-bound ops
def conv1_fwd(image, w, bias, y):
    """image fp32 [N,1,H,W] -> y bf16 [N,H-2,W-2,64], bias + ReLU fused (network.py:131)."""
    _need_cuda(image, w, bias, y)
    n, _, h, wd = image.shape
    with _timed("hbm:conv1_fwd", n * (h * wd * 4.0 + (h - 2) * (wd - 2) * 128.0)):
        _chk(_lib.load().ub_conv1_fwd(_ptr(image), _ptr(w), _ptr(bias), _ptr(y), n, h, wd, _stream()), "conv1_fwd")
    return y


_CONV1_WS = {}


def conv1_wgrad(image, dy, dw=None, db=None, part=None, ws=None):
    """conv11c weight + bias gradient.  part: [partial_rows(PROD_CONV1, n, h, w), 640] receives the partial rows
    (dW[64][9] then db[64]); otherwise they are reduced here and ADDED to dw / db."""
    n, _, h, wd = image.shape
    own = part is None
    if own:
        part = _part(partial_rows(PROD_CONV1, n, h, wd), CONV1_STRIDE, image)
    if ws is None:      # scratch for the bf16 image copy (tensor-pipe path)
        key = (n, h, wd, str(image.device))
        ws = _CONV1_WS.get(key)
        if ws is None:
            ws = _CONV1_WS[key] = torch.empty(_lib.load().ub_conv1_wgrad_workspace_bytes(n, h, wd), dtype=torch.uint8,
                                              device=image.device)
    with _timed("hbm:conv1_wgrad", n * (h * wd * 4.0 + (h - 2) * (wd - 2) * 128.0)):
        _chk(_lib.load().ub_conv1_wgrad(_ptr(image), _ptr(dy), _ptr(part), _ptr(ws), n, h, wd, _stream()), "conv1_wgrad")
        STATS.launches += 1   # the image conversion kernel
    if own:
        reduce_partials(part, part.shape[0], CONV1_STRIDE, 576, dw, True)
        reduce_partials(part, part.shape[0], CONV1_STRIDE, 64, db, True, col=576)


def pool2x2_fwd(x, y, idx, cat=None, c_off=0):
    _need_cuda(x, y, idx)
    n, h, w, c = x.shape
    # 4 x 2 B read + 2 B value + 1 B index per output element (+ 2 B for the skip copy)
    with _timed("hbm:pool_fwd", n * (h // 2) * (w // 2) * c * (13.0 if cat is not None else 11.0)):
        _chk(_lib.load().ub_pool2x2_fwd(_ptr(x), _ptr(y), _ptr(idx), _t4p(cat), int(c_off), n, h, w, c, _stream()), "pool2x2_fwd")
    return y, idx


def pool2x2_bwd(dy, dcat, c_off, idx, pooled, dx, dbias=None, dbias_part=None):
    """dbias_part: [partial_rows(PROD_POOL, n, h, w, c), c] receives partial rows of the column sums of dx (bias gradient
    of the conv that fed the pool); dbias (fp32 [C]): the same, reduced here and ADDED to it."""
    n, h, w, c = dx.shape
    own = dbias is not None and dbias_part is None
    if own:
        dbias_part = _part(partial_rows(PROD_POOL, n, h, w, c), c, dx)
    # 2 B dY + 1 B index read + 4 x 2 B dX written per pooled element (the second gradient producer
    # and the ReLU mask this kernel also reads are not counted)
    with _timed("hbm:pool_bwd", n * (h // 2) * (w // 2) * c * 11.0):
        _chk(_lib.load().ub_pool2x2_bwd(_ptr(dy), _t4p(dcat), int(c_off), _ptr(idx), _ptr(pooled), _ptr(dx), _ptr(dbias_part),
                                         n, h, w, c, _stream()), "pool2x2_bwd")
    if own:
        reduce_partials(dbias_part, dbias_part.shape[0], c, c, dbias, True)
    return dx


def head_fwd(x, w, bias, logits=None, argmax=None):
    n, h, wd, _ = x.shape
    _chk(_lib.load().ub_head_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(logits), _ptr(argmax), n, h, wd, _stream()), "head_fwd")


def head_bwd(x, w, dlogits, dx, dw=None, db=None, dbias_x=None, part=None, with_dx_sums=True):
    """part: [partial_rows(PROD_HEAD, n, h, w), 194] receives the partial rows (dW[2][64], db[2], column sums of dx[64] =
    bias gradient of conv12e); otherwise they are reduced here and ADDED to dw / db / dbias_x."""
    n, h, wd, _ = x.shape
    own = part is None
    if own:
        part = _part(partial_rows(PROD_HEAD, n, h, wd), HEAD_STRIDE, x)
        with_dx_sums = dbias_x is not None
    with _timed("hbm:head_bwd", n * h * wd * (128.0 + 8.0 + 128.0)):      # x, dlogits in; dx out
        _chk(_lib.load().ub_head_bwd(_ptr(x), _ptr(w), _ptr(dlogits), _ptr(dx), _ptr(part), int(with_dx_sums), n, h, wd,
                                      _stream()), "head_bwd")
    if own:
        reduce_partials(part, part.shape[0], HEAD_STRIDE, 128, dw, True)
        reduce_partials(part, part.shape[0], HEAD_STRIDE, 2, db, True, col=128)
        if dbias_x is not None:
            reduce_partials(part, part.shape[0], HEAD_STRIDE, 64, dbias_x, True, col=130)


LOSS_BCE, LOSS_SOFTMAX_CE = 0, 1


def loss_workspace(device):
    nbytes = _lib.load().ub_loss_workspace_bytes()
    return torch.zeros(nbytes // 4, dtype=torch.float32, device=device)


def loss_fwd_bwd(kind, logits, label, weight, dlogits, loss_out, workspace, grad_scale):
    """logits fp32 [B,2,Hp,Wp]; label [B,L,L] uint8|int64; weight [B,L,L] fp32 (trainer.py:60-75)."""
    _need_cuda(logits, label, weight, loss_out, workspace)
    b, _, hp, wp = logits.shape
    l = label.shape[-1]
    lb = {torch.uint8: 1, torch.int64: 8}[label.dtype]
    # 2 logits + label + weight read, 2 dlogits written per pixel (fp32 logits: 21 B with uint8 labels)
    with _timed("hbm:loss", b * l * l * (8.0 + lb + 4.0 + (8.0 if dlogits is not None else 0.0))):
        _chk(_lib.load().ub_loss_fwd_bwd(kind, _ptr(logits), _ptr(label), lb, _ptr(weight), _ptr(dlogits), _ptr(loss_out),
                                          _ptr(workspace), b, hp, wp, l, float(grad_scale), _stream()), "loss_fwd_bwd")


def colsum(g, out=None, part=None):
    """Column sums over pixels of the bf16 NHWC view g (dense in n,y,x up to the pixel stride).  part:
    [partial_rows(PROD_COLSUM, n, h, w, c), c] receives the partial rows; out (fp32 [C]): reduced here, out += sums."""
    n, h, w, c = g.shape
    ld = g.stride(2)
    if g.stride(1) != w * ld or (n > 1 and g.stride(0) != h * w * ld):
        raise RuntimeError("colsum needs a view that is dense over pixels")
    own = part is None
    if own:
        part = _part(partial_rows(PROD_COLSUM, n, h, w, c), c, g)
    with _timed("hbm:colsum", n * h * w * c * 2.0):
        _chk(_lib.load().ub_colsum(_ptr(g), _ptr(part), n * h * w, c, ld, _stream()), "colsum")
    if own:
        reduce_partials(part, part.shape[0], c, c, out, True)


def sgd_momentum(p, g, buf, lr, momentum, first_step):
    with _timed("hbm:sgd", p.numel() * 20.0):                              # 20 B per parameter
        _chk(_lib.load().ub_sgd_momentum(_ptr(p), _ptr(g), _ptr(buf), p.numel(), float(lr), float(momentum),
                                          int(first_step), _stream()), "sgd_momentum")


def naive_conv3x3(x, w_fwd, bias, relu):
    n, h, w, cin = x.shape
    cout = w_fwd.shape[0]
    y = torch.empty((n, h - 2, w - 2, cout), dtype=torch.float32, device=x.device)
    _chk(_lib.load_diag().ub_naive_conv3x3(_ptr(x), _ptr(w_fwd), _ptr(bias), _ptr(y), n, h, w, cin, cout, int(relu),
                                       _stream()), "naive_conv3x3")
    return y


def probe_run(cap_rows=128, which=0):
    buf = (C.c_float * (cap_rows * 6))(*([-1.0] * (cap_rows * 6)))
    _chk(_lib.load_diag().ub_probe_run(which, buf, cap_rows * 6), "probe")
    rows = []
    for i in range(cap_rows):
        r = list(buf[i * 6:(i + 1) * 6])
        if r[0] < 0:
            break
        rows.append(r)
    return rows
