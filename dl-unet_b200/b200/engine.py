"""Execution engine of the B200 U-Net path: owns the flat parameter / gradient arenas, the
packed bf16 weight caches and the per-shape activation workspaces, and issues the explicit
forward and backward kernel schedules (the replacement for the ATen dispatch + autograd tape
of network.py:129-192 / trainer.py:77).

Data layout in HBM
  * activations: NHWC bf16, one buffer per layer output; the decoder's concat buffers
    [B, Hu, Hu, 2*C] are written in place by the pool kernel (skip half, at the
    crop_and_concat offset; border pre-zeroed once) and the up-convolution epilogue (upper half),
    so network.py:126-127's pad + cat are never materialised.
  * parameters: one flat fp32 arena in state_dict order (the nn.Parameters are views of it),
    one flat fp32 gradient arena in BACKWARD-COMPLETION order (finalconv first) so gradient
    buckets become ready front to back, one flat momentum arena.
  * derived caches: bf16 K-major packed weights (forward and dgrad operand per layer),
    refreshed when the masters change.
"""
import torch

from . import ops
from .geometry import LAYERS, LAYER_BY_NAME, LEVEL_C, Geometry, weight_shape

# encoder levels whose 2x2 max pool runs inside the epilogue of conv{l}2c (the haloed-patch kernel: <= 256 channels);
# UB_FUSED_POOL_LEVELS=0 restores the separate pool kernel everywhere (experiments)
FUSED_POOL_LEVELS = int(__import__("os").environ.get("UB_FUSED_POOL_LEVELS", "2"))

# UB_STREAMK_INFERENCE=1: stream-K in inference too (+1-2 % throughput; predictions then depend on the batch size in the
# last bf16 bit of a few activations)
STREAMK_INFERENCE = int(__import__("os").environ.get("UB_STREAMK_INFERENCE", "0"))

PARAM_NAMES = [f"{n}.{p}" for (n, _, _, _) in LAYERS for p in ("weight", "bias")]
# order in which backward finishes each layer's gradients
BACKWARD_LAYER_ORDER = [n for (n, _, _, _) in reversed(LAYERS)]


def _align(n, a=64):
    return (n + a - 1) // a * a


class ParamStore:
    """Flat fp32 arenas for the 46 parameters, their gradients and momentum buffers."""

    def __init__(self, module):
        self.module = module
        self.device = None
        self.p_arena = self.g_arena = self.m_arena = None
        # one layout for parameters, gradients and momentum: BACKWARD-completion order (finalconv
        # first), each tensor padded to 64 elements, so (a) gradient buckets fill front to back and
        # (b) the optimizer is a single elementwise kernel over [0, total)
        self.numel, self.g_off = {}, {}
        for name in PARAM_NAMES:
            lname, pname = name.split(".")
            shape = weight_shape(lname) if pname == "weight" else (LAYER_BY_NAME[lname][2],)
            n = 1
            for d in shape:
                n *= d
            self.numel[name] = n
        goff = 0
        for lname in BACKWARD_LAYER_ORDER:
            for pname in ("weight", "bias"):
                name = f"{lname}.{pname}"
                self.g_off[name] = goff
                goff += _align(self.numel[name])
        self.p_off = self.g_off
        self.total = goff
        self.packs = {}
        self.pack_versions = None
        self.dirty = True
        self.first_step = True

    def params(self):
        return {name: self.module.get_parameter(name) for name in PARAM_NAMES}

    def adopt(self, device):
        """Make every nn.Parameter a view of the flat arena on `device` (idempotent).  Called at
        the top of each forward: `.to(device)` and `load_state_dict` may have re-pointed them."""
        ps = self.params()
        ok = self.p_arena is not None and self.p_arena.device == device
        if ok:
            base = self.p_arena.data_ptr()
            for name, p in ps.items():
                if p.data_ptr() != base + 4 * self.p_off[name] or p.dtype != torch.float32:
                    ok = False
                    break
        if ok:
            return
        for name, p in ps.items():
            if p.device != device:
                raise RuntimeError(f"parameter {name} is on {p.device}, input on {device}: call unet.to(device) first")
        arena = torch.zeros(self.total, dtype=torch.float32, device=device)
        for name, p in ps.items():
            view = arena[self.p_off[name]:self.p_off[name] + self.numel[name]].view(p.shape)
            view.copy_(p.detach().to(torch.float32))
            p.data = view
        self.p_arena = arena
        self.g_arena = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.m_arena = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.jobs = None
        self.arena_id = getattr(self, "arena_id", 0) + 1   # workspaces cache job tables that point into the arenas
        self.packs = {}
        self.device = device
        self.dirty = True
        self.first_step = True

    def p(self, name):
        return self.p_arena[self.p_off[name]:self.p_off[name] + self.numel[name]]

    def g(self, name):
        return self.g_arena[self.g_off[name]:self.g_off[name] + self.numel[name]]

    def grad_view(self, name):
        lname, pname = name.split(".")
        shape = weight_shape(lname) if pname == "weight" else (LAYER_BY_NAME[lname][2],)
        return self.g(name).view(shape)

    def versions(self):
        return tuple(p._version for p in self.params().values())

    def _layer_entries(self, slab_of=None, with_name=False):
        """Job-table rows of the 21 conv / up-conv layers (ops.JobTable).  slab_of(lname) -> (slab workspace, n_slabs);
        None: a pack-only table (the gradient columns are placeholders)."""
        entries = []
        for (lname, kind, cin, cout) in LAYERS:
            if kind not in ("conv", "up"):
                continue
            w = self.p(lname + ".weight")
            if lname not in self.packs:
                fshape = (cout, 9 * cin) if kind == "conv" else (4 * cout, cin)
                dshape = (cin, 9 * cout) if kind == "conv" else (cin, 4 * cout)
                self.packs[lname] = (torch.empty(fshape, dtype=torch.bfloat16, device=self.device),
                                     torch.empty(dshape, dtype=torch.bfloat16, device=self.device))
            slab, n_slabs = slab_of(lname) if slab_of is not None else (self.g(lname + ".weight"), 1)
            entries.append((w, self.packs[lname][0], self.packs[lname][1], slab, self.g(lname + ".weight"),
                            0 if kind == "conv" else 2, cout, cin, n_slabs) + ((lname,) if with_name else ()))
        return entries

    def _job_table(self):
        """One-launch pack over all 21 conv / up-conv layers; rebuilt with the arenas."""
        if getattr(self, "jobs", None) is None:
            self.jobs = ops.JobTable(self._layer_entries(), self.device)
        return self.jobs

    def refresh_packs(self):
        """Re-derive the bf16 GEMM operands (one launch for all layers) when the fp32 masters changed."""
        v = self.versions()
        if not self.dirty and v == self.pack_versions:
            return
        self._job_table().pack_all()
        self.pack_versions = v
        self.dirty = False


class Workspace:
    """Activation (and, for training, gradient) buffers for one (batch, size) shape."""

    def __init__(self, batch, size, device, train):
        g = Geometry(size)
        self.geom, self.B, self.S, self.train = g, batch, size, train
        self.busy = False
        bf = dict(dtype=torch.bfloat16, device=device)
        B = batch
        a = {}
        for l in (1, 2, 3, 4):
            C = LEVEL_C[l]
            a[f"e{l}1"] = torch.empty(B, g.enc1[l], g.enc1[l], C, **bf)
            # levels 1-2 pool inside the conv epilogue: the full-resolution conv{l}2c output never exists in HBM
            # (backward needs only the pooled values, the argmax codes and the layer's INPUT)
            a[f"e{l}2"] = torch.empty(B, g.enc2[l], g.enc2[l], C, **bf) if l > FUSED_POOL_LEVELS else None
            a[f"p{l}"] = torch.empty(B, g.pool[l], g.pool[l], C, **bf)
            a[f"i{l}"] = torch.empty(B, g.pool[l], g.pool[l], C, dtype=torch.uint8, device=device)
            # zero once: the padded border of the skip half is never written afterwards
            a[f"cat{l}"] = torch.zeros(B, g.up[l], g.up[l], 2 * C, **bf)
            a[f"d{l}1"] = torch.empty(B, g.dec1[l], g.dec1[l], C, **bf)
            a[f"d{l}2"] = torch.empty(B, g.dec2[l], g.dec2[l], C, **bf)
        a["e51"] = torch.empty(B, g.enc1[5], g.enc1[5], 1024, **bf)
        a["e52"] = torch.empty(B, g.enc2[5], g.enc2[5], 1024, **bf)
        self.act = a
        self.grad = {}
        if train:
            for k, t in a.items():
                if k[0] != "i" and t is not None:
                    self.grad[k] = torch.empty_like(t)
            for l in range(1, FUSED_POOL_LEVELS + 1):
                self.grad[f"e{l}2"] = torch.empty(B, g.enc2[l], g.enc2[l], LEVEL_C[l], **bf)
        self.image = None
        self.slabs = None   # weight-gradient slab workspace + unpack job table, built on the first backward

    def wgrad_input_size(self, lname):
        """Spatial size of the INPUT of a conv / up-conv layer (its weight gradient reduces over the output pixels)."""
        g = self.geom
        if lname.startswith("upconv"):
            return g.up[int(lname[-1])] // 2
        lvl, pos, side = int(lname[4]), int(lname[5]), lname[6]
        out = ((g.enc1 if pos == 1 else g.enc2) if side == "c" else (g.dec1 if pos == 1 else g.dec2))[lvl]
        return out + 2

    def bias_plan(self):
        """Where every bias gradient (and the two small weight gradients) of the backward schedule comes from:
        name -> (producer, ub_partial_rows arguments, row stride, column offset in the row, length, buffer key).
        Mirrors Engine.backward."""
        g, B = self.geom, self.B
        plan = {}
        head = ("head", (ops.PROD_HEAD, B, g.out, g.out), ops.HEAD_STRIDE)
        plan["finalconv.weight"] = head + (0, 130, "head")          # weight [2][64] + bias [2]: contiguous in the arena
        plan["conv12e.bias"] = head + (130, 64, "head")
        for l in (1, 2, 3, 4):
            C = LEVEL_C[l]
            if l > 1:   # the up-convolution below this level consumes d{l}2: its dgrad sums the columns of its dx
                plan[f"conv{l}2e.bias"] = ("updgrad", (ops.PROD_UPDGRAD, B, g.dec2[l], g.dec2[l], C), C, 0, C, None)
            # conv{l}2e's dgrad sums its dx = the output gradient of conv{l}1e; conv{l}1e's dgrad that of upconv{l}
            plan[f"conv{l}1e.bias"] = ("dgrad", (ops.PROD_DGRAD, B, g.dec1[l], g.dec1[l], C, C, 0), C, 0, C, None)
            plan[f"upconv{l}.bias"] = ("dgrad", (ops.PROD_DGRAD, B, g.up[l], g.up[l], 2 * C, C, C), C, 0, C, None)
            plan[f"conv{l}2c.bias"] = ("pool", (ops.PROD_POOL, B, g.enc2[l], g.enc2[l], C), C, 0, C, None)
            if l > 1:
                plan[f"conv{l}1c.bias"] = ("dgrad", (ops.PROD_DGRAD, B, g.enc1[l], g.enc1[l], C, C, 0), C, 0, C, None)
        plan["conv52c.bias"] = ("updgrad", (ops.PROD_UPDGRAD, B, g.enc2[5], g.enc2[5], 1024), 1024, 0, 1024, None)
        plan["conv51c.bias"] = ("dgrad", (ops.PROD_DGRAD, B, g.enc1[5], g.enc1[5], 1024, 1024, 0), 1024, 0, 1024, None)
        plan["conv11c.weight"] = ("conv1", (ops.PROD_CONV1, B, self.S, self.S), ops.CONV1_STRIDE, 0, 640, None)
        return plan

    def grad_slabs(self, store):
        """Deterministic reductions: the wgrad kernel of every layer stores `n` partial gradients (slabs) side by side,
        every bias-gradient producer one row of partial sums per CTA; slab and row counts depend on the layer shape at
        this batch, so the workspaces and the unpack job table (which sums them in fixed order) live here.
        -> {"view": layer -> slab workspace, "part": parameter -> partial-row buffer, "table", "range": layer -> jobs}"""
        if self.slabs is None or self.slabs["arena_id"] != store.arena_id:
            dev = store.device
            counts, off, total = {}, {}, 0
            for (lname, kind, cin, cout) in LAYERS:
                if kind not in ("conv", "up"):
                    continue
                h = self.wgrad_input_size(lname)
                counts[lname] = ops.wgrad_slabs(ops.PACK_CONV_FWD if kind == "conv" else ops.PACK_UP_FWD, self.B, h, h, cin, cout)
                off[lname] = total
                total += counts[lname] * store.numel[lname + ".weight"]
            arena = torch.empty(total, dtype=torch.float32, device=dev)
            view = lambda l: arena[off[l]:off[l] + counts[l] * store.numel[l + ".weight"]]
            # partial-row buffers of the bias producers
            plan, bufs, part = self.bias_plan(), {}, {}
            for name, (prod, qargs, stride, col, length, key) in plan.items():
                key = key or name
                if key not in bufs:
                    bufs[key] = torch.empty(ops.partial_rows(*qargs), stride, dtype=torch.float32, device=dev)
                part[name] = bufs[key]
            weights = {e[-1]: e[:-1] for e in store._layer_entries(lambda l: (view(l), counts[l]), with_name=True)}
            dummy = store.g_arena[:1]
            entries, ranges = [], {}
            for (lname, kind, cin, cout) in LAYERS:
                first = len(entries)
                if lname in weights:
                    entries.append(weights[lname])
                for pname in ("weight", "bias"):
                    name = f"{lname}.{pname}"
                    if name in plan:
                        prod, qargs, stride, col, length, key = plan[name]
                        buf = part[name]
                        dst = store.g_arena[store.g_off[name]:store.g_off[name] + length]   # may span weight + bias
                        entries.append((dummy, dummy, dummy, buf.view(-1)[col:], dst, 3, length, stride, buf.shape[0]))
                ranges[lname] = (first, len(entries))
            table = ops.JobTable(entries, dev)
            self.slabs = {"arena": arena, "view": view, "counts": counts, "part": part, "bufs": bufs, "table": table,
                          "range": ranges, "arena_id": store.arena_id}
        return self.slabs


class Engine:
    def __init__(self, module):
        self.store = ParamStore(module)
        self.workspaces = {}
        self.grad_ready_hook = None  # callable(start, end): gradient arena slice [start, end) is final
        self.grad_done_hook = None   # callable(): every gradient has been produced

    # ------------------------------------------------------------------ workspaces
    def workspace(self, batch, size, device, train):
        key = (batch, size, str(device), train)
        pool = self.workspaces.setdefault(key, [])
        for ws in pool:
            if not ws.busy:
                return ws
        ws = Workspace(batch, size, device, train)
        pool.append(ws)
        return ws

    # ------------------------------------------------------------------ forward
    def forward(self, image, train, argmax_out=None, want_logits=True):
        """image: fp32 [B,1,S,S] CUDA.  Returns (logits fp32 [B,2,So,So] or None, workspace).

        Inference (train=False) runs with stream-K off: whole work units keep one K order whatever the batch size, so
        the prediction for a tile does not depend on what else is in its batch (tests/test_gpu_dropin.py: the tile
        route equals per-tile predict() bit for bit); training takes the 1-2 % stream-K gives -- its split points
        depend on the batch shape only, so it stays bit-reproducible from run to run."""
        prev = ops.get_option(ops.OPT_STREAM_K, 1)
        if not train and not STREAMK_INFERENCE and prev != 0:
            ops.set_option(ops.OPT_STREAM_K, 0)
            try:
                return self._forward(image, train, argmax_out, want_logits)
            finally:
                ops.set_option(ops.OPT_STREAM_K, prev)
        return self._forward(image, train, argmax_out, want_logits)

    def _forward(self, image, train, argmax_out, want_logits):
        if image.dim() != 4 or image.shape[1] != 1 or image.shape[2] != image.shape[3]:
            raise RuntimeError(f"expected a [B,1,S,S] image batch, got {tuple(image.shape)}")
        if not image.is_cuda:
            raise RuntimeError("the B200 U-Net path has no CPU implementation: move the module and the input to a CUDA device")
        st = self.store
        st.adopt(image.device)
        st.refresh_packs()
        image = image.contiguous().float()
        B, _, S, _ = image.shape
        ws = self.workspace(B, S, image.device, train)
        g, a = ws.geom, ws.act
        ws.image = image
        P, pk = st.p, st.packs

        ops.conv1_fwd(image, P("conv11c.weight"), P("conv11c.bias"), a["e11"])
        x = None
        for l in (1, 2, 3, 4):
            C = LEVEL_C[l]
            if l > 1:
                ops.conv3x3_fwd(x, pk[f"conv{l}1c"][0], P(f"conv{l}1c.bias"), a[f"e{l}1"])
            if l <= FUSED_POOL_LEVELS:
                # conv{l}2c + max pool + argmax codes + skip copy in one launch (network.py:132-133, 138-139)
                ops.conv3x3_fwd_pool(a[f"e{l}1"], pk[f"conv{l}2c"][0], P(f"conv{l}2c.bias"), C, None, a[f"p{l}"], a[f"i{l}"],
                                     a[f"cat{l}"][..., :C], g.coff[l])
            else:
                ops.conv3x3_fwd(a[f"e{l}1"], pk[f"conv{l}2c"][0], P(f"conv{l}2c.bias"), a[f"e{l}2"])
                ops.pool2x2_fwd(a[f"e{l}2"], a[f"p{l}"], a[f"i{l}"], a[f"cat{l}"][..., :C], g.coff[l])
            x = a[f"p{l}"]
        ops.conv3x3_fwd(x, pk["conv51c"][0], P("conv51c.bias"), a["e51"])
        ops.conv3x3_fwd(a["e51"], pk["conv52c"][0], P("conv52c.bias"), a["e52"])
        x = a["e52"]
        for l in (4, 3, 2, 1):
            C = LEVEL_C[l]
            ops.upconv2x2_fwd(x, pk[f"upconv{l}"][0], P(f"upconv{l}.bias"), a[f"cat{l}"][..., C:])
            ops.conv3x3_fwd(a[f"cat{l}"], pk[f"conv{l}1e"][0], P(f"conv{l}1e.bias"), a[f"d{l}1"])
            if l > 1:
                ops.conv3x3_fwd(a[f"d{l}1"], pk[f"conv{l}2e"][0], P(f"conv{l}2e.bias"), a[f"d{l}2"])
            x = a[f"d{l}2"]
        logits = None
        if want_logits:
            logits = torch.empty(B, 2, g.out, g.out, dtype=torch.float32, device=image.device)
        # conv12e + finalconv in one launch: the head reads the fp32 accumulators (one bf16 rounding less on the
        # logits); inference never writes the 64-channel activation to HBM
        ops.conv3x3_fwd_head(a["d11"], pk["conv12e"][0], P("conv12e.bias"), a["d12"] if train else None,
                             P("finalconv.weight"), P("finalconv.bias"), logits, argmax_out)
        return logits, ws

    # ------------------------------------------------------------------ backward
    def _ready(self, lname):
        if self.grad_ready_hook is not None:
            st = self.store
            s = st.g_off[lname + ".weight"]
            e = st.g_off[lname + ".bias"] + _align(st.numel[lname + ".bias"])
            self.grad_ready_hook(s, e)

    def _finish(self, lname):
        """Every gradient of `lname` has been produced as slabs / partial rows.  Single device: they are all reduced by
        one launch at the end of backward.  Data parallel: the layer joins the pending set, which the all-reducer has
        reduced (flush_pending_unpack, one launch per bucket) right before it sends the bucket."""
        if self.grad_ready_hook is not None:
            self._pending.append(lname)
        self._ready(lname)

    def flush_pending_unpack(self):
        """Slabs / partial rows of the layers finished since the last flush -> final gradients, one launch.  Backward
        finishes layers in reverse definition order, so their jobs are one contiguous range of the table."""
        if self._pending:
            r = self._slabs["range"]
            i = min(r[l][0] for l in self._pending)
            j = max(r[l][1] for l in self._pending)
            self._slabs["table"].unpack_layer(i, self._accumulate, j)
            self._pending = []

    def _conv_bwd(self, lname, x, dy, dx, relu_src, bias_done=False, dx_bias=None, dx_bias_c0=0):
        """wgrad + bias grad (+ dgrad when dx is not None) of one 3x3 conv layer.  bias_done: the kernel that
        produced dy (pool / head backward, or the dgrad of the layer after this one) already left the partial rows of
        its column sums (this layer's bias gradient).  dx_bias: name of the parameter that receives the column sums of
        dx (channels >= dx_bias_c0) -- the bias gradient of the layer that produced x, fused into this layer's dgrad."""
        st = self.store
        ops.conv3x3_wgrad(x, dy, self._slabs["view"](lname))
        if not bias_done:
            ops.colsum(dy, part=self._slabs["part"][lname + ".bias"])
        self._finish(lname)
        if dx is not None:
            ops.conv3x3_dgrad(dy, st.packs[lname][1], dx, relu_src, dbias_c0=dx_bias_c0,
                              dbias_part=self._slabs["part"][dx_bias] if dx_bias is not None else None)

    def backward(self, ws, dlogits):
        """dlogits: fp32 [B,2,So,So] (d loss / d logits).  Fills the gradient arena; if every
        param.grad still is its arena view (no zero_grad since the last backward) the new
        gradients are accumulated on top, as autograd would.  Nothing is zeroed and nothing is accumulated with
        atomics: every kernel stores slabs / partial rows, the unpack job table sums them in fixed order."""
        st = self.store
        a, gr, g = ws.act, ws.grad, ws.geom
        self._accumulate = all(
            p.grad is not None and p.grad.data_ptr() == st.grad_view(n).data_ptr() for n, p in st.params().items())
        self._slabs = ws.grad_slabs(st)
        self._pending = []
        part = self._slabs["part"]
        P, pk = st.p, st.packs
        dlogits = dlogits.contiguous().float()

        ops.head_bwd(a["d12"], P("finalconv.weight"), dlogits, gr["d12"], part=part["finalconv.weight"])
        self._finish("finalconv")
        dy = gr["d12"]
        for l in (1, 2, 3, 4):
            C = LEVEL_C[l]
            # each dgrad also produces the bias gradient of the layer whose output gradient it computes
            # conv{l}2e's own bias gradient came with the kernel that produced dy: head_bwd (l = 1), upconv{l-1}'s dgrad
            self._conv_bwd(f"conv{l}2e", a[f"d{l}1"], dy, gr[f"d{l}1"], a[f"d{l}1"], bias_done=True,
                           dx_bias=f"conv{l}1e.bias")
            self._conv_bwd(f"conv{l}1e", a[f"cat{l}"], gr[f"d{l}1"], gr[f"cat{l}"], None, bias_done=True,
                           dx_bias=f"upconv{l}.bias", dx_bias_c0=C)
            # up-convolution: its output is the upper channel half of the concat buffer
            up_name = f"upconv{l}"
            src = a[f"d{l + 1}2"] if l < 4 else a["e52"]
            dsrc = gr[f"d{l + 1}2"] if l < 4 else gr["e52"]
            dyu = gr[f"cat{l}"][..., C:]
            ops.upconv2x2_wgrad(src, dyu, self._slabs["view"](up_name))
            self._finish(up_name)                    # its bias gradient came with conv{l}1e's dgrad
            # ... and its dgrad also leaves the bias gradient of the conv that produced its input (conv{l+1}2e / conv52c)
            ops.upconv2x2_dgrad(dyu, pk[up_name][1], dsrc, src,
                                dbias_part=part[f"conv{l + 1}2e.bias" if l < 4 else "conv52c.bias"])
            dy = dsrc
        # bottleneck
        self._conv_bwd("conv52c", a["e51"], gr["e52"], gr["e51"], a["e51"], bias_done=True, dx_bias="conv51c.bias")
        self._conv_bwd("conv51c", a["p4"], gr["e51"], gr["p4"], None, bias_done=True)
        for l in (4, 3, 2, 1):
            C = LEVEL_C[l]
            H = g.enc2[l]
            if H & 1:
                gr[f"e{l}2"].zero_()  # floor-mode pool: the last row / column gets no gradient
            ops.pool2x2_bwd(gr[f"p{l}"], gr[f"cat{l}"][..., :C], g.coff[l], a[f"i{l}"], a[f"p{l}"], gr[f"e{l}2"],
                            dbias_part=part[f"conv{l}2c.bias"])
            self._conv_bwd(f"conv{l}2c", a[f"e{l}1"], gr[f"e{l}2"], gr[f"e{l}1"], a[f"e{l}1"], bias_done=True,
                           dx_bias=f"conv{l}1c.bias" if l > 1 else None)     # conv11c: conv1_wgrad sums its own bias
            if l > 1:
                self._conv_bwd(f"conv{l}1c", a[f"p{l - 1}"], gr[f"e{l}1"], gr[f"p{l - 1}"], None, bias_done=True)
            else:
                ops.conv1_wgrad(ws.image, gr["e11"], part=part["conv11c.weight"])
                self._finish("conv11c")
        if self.grad_ready_hook is None:
            self._slabs["table"].unpack_all(self._accumulate)   # single device: all 46 gradients in one launch
        if self.grad_done_hook is not None:
            self.grad_done_hook()

    def assign_grads(self, accumulate_existing=True):
        """Expose the gradient arena through param.grad (views; no copies)."""
        st = self.store
        for name, p in st.params().items():
            v = st.grad_view(name)
            if p.grad is None or not accumulate_existing:
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                p.grad.add_(v)
            # else: p.grad already is this view and was overwritten -> caller zero_grad()s per step


class _Lease:
    """Marks a workspace busy for as long as an autograd graph may still need its activations."""

    def __init__(self, ws):
        self.ws = ws
        ws.busy = True

    def release(self):
        if self.ws is not None:
            self.ws.busy = False
            self.ws = None

    def __del__(self):
        self.release()


class _UnetFn(torch.autograd.Function):
    """Whole-network autograd node: forward = Engine.forward, backward = Engine.backward.  The 46
    parameter gradients are delivered through param.grad views of the flat arena (so a fused
    optimizer and bucketed all-reduce can work on contiguous memory) and not through autograd's
    return values."""

    @staticmethod
    def forward(ctx, engine, image, anchor):
        logits, ws = engine.forward(image, train=True)
        ctx.engine, ctx.ws = engine, ws
        ctx.lease = _Lease(ws)  # the workspace is released when backward ran or the graph died
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        engine, ws = ctx.engine, ctx.ws
        try:
            engine.backward(ws, dlogits)
            engine.assign_grads()
        finally:
            ctx.lease.release()
        return None, None, None


def unet_forward(engine, image):
    """Autograd-aware forward used by network.Unet.forward."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in engine.store.module.parameters()):
        if image.requires_grad:
            raise RuntimeError("gradients w.r.t. the input image are not part of the reference path (network.py:131)")
        # `anchor` makes the output require grad so that autograd calls our backward
        anchor = next(p for p in engine.store.module.parameters() if p.requires_grad)
        return _UnetFn.apply(engine, image, anchor)
    logits, _ = engine.forward(image, train=False)
    return logits
