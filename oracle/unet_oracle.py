"""CPU oracle for the U-Net hot path -- TEST INFRASTRUCTURE ONLY.

A plain fp32 restatement (functional PyTorch on the CPU, no nn.Module, no custom kernels) of
the reference's training / inference path.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this file; the product path
(dl-unet_b200/) never does and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against the reference itself: oracle/validate_against_reference.py imports
/root/reference/{network,trainer}.py in the build container and checks bit-equality of weights,
logits, loss, all 46 gradients and the SGD update, then writes tests/golden/*.npz which travel to
the GPU box (where /root/reference does not exist).

Each function cites the reference lines it restates (paths relative to the reference repo).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# (name, kind, cin, cout) in construction order -- network.py:23-58
LAYERS = [
    ("conv11c", "conv", 1, 64), ("conv12c", "conv", 64, 64),
    ("conv21c", "conv", 64, 128), ("conv22c", "conv", 128, 128),
    ("conv31c", "conv", 128, 256), ("conv32c", "conv", 256, 256),
    ("conv41c", "conv", 256, 512), ("conv42c", "conv", 512, 512),
    ("conv51c", "conv", 512, 1024), ("conv52c", "conv", 1024, 1024),
    ("upconv4", "up", 1024, 512),
    ("conv41e", "conv", 1024, 512), ("conv42e", "conv", 512, 512),
    ("upconv3", "up", 512, 256),
    ("conv31e", "conv", 512, 256), ("conv32e", "conv", 256, 256),
    ("upconv2", "up", 256, 128),
    ("conv21e", "conv", 256, 128), ("conv22e", "conv", 128, 128),
    ("upconv1", "up", 128, 64),
    ("conv11e", "conv", 128, 64), ("conv12e", "conv", 64, 64),
    ("finalconv", "head", 64, 2),
]
PARAM_NAMES = [f"{n}.{p}" for (n, _, _, _) in LAYERS for p in ("weight", "bias")]


def init_std(name):
    """Gaussian std of each weight -- network.py:70-105.  Note the operator precedence the
    reference actually has: 2 / sqrt(N), not sqrt(2 / N); conv11c uses sqrt(2)."""
    spec = {n: (k, ci, co) for (n, k, ci, co) in LAYERS}
    if name == "conv11c":
        return 2 ** 0.5                                   # network.py:70
    kind, cin, cout = spec[name]
    if name in ("conv41e", "conv31e", "conv21e", "conv11e"):
        # N = Cskip*3^2 + Cup*2^2 with Cskip = Cup = cin/2  -- network.py:87,92,97,102
        return 2 / (cin // 2 * 9.0 + cin // 2 * 4.0) ** 0.5
    # every other layer (up-convs and the 1x1 head included): N = cin * 3^2 -- network.py:71-105
    return 2 / (cin * 9.0) ** 0.5


def init_state_dict(seed):
    """Reproduce `torch.manual_seed(seed); Unet().state_dict()` -- network.py:20-105.
    RNG consumption order: the 23 default nn.Conv2d / nn.ConvTranspose2d constructors in
    definition order (kaiming-uniform weight then uniform bias each), then the 23 normal_ draws
    in the order of network.py:70-105, which is the definition order as well."""
    torch.manual_seed(seed)
    sd = {}
    for (name, kind, cin, cout) in LAYERS:                # network.py:23-58
        k = {"conv": 3, "up": 2, "head": 1}[kind]
        shape = (cin, cout, k, k) if kind == "up" else (cout, cin, k, k)
        w = torch.empty(shape)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))           # nn.Conv2d.reset_parameters
        fan_in = shape[1] * k * k                                    # torch's fan_in uses dim 1 for both kinds
        bound = 1 / math.sqrt(fan_in)
        b = torch.empty(cout)
        torch.nn.init.uniform_(b, -bound, bound)
        sd[name + ".weight"], sd[name + ".bias"] = w, b
    for (name, _, _, _) in LAYERS:                        # network.py:70-105
        sd[name + ".weight"] = torch.empty_like(sd[name + ".weight"]).normal_(mean=0, std=init_std(name))
    return sd


def crop_offset(ha, hb):
    """c = int((HA - HB) * 0.5) -- network.py:124-125 (truncation toward zero, sign kept)."""
    return int((ha - hb) * 0.5)


def crop_and_concat(a, b):
    """network.py:108-127: negative c zero-pads the skip tensor, positive c centre-crops it;
    an odd size difference makes torch.cat raise RuntimeError, as in the reference."""
    c = crop_offset(a.shape[2], b.shape[2])
    a = F.pad(a, (-c, -c, -c, -c))
    return torch.cat((a, b), 1)


def forward(sd, t, taps=None):
    """Unet.forward -- network.py:129-192.  `taps`, if a dict, receives named intermediates."""
    def conv(name, t, relu=True):
        t = F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"])
        t = F.relu(t) if relu else t
        if taps is not None:
            taps[name] = t
        return t

    def up(name, t):
        t = F.conv_transpose2d(t, sd[name + ".weight"], sd[name + ".bias"], stride=2)
        if taps is not None:
            taps[name] = t
        return t

    skips = []
    for lvl in "1234":                                   # network.py:131-152
        t = conv(f"conv{lvl}1c", t)
        t = conv(f"conv{lvl}2c", t)
        t = F.max_pool2d(t, kernel_size=2, stride=2)
        if taps is not None:
            taps[f"pool{lvl}"] = t
        skips.append(t)                                  # the POOLED tensor is the skip (quirk 1)
    t = conv("conv51c", t)                               # network.py:155-156
    t = conv("conv52c", t)
    for lvl in "4321":                                   # network.py:159-188
        t = up(f"upconv{lvl}", t)
        t = crop_and_concat(skips[int(lvl) - 1], t)
        t = conv(f"conv{lvl}1e", t)
        t = conv(f"conv{lvl}2e", t)
    return conv("finalconv", t, relu=False)              # network.py:190


def output_size(s):
    """Spatial size of the logits for an s x s input, or None if the reference would raise
    (odd pre-pool size is silently floored by max_pool2d; an odd skip/up difference raises)."""
    sizes = []
    for _ in range(4):
        s = s - 4
        if s <= 0:
            return None
        s = s // 2
        sizes.append(s)
    s = s - 4
    if s <= 0:
        return None
    for lvl in (3, 2, 1, 0):
        s = 2 * s
        if (sizes[lvl] - s) % 2:
            return None
        s = s - 4
        if s <= 0:
            return None
    return s


def center_crop_preds(preds, label_size):
    """trainer.py:60-61 / tester.py:29-30."""
    pad = int((preds.shape[-1] - label_size) / 2)
    return preds[:, :, pad:label_size + pad, pad:label_size + pad]


def loss_bce(preds, labels, weight):
    """trainer.py:63-75 with the per-pixel weight broadcast (identical to the reference at B = 1;
    SURVEY.md 0.1-3): mean over B*2*H*W of w[b,h,w] * BCEWithLogits(z_c, onehot_c)."""
    ll = torch.empty_like(preds)
    ll[:, 0] = 1 - labels[:, 0]                          # trainer.py:64
    ll[:, 1] = labels[:, 0]                              # trainer.py:65
    return F.binary_cross_entropy_with_logits(preds, ll, weight=weight.unsqueeze(1))


def loss_softmax_ce(preds, labels, weight):
    """BASELINE config 5: mean over B*H*W of w * (logsumexp(z) - z_label)."""
    return (F.cross_entropy(preds, labels[:, 0].long(), reduction="none") * weight).mean()


def train_step(sd, mom, images, labels, weight, lr=1e-4, momentum=0.99, kind="bce"):
    """One iteration of the hot loop, trainer.py:54-78: zero_grad, forward, crop, loss, backward,
    SGD-momentum step (optim.SGD(lr=1e-4, momentum=0.99), trainer.py:30).  Updates sd / mom in
    place; returns (loss, logits, grads)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits = forward(params, images)
    preds = center_crop_preds(logits, labels.shape[-1])
    loss = loss_bce(preds, labels, weight) if kind == "bce" else loss_softmax_ce(preds, labels, weight)
    grads = dict(zip(params.keys(), torch.autograd.grad(loss, list(params.values()))))
    with torch.no_grad():
        for k in sd:
            g = grads[k]
            if k not in mom:
                mom[k] = g.clone()                       # first step: buf = g
            else:
                mom[k].mul_(momentum).add_(g)            # buf = m*buf + g
            sd[k].add_(mom[k], alpha=-lr)                # p += (-lr)*buf, as optim.SGD does
    return loss.detach(), logits.detach(), grads


def predict_mask(sd, image, label_size):
    """tester.py:27-30: forward, centre-crop, argmax over the 2 channels (ties -> class 0)."""
    with torch.no_grad():
        return center_crop_preds(forward(sd, image), label_size).argmax(dim=1)


def pool_with_indices(t):
    """F.max_pool2d(t, 2, 2, return_indices=True): ATen's first-maximum tie rule, flat h*W+w."""
    return F.max_pool2d(t, 2, 2, return_indices=True)


# ------------------------------------------------------------------ synthetic workload
def synthetic_batch(batch, size=572, seed=0, w0=10.0, sigma=5.0):
    """Synthetic tiles for the benchmark / parity tests: uniform image,
    blob labels, border-emphasis weight map 1 + w0*exp(-d^2/(2 sigma^2)) built from a cheap
    distance-to-foreground proxy (the real map generators are CPU code outside this path)."""
    out = output_size(size)
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(batch, 1, size, size, generator=g)
    yy, xx = torch.meshgrid(torch.arange(out), torch.arange(out), indexing="ij")
    labels = torch.zeros(batch, 1, out, out, dtype=torch.int64)
    dist = torch.full((batch, out, out), 1e9)
    for b in range(batch):
        n_blobs = max(2, out // 24)
        for _ in range(n_blobs):
            cy, cx = (torch.rand(2, generator=g) * out).tolist()
            ry, rx = (4 + torch.rand(2, generator=g) * max(4.0, out / 6)).tolist()
            d = torch.sqrt(((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2)
            labels[b, 0][d <= 1] = 1
            dist[b] = torch.minimum(dist[b], (d - 1).clamp_min(0) * min(ry, rx))
    weight = 1.0 + w0 * torch.exp(-(dist ** 2) / (2 * sigma ** 2)) * (1 - labels[:, 0]).float()
    return images, labels, weight.float()


def to_numpy_tree(d):
    return {k: v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v) for k, v in d.items()}
