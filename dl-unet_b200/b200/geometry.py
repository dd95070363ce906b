"""Shape arithmetic of the valid-convolution U-Net (pure Python, no torch): layer table,
per-level sizes, crop_and_concat offsets, tile planning.  Mirrors network.py:23-58 (layers),
network.py:108-127 (crop rule) and functions.py:121-146 (input_size_compute)."""

# channels of encoder level l = 1..5
LEVEL_C = {1: 64, 2: 128, 3: 256, 4: 512, 5: 1024}

# (name, kind, cin, cout) in definition order (= state_dict order) -- network.py:23-58
LAYERS = [
    ("conv11c", "conv1", 1, 64), ("conv12c", "conv", 64, 64),
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
LAYER_BY_NAME = {n: (k, ci, co) for (n, k, ci, co) in LAYERS}


def weight_shape(name):
    kind, cin, cout = LAYER_BY_NAME[name]
    if kind == "up":
        return (cin, cout, 2, 2)
    if kind == "head":
        return (cout, cin, 1, 1)
    return (cout, cin, 3, 3)


def crop_offset(ha, hb):
    """c = int((HA - HB) * 0.5): network.py:124-125 -- truncation toward zero, sign kept.
    c < 0: the skip tensor is zero-padded by -c per side; c > 0: centre-cropped by c."""
    return int((ha - hb) * 0.5)


class Geometry:
    """All spatial sizes for an S x S input.  Raises RuntimeError where the reference would
    (torch.cat size mismatch on an odd skip/up difference, network.py:127) or where a layer
    would see a non-positive size."""

    def __init__(self, size):
        s = int(size)
        self.S = s
        self.enc1, self.enc2, self.pool = {}, {}, {}
        for l in (1, 2, 3, 4):
            self.enc1[l] = s - 2
            self.enc2[l] = s - 4
            if s - 4 < 2:
                raise RuntimeError(f"input size {size} is too small for the U-Net (level {l})")
            s = (s - 4) // 2            # F.max_pool2d floors odd sizes
            self.pool[l] = s
        self.enc1[5], self.enc2[5] = s - 2, s - 4
        if s - 4 < 1:
            raise RuntimeError(f"input size {size} is too small for the U-Net (bottleneck)")
        s = s - 4
        self.up, self.dec1, self.dec2, self.coff = {}, {}, {}, {}
        for l in (4, 3, 2, 1):
            s = 2 * s
            self.up[l] = s
            c = crop_offset(self.pool[l], s)
            if self.pool[l] - 2 * c != s:
                # same failure the reference hits in torch.cat (network.py:127)
                raise RuntimeError(
                    f"Sizes of tensors must match except in dimension 1. Expected size {self.pool[l] - 2 * c} "
                    f"but got size {s} for tensor number 1 in the list. (crop_and_concat at level {l}, input {size})")
            self.coff[l] = c
            self.dec1[l], self.dec2[l] = s - 2, s - 4
            if s - 4 < 1:
                raise RuntimeError(f"input size {size} is too small for the U-Net (decoder level {l})")
            s = s - 4
        self.out = s


def input_size_compute(original_size):
    """functions.py:121-146: smallest valid (input, output) pair whose output covers original_size."""
    lowest = 20
    while True:
        inp = (((lowest * 2 + 4) * 2 + 4) * 2 + 4) * 2 + 4
        out = ((((lowest - 4) * 2 - 4) * 2 - 4) * 2 - 4) * 2 - 4
        if out >= original_size:
            return original_size, inp, out
        lowest += 2


def conv_flops_per_tile(size=572):
    """Dense forward FLOPs (2*M*N*K) per layer for one tile -- the SURVEY.md 2.1 table."""
    g = Geometry(size)
    out = {}
    for (name, kind, cin, cout) in LAYERS:
        if kind in ("conv", "conv1"):
            lvl, pos, side = int(name[4]), int(name[5]), name[6]
            h = (g.enc1 if pos == 1 else g.enc2)[lvl] if side == "c" else (g.dec1 if pos == 1 else g.dec2)[lvl]
            out[name] = 2.0 * h * h * cout * cin * 9
        elif kind == "up":
            lvl = int(name[-1])
            hin = g.up[lvl] // 2
            out[name] = 2.0 * hin * hin * cin * cout * 4
        else:
            out[name] = 2.0 * g.out * g.out * cin * cout
    return out
