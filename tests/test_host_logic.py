"""CPU: host-side logic of the drop-in (geometry, module surface, error behaviour, C ABI exports)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import unet_oracle as O
from b200 import lib
from b200.dist import bucket_bounds, partition
from b200.geometry import Geometry, conv_flops_per_tile, crop_offset, input_size_compute

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_geometry_572():
    g = Geometry(572)
    assert g.out == 388
    assert g.pool == {1: 284, 2: 140, 3: 68, 4: 32}
    assert g.up == {4: 56, 3: 104, 2: 200, 1: 392}
    assert g.coff == {4: -12, 3: -18, 2: -30, 1: -54}        # zero padding of the pooled skip (quirk 1)


def test_geometry_crop_branch_and_errors():
    g = Geometry(188)
    assert g.out == 4 and g.coff == {4: 0, 3: 6, 2: 18, 1: 42}  # positive = centre crop
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
        Geometry(204)
    with pytest.raises(RuntimeError):
        Geometry(100)
    for s in range(188, 700, 2):
        ours = None
        try:
            ours = Geometry(s).out
        except RuntimeError:
            pass
        assert ours == O.output_size(s), s


def test_crop_offset_matches_reference_rule():
    for ha in range(1, 40):
        for hb in range(1, 40):
            assert crop_offset(ha, hb) == int((ha - hb) * 0.5)


def test_input_size_compute():
    assert input_size_compute(388) == (388, 572, 388)
    assert input_size_compute(512)[1:] == (700, 516)
    assert input_size_compute(196)[1:] == (380, 196)
    assert input_size_compute(4096)[1] == 4284


def test_flop_table():
    f = conv_flops_per_tile(572)
    assert abs(sum(f.values()) / 1e9 - 300.86) < 0.05
    assert abs(f["conv41e"] / 1e9 - 27.519) < 0.01


def test_unet_module_surface():
    from network import Unet
    torch.manual_seed(0)
    net = Unet()
    sd = O.init_state_dict(0)
    assert list(net.state_dict().keys()) == O.PARAM_NAMES
    assert all(torch.equal(sd[k], v) for k, v in net.state_dict().items())   # same init as the reference
    assert sum(p.numel() for p in net.parameters()) == 31030658
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        net(torch.zeros(1, 1, 188, 188))
    a, b = torch.randn(1, 2, 8, 8), torch.randn(1, 3, 4, 4)
    assert torch.equal(net.crop_and_concat(a, b), O.crop_and_concat(a, b))
    a, b = torch.randn(1, 2, 32, 32), torch.randn(1, 3, 56, 56)
    assert torch.equal(net.crop_and_concat(a, b), O.crop_and_concat(a, b))
    with pytest.raises(RuntimeError):
        net.crop_and_concat(torch.zeros(1, 1, 5, 5), torch.zeros(1, 1, 8, 8))
    # state_dict round trip in the reference's .pth format
    net2 = Unet()
    net2.load_state_dict(net.state_dict())
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))


def test_arena_layout_is_backward_ordered():
    from network import Unet
    st = Unet().engine.store
    order = sorted(st.g_off, key=st.g_off.get)
    assert order[0] == "finalconv.weight" and order[-1] == "conv11c.bias"
    assert all(v % 64 == 0 for v in st.g_off.values())
    assert st.total >= 31030658


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "unet_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void|long long|const char\*)\s+(ub_[a-z0-9_]+)\s*\(", header, re.M))
    handle = lib.load()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/unet_b200.h but not exported"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert handle.ub_version() == 1
    assert b"no CUDA driver" in handle.ub_error_string(-3)


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b200 import ops
    with pytest.raises(RuntimeError):
        ops.conv1_fwd(torch.zeros(1, 1, 8, 8), torch.zeros(64, 1, 3, 3), torch.zeros(64), torch.zeros(1, 6, 6, 64))


def test_partition_and_buckets():
    for n in (121, 128, 7):
        for w in (1, 2, 4, 8):
            parts = [partition(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in parts) - min(e - s for s, e in parts) <= 1
    ends = [100, 300, 1000, 1100, 5000]
    b = bucket_bounds(5000, ends, 800)
    assert b == [(0, 1000), (1000, 5000)]
    assert bucket_bounds(5000, ends, 10 ** 9) == [(0, 5000)]


def test_drop_in_functions_host_side(golden_dir):
    import functions
    m = functions.evaluation_metrics(torch.tensor([[1, 0], [1, 1]]), torch.tensor([[1, 1], [0, 1]]))
    assert m.shape == (2, 1) and abs(m[0, 0] - 0.5) < 1e-12 and abs(m[1, 0] - 0.5) < 1e-12
    assert functions.input_size_compute(torch.zeros(1, 1, 388, 388)) == (388, 572, 388)
    if not torch.cuda.is_available():
        # the weight maps are device kernels: without a CUDA device they raise, they do not fall back to the host
        g = np.load(os.path.join(golden_dir, "class_balance.npz"))
        with pytest.raises(RuntimeError, match="no CPU path"):
            functions.class_balance(torch.from_numpy(g["lab"]))
        with pytest.raises(RuntimeError, match="no CPU path"):
            functions.weighted_map(torch.from_numpy(g["lab"]))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted (GPU box)")
def test_crop_sampler_matches_reference_dataset_init(tmp_path):
    """b200.augment.CropSampler (anchors + weighted crop distribution + draw order) against the reference's own
    ImageDataset.__init__ / __getitem__ code (data.py:66-82, 97-104, 115) on a seeded frame written to disk."""
    import subprocess
    import sys
    script = r'''
import sys, types, os, numpy as np
sys.modules.setdefault("wget", types.ModuleType("wget"))
sys.path.insert(0, "/root/reference")
import warnings; warnings.simplefilter("ignore")
import cv2 as cv
import data as R
root = sys.argv[1]
rng = np.random.RandomState(5)
yy, xx = np.indices((200, 200))
gt = np.zeros((200, 200), np.uint16)
for k in range(12):
    cy, cx = rng.rand(2) * 200; ry, rx = 6 + rng.rand(2) * 30
    gt[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = k + 1
os.makedirs(os.path.join(root, "01")); os.makedirs(os.path.join(root, "01_GT", "SEG"))
cv.imwrite(os.path.join(root, "01", "t000.tif"), (rng.rand(200, 200) * 255).astype(np.uint8))
cv.imwrite(os.path.join(root, "01_GT", "SEG", "man_seg000.tif"), gt)
ds = R.ImageDataset(root, crop=100, ISBI2012=True)
np.save(os.path.join(root, "target.npy"), np.asarray(ds.target[0]))
np.save(os.path.join(root, "dist.npy"), np.asarray(ds.target_weighted_crop_distribution[0]))
np.save(os.path.join(root, "pairs.npy"), np.asarray(ds.pairs))
'''
    r = subprocess.run([sys.executable, "-c", script, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    from b200.augment import CropSampler
    target = np.load(tmp_path / "target.npy")
    cs = CropSampler(target[None], crop=100, skip=10)
    assert np.array_equal(np.asarray(cs.pairs), np.load(tmp_path / "pairs.npy"))
    np.testing.assert_allclose(cs.distribution[0], np.load(tmp_path / "dist.npy"), rtol=1e-9, atol=1e-15)
    np.random.seed(3)
    d = cs.draw([0, 0])
    assert d.shape == (2, 4) and (d[:, 3] % 30 == 0).all() and (d[:, 1:3] >= 0).all() and (d[:, 1:3] <= 100).all()


def test_c_abi_host_side_queries_and_argument_checks():
    """Entry points that return before any launch: workspace-size queries and argument validation (no GPU needed)."""
    import ctypes as C
    from b200 import lib
    h = lib.load()
    # weight maps: counts (8 B / sample) + component labels (4 B / px) + 2 candidates (8 B / px), each 256-aligned
    al = lambda v: (v + 255) & ~255
    assert h.ub_weight_map_workspace_bytes(16, 388, 388) == al(8 * 16) + al(4 * 16 * 388 * 388) + al(8 * 16 * 388 * 388)
    assert h.ub_weight_map_workspace_bytes(0, 388, 388) == 0
    assert h.ub_augment_workspace_bytes(16, 388, 572, C.c_double(10.0)) > 2 * 16 * 2 * 800 * 800 * 8   # two fp64 windows
    assert h.ub_augment_workspace_bytes(16, 388, 572, C.c_double(40.0)) == 0                         # sigma beyond the smem tile
    assert h.ub_loss_workspace_bytes() > 0
    null = C.c_void_p(0)
    assert h.ub_weight_map(null, 8, 1, 16, 16, C.c_double(20), C.c_double(25), 1, null, null, null, null) == -5      # UB_E_ARG
    assert h.ub_mask_metrics(null, 1, null, 1, 10, null, null) == -5
    assert h.ub_mirror_pad(null, null, 1, 10, 20, null) == -5
    assert b"null" in h.ub_error_string(-5).lower() or h.ub_error_string(-5)
