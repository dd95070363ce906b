"""CPU: the oracle against the golden vectors produced from the reference itself
(oracle/validate_against_reference.py), and -- when /root/reference is mounted -- against the
reference directly."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import unet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def digest(g):
    g = g.detach().double().flatten()
    head = (g[:5].tolist() + [0.0] * 5)[:5]
    return np.array([g.norm().item(), g.sum().item(), g.abs().max().item()] + head)


@pytest.mark.parametrize("name", ["unet_S188_B1_seed0", "unet_S220_B1_seed1"])
def test_oracle_reproduces_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    S, B, seed = int(g["S"]), int(g["B"]), int(g["seed"])
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.init_state_dict(seed)
    images, labels, weight = O.synthetic_batch(B, S, seed=int(g["data_seed"]))
    logits = O.forward(sd, images)
    np.testing.assert_allclose(logits.numpy(), g["logits0"], rtol=1e-5, atol=1e-4)
    mom = {}
    for step in range(2):
        loss, _, grads = O.train_step(sd, mom, images, labels, weight)
        assert abs(loss.item() - float(g[f"loss{step}"])) <= 1e-5 * abs(float(g[f"loss{step}"]))
        d = np.stack([digest(grads[k]) for k in O.PARAM_NAMES])
        ref = g[f"grad_digest{step}"]
        np.testing.assert_allclose(d[:, 0], ref[:, 0], rtol=1e-4)           # per-parameter gradient norms
        np.testing.assert_allclose(d[:, 2], ref[:, 2], rtol=1e-4)           # per-parameter max |g|
    after = O.forward(sd, images)
    np.testing.assert_allclose(after.detach().numpy(), g["logits_after2"], rtol=1e-4, atol=1e-3)


def test_pool_tie_rule_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pool_ties.npz"))
    v, i = O.pool_with_indices(torch.from_numpy(g["x"]))
    assert np.array_equal(v.numpy(), g["v"]) and np.array_equal(i.numpy(), g["i"])


def test_output_sizes_and_crop_rule():
    assert [O.output_size(s) for s in (188, 220, 252, 572, 204)] == [4, 36, 68, 388, None]
    assert O.crop_offset(32, 56) == -12 and O.crop_offset(284, 392) == -54 and O.crop_offset(20, 8) == 6
    assert O.crop_offset(5, 8) == -1  # truncation toward zero -> torch.cat then raises
    with pytest.raises(RuntimeError):
        O.crop_and_concat(torch.zeros(1, 1, 5, 5), torch.zeros(1, 1, 8, 8))


def test_init_std_quirk():
    sd = O.init_state_dict(0)
    assert abs(sd["conv12c.weight"].std().item() - 2 / (576 ** 0.5)) < 2e-3   # 2/sqrt(N), not sqrt(2/N)
    assert abs(sd["conv11c.weight"].std().item() - 2 ** 0.5) < 0.15


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted (GPU box)")
def test_oracle_bit_equal_to_reference():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "validate_against_reference.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["weightmap_blobs_96", "weightmap_noise_48", "weightmap_single_64", "weightmap_blobs_388"])
def test_weightmap_oracle_reproduces_reference_golden(golden_dir, name):
    """oracle/weightmap_oracle.py (scipy, no OpenCV) against the outputs of the reference's own weighted_map /
    class_balance (functions.py:7-117) stored by validate_against_reference.py.  Tolerance 2.5e-6 absolute: OpenCV's
    float sqrt is 1 ulp off on part of the pixels, everything else is bit-equal (see the oracle's header)."""
    import weightmap_oracle as W
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    lab = W.synthetic_labels(int(g["batch"]), int(g["size"]), int(g["seed"]), str(g["kind"]))
    for tag, dt in (("i64", np.int64), ("f32", np.float32)):
        w = W.weighted_map(lab.astype(dt))
        assert w.dtype == np.float32 and float(np.abs(w - g["w_" + tag]).max()) <= 2.5e-6
        assert (w == g["w_" + tag]).mean() > 0.5
    assert np.array_equal(W.class_balance(lab), g["cb"])
    # the integer-label quirk (functions.py:37-40): w_c is truncated, so background weights lose the class ratio
    bg = lab == 0
    assert np.all(g["w_i64"][bg] <= g["w_f32"][bg] + 1e-6)


def test_weightmap_oracle_raises_like_reference():
    import weightmap_oracle as W
    with pytest.raises(IndexError):
        W.weighted_map(np.zeros((1, 8, 8), dtype=np.int64))      # single class: counts[1] (functions.py:40)
    with pytest.raises(IndexError):
        W.class_balance(np.ones((1, 8, 8), dtype=np.int64))


@pytest.mark.parametrize("name", ["augment_crop100", "augment_crop388"])
def test_augment_oracle_reproduces_reference_golden(golden_dir, name):
    """oracle/augment_oracle.py (own spline rotation / Gaussian / bilinear code, no scipy.ndimage transforms) against
    the output of the reference's unmodified ImageDataset.__getitem__ (data.py:93-137) for the recorded draws."""
    import augment_oracle as A
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    size, crop, seed = int(g["size"]), int(g["crop"]), int(g["seed"])
    img, gt = A.synthetic_dataset_image(size, seed)
    _, x, y, rot = (int(v) for v in g["draws"])
    S = A.input_size_compute(crop)[1]
    noise = A.noise_fields(int(g["noise_seed"]), S)
    o_img, o_gt = A.augment_sample(img, gt, x, y, rot, noise[0], noise[1], crop=crop)
    assert o_img.dtype == np.float32 and o_img.shape == (S, S) and o_gt.shape == (crop, crop)
    assert np.array_equal(o_img, g["image"]) and np.array_equal(o_gt, g["label"])


def test_augment_oracle_pieces_match_scipy():
    import augment_oracle as A
    from scipy import ndimage
    rng = np.random.RandomState(0)
    img = (ndimage.gaussian_filter(rng.rand(60, 60), 2) * 255).astype(np.uint8)
    pad = np.pad(img, 90, mode="reflect")                                   # pad > n - 1: repeated reflection
    idx = A.reflect_index(np.arange(60 + 180), 90, 60)
    assert np.array_equal(pad, img[idx][:, idx])
    for deg in (0, 30, 150, 270):
        ref = ndimage.rotate(pad, deg)
        h, w = ref.shape
        crop = ref[h // 2 - 45:h // 2 + 45, w // 2 - 45:w // 2 + 45]
        assert np.array_equal(A.rotate_center_crop(pad, deg, 90), crop), deg
    field = rng.rand(90, 90)
    assert np.abs(A.gaussian_filter_zero(field, 10) - ndimage.gaussian_filter(field, 10, mode="constant", cval=0)).max() < 1e-14
