"""GPU: the reference's call surface end to end -- `training(unet, train_loader, val_loader, epochs, batch_size,
device, fold_dir, DATASET)` (trainer.py:15) and `testing(unet, test_loader, batch_size, device, output_dir)`
(tester.py:13) -- running on the B200 kernels with device-side weight maps and metrics, against the CPU oracle
stepping through the same batches (trainer.py:54-78 restated in oracle/unet_oracle.py, weight maps from
oracle/weightmap_oracle.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
S, L = 220, 36


def loaders(n_train, n_val, seed):
    """Lists of (images f32 [1,1,S,S] CPU, labels i64 [1,1,L,L] CPU): what the reference's DataLoader yields
    at batch size 1 (the size at which its weight broadcast is well defined)."""
    import weightmap_oracle as W
    g = torch.Generator().manual_seed(seed)
    out, k = [], seed * 100
    while len(out) < n_train + n_val:
        full = torch.from_numpy(W.synthetic_labels(1, S, k, "blobs"))[:, None]               # blobs over the whole tile
        k += 1
        h = (S - L) // 2
        lab = full[:, :, h:h + L, h:h + L].contiguous()                                        # label = its centre
        if not 0.1 * L * L < int(lab.sum()) < 0.9 * L * L:                                     # both classes present
            continue
        img = 0.2 + 0.5 * full.float() + 0.3 * torch.rand(1, 1, S, S, generator=g)
        out.append((img, lab))
    return out[:n_train], out[n_train:]


@pytest.mark.parametrize("hela", [True, False])
def test_training_drop_in_matches_oracle_epochs(tmp_path, hela, capsys):
    import unet_oracle as O
    import weightmap_oracle as W
    import trainer
    from network import Unet
    torch.set_num_threads(os.cpu_count() or 1)
    train, val = loaders(3, 2, 5)
    dataset = trainer.HELA if hela else "ISBI2012-from-argparse"        # identity test, as upstream (trainer.py:68)
    torch.manual_seed(4)
    net = Unet().to(DEV)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    trainer.training(net, train, val, 1, 1, DEV, str(tmp_path), dataset)   # epochs 0 and 1 (range(epochs + 1))
    out = capsys.readouterr().out
    assert "Training is finished" in out and "Model has been saved" in out
    # ---- the same two epochs on the oracle
    wm = (lambda l: W.weighted_map(l)) if hela else (lambda l: W.class_balance(l))
    mom, losses, vlosses = {}, [], []
    for epoch in range(2):
        tot = 0.0
        for img, lab in train:
            w = torch.from_numpy(wm(lab[:, 0].numpy()))
            loss, _, _ = O.train_step(sd, mom, img, lab, w)
            tot += loss.item()
        losses.append(tot / len(train))
        vt = 0.0
        for img, lab in val:
            w = torch.from_numpy(wm(lab[:, 0].numpy()))
            vt += O.loss_bce(O.center_crop_preds(O.forward(sd, img), L), lab, w).item()
        vlosses.append(vt / len(val))
    got = np.loadtxt(tmp_path / "progress" / "loss.out")
    gotv = np.loadtxt(tmp_path / "progress" / "loss_val.out")
    # epoch 0 is the same computation up to bf16 (north star: loss within 1e-2 relative); epoch 1 follows weights that
    # already took three bf16-gradient SGD steps from a chaotic start (loss 100+ -> 10), hence the looser bound
    assert abs(got[0] - losses[0]) <= 2e-2 * abs(losses[0]), (got, losses)
    assert abs(got[1] - losses[1]) <= 0.15 * abs(losses[1]), (got, losses)
    assert abs(gotv[0] - vlosses[0]) <= 5e-2 * abs(vlosses[0]), (gotv, vlosses)
    for f in ("train_eval_iou", "train_eval_pe", "val_eval_iou", "val_eval_pe"):
        assert np.loadtxt(tmp_path / "progress" / (f + ".out")).shape == (2,)
    ck = torch.load(tmp_path / "models" / "unet_weight_save_best.pth", map_location="cpu")
    assert list(ck.keys()) == O.PARAM_NAMES and all(v.dtype == torch.float32 for v in ck.values())
    fresh = Unet()
    fresh.load_state_dict(ck)                                             # the 46-key checkpoint round-trips


def test_testing_drop_in_writes_reference_outputs(tmp_path, capsys):
    import unet_oracle as O
    import tester
    import functions as Fn
    from network import Unet
    torch.manual_seed(6)
    net = Unet().to(DEV)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    _, test = loaders(0, 2, 9)
    tester.testing(net, test, 1, DEV, str(tmp_path))
    assert "Testing is finished" in capsys.readouterr().out
    for sub, stem in (("images", "image"), ("preds", "pred"), ("labels", "label")):
        assert sorted(os.listdir(tmp_path / sub)) == [f"{stem}0.tif", f"{stem}1.tif"]
    iou = np.loadtxt(tmp_path / "test_iou.out")
    pe = np.loadtxt(tmp_path / "test_pe.out")
    img, lab = test[0]
    omask = O.predict_mask(sd, img, L)
    ref = Fn.evaluation_metrics(omask[0], lab[0, 0])                      # CPU route of the reference formulas
    # statistics of the FIRST image only (tester.py:42 drops the concatenation), std over a single value = 0
    assert iou[1] == 0 and pe[1] == 0
    assert abs(pe[0] - ref[1, 0]) <= 0.02 and (np.isnan(iou[0]) == np.isnan(ref[0, 0])) and (np.isnan(iou[0]) or abs(iou[0] - ref[0, 0]) <= 0.05)


def test_async_checkpoint_equals_torch_save(tmp_path):
    """b200.checkpoint.AsyncCheckpointer writes what torch.save(unet.state_dict()) writes (46 keys, fp32, same values),
    captures the weights of the moment save() was called even if training goes on, and round-trips."""
    import unet_oracle as O
    from network import Unet
    from b200 import train as T
    from b200.checkpoint import AsyncCheckpointer
    torch.manual_seed(2)
    net = Unet().to(DEV)
    images, labels, weight = O.synthetic_batch(1, 220, seed=3)
    step = T.TrainStep(net, T.FusedSGD(net), graph=False)
    step(images.to(DEV), labels.to(DEV), weight.to(DEV))
    ref = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    ck = AsyncCheckpointer()
    ck.save(net, str(tmp_path / "a.pth"))
    for _ in range(3):                                            # training continues while the file is written
        step(images.to(DEV), labels.to(DEV), weight.to(DEV))
    ck.wait()
    got = torch.load(tmp_path / "a.pth")
    assert list(got.keys()) == O.PARAM_NAMES
    for k in ref:
        assert got[k].dtype == torch.float32 and got[k].shape == ref[k].shape and torch.equal(got[k], ref[k]), k
    assert not torch.equal(ref["conv12c.weight"], net.state_dict()["conv12c.weight"].cpu())   # weights did move on
    fresh = Unet()
    fresh.load_state_dict(got)


def test_eval_after_graph_replay_uses_current_weights():
    """A CUDA-graph replay updates the fp32 masters without running any Python: the bf16 operand caches must be
    re-derived before the next forward outside the graph (trainer.py:95-121 validates after every epoch).  step, eval,
    step, eval -- the second eval must equal a forward of a fresh module loaded with the current state_dict."""
    import unet_oracle as O
    from network import Unet
    from b200 import train as T
    torch.manual_seed(8)
    net = Unet().to(DEV)
    images, labels, weight = (t.to(DEV) for t in O.synthetic_batch(1, 220, seed=5))
    step = T.TrainStep(net, T.FusedSGD(net), graph=True)
    for _ in range(3):                       # eager first step, capture, first replay
        step(images, labels, weight)
    with torch.no_grad():
        ev1 = net(images).clone()
    step(images, labels, weight)             # replay: masters move, no Python between pack_all and sgd
    with torch.no_grad():
        ev2 = net(images).clone()
        mask2 = net.predict(images)
    fresh = Unet()
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
    fresh.to(DEV)
    with torch.no_grad():
        want = fresh(images)
    assert torch.equal(ev2, want)
    assert torch.equal(mask2, want.argmax(1))
    assert not torch.equal(ev1, ev2)         # the step did change the weights


def test_testing_tile_route_equals_per_tile_predict(tmp_path):
    """testing(..., tile_in=...) -- the overlap-tile route of the drop-in (b200.tiling.TileSegmenter behind the
    reference's signature): the written prediction equals the per-tile predict() masks stitched together."""
    import numpy as np
    import tester
    from network import Unet
    from b200 import tiling
    torch.manual_seed(11)
    net = Unet().to(DEV)
    size, S = 300, 508                      # label / mirror-padded input size, as ImageDataset_test would deliver
    raw = torch.rand(size, size)
    pad = (S - size) // 2
    image = torch.nn.functional.pad(raw[None, None], (pad,) * 4, mode="reflect")
    label = (torch.rand(1, 1, size, size) > 0.5).long()
    tester.testing(net, [(image, label)], 1, DEV, str(tmp_path), tile_in=252)
    tiles = tiling.cut_tiles(raw.to(DEV), 252)
    ref = torch.cat([net.predict(tiles[i:i + 1]) for i in range(tiles.shape[0])]).to(torch.uint8)
    want = tiling.stitch(ref, size, size, 252).cpu().numpy()
    from PIL import Image
    got = np.array(Image.open(tmp_path / "preds" / "pred0.tif"))
    got = got[..., 0] if got.ndim == 3 else got
    assert np.array_equal(got > 0, want > 0)
    assert os.path.exists(tmp_path / "test_iou.out")
