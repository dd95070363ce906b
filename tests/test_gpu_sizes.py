"""GPU: the whole path at every kind of input size the reference accepts -- the smallest (188), sizes whose encoder
maps are not multiples of the conv / pool tile shapes, batch sizes that do not divide anything -- against the CPU
oracle.  Exercises the tile tails of every kernel (TMA clipping, fused pool / head epilogues, floor-mode pooling)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def max_rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


def bf16_weights(sd):
    return {k: (v.to(torch.bfloat16).float() if k.endswith("weight") and k not in ("conv11c.weight", "finalconv.weight") else v)
            for k, v in sd.items()}


@pytest.mark.parametrize("S,B", [(188, 1), (188, 3), (220, 2), (252, 5), (284, 2), (316, 1), (348, 3), (380, 1)])
def test_forward_and_step_at_size(S, B):
    import unet_oracle as O
    from network import Unet
    from b200 import train as T
    torch.set_num_threads(os.cpu_count() or 1)
    sd = bf16_weights(O.init_state_dict(S % 7))
    images, labels, weight = O.synthetic_batch(B, S, seed=S + B)
    images = images.to(torch.bfloat16).float()
    net = Unet()
    net.load_state_dict(sd)
    net.to(DEV)
    opt = T.FusedSGD(net)
    opt.zero_grad()
    logits = net(images.to(DEV))
    out = O.output_size(S)
    assert logits.shape == (B, 2, out, out)
    loss = T.weighted_bce_with_logits(logits, labels.to(DEV), weight.to(DEV))
    loss.backward()
    sd_o = {k: v.clone() for k, v in sd.items()}
    oloss, ologits, ograds = O.train_step(sd_o, {}, images, labels, weight)
    assert max_rel(logits.detach(), ologits) < 1.5e-2, max_rel(logits.detach(), ologits)
    assert abs(loss.item() - oloss.item()) <= 1e-2 * abs(oloss.item())
    # small maps have few pixels per ReLU / arg-max decision: the gradient bound is loose here, tight at 572
    worst = max(((p.grad.double().cpu() - ograds[k].double()).norm() / (ograds[k].double().norm() + 1e-300)).item()
                for k, p in net.named_parameters())
    assert worst < (0.6 if S == 188 else 0.3), worst                      # 188 -> a 4 x 4 output: a handful of pixels decide
    mask = net.predict(images.to(DEV), labels.shape[-1])
    assert torch.equal(mask.cpu(), logits.detach().argmax(1).cpu())       # fused argmax == argmax of the logits


def test_unsupported_sizes_raise_like_reference():
    from network import Unet
    net = Unet().to(DEV)
    for S in (204, 190, 300):
        with pytest.raises(RuntimeError):
            net(torch.zeros(1, 1, S, S, device=DEV))


def _nhwc_bf16(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


@pytest.mark.parametrize("S,B", [(188, 3), (220, 2), (252, 5), (284, 2), (316, 1), (348, 3), (380, 1)])
def test_backward_tails_teacher_forced(S, B):
    """The tile tails of every dgrad / wgrad / pool-backward / loss kernel at odd sizes, pinned at the north-star
    gradient tolerance (2e-2 relative).  On tiny maps a single ReLU / arg-max flip in OUR forward changes whole
    gradient tensors, which is why test_forward_and_step_at_size can only bound them loosely; here the forward state
    the backward kernels read -- every saved activation, the pool argmax codes, the skip placement, d loss / d logits --
    is the ORACLE's (rounded to bf16), so what is compared is the backward arithmetic alone (trainer.py:77)."""
    import torch.nn.functional as F
    import unet_oracle as O
    from network import Unet
    from b200.geometry import LEVEL_C
    torch.set_num_threads(os.cpu_count() or 1)
    sd = bf16_weights(O.init_state_dict(S % 5))
    images, labels, weight = O.synthetic_batch(B, S, seed=2 * S + B)
    images = images.to(torch.bfloat16).float()
    # ---- oracle forward with every intermediate, its gradients and d loss / d logits
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    taps = {}
    ologits = O.forward(params, images, taps)
    ologits.retain_grad()
    oloss = O.loss_bce(O.center_crop_preds(ologits, labels.shape[-1]), labels, weight)
    oloss.backward()
    ograds = {k: v.grad for k, v in params.items()}
    # ---- our forward (allocates the workspace), then its saved state is replaced by the oracle's
    net = Unet()
    net.load_state_dict(sd)
    net.to(DEV)
    eng = net.engine
    with torch.no_grad():
        _, ws = eng.forward(images.to(DEV), train=True)
        a, g = ws.act, ws.geom
        dev = lambda t: t.detach().to(DEV)
        for l in (1, 2, 3, 4):
            C = LEVEL_C[l]
            a[f"e{l}1"].copy_(_nhwc_bf16(dev(taps[f"conv{l}1c"])))
            full = taps[f"conv{l}2c"].detach()
            if a[f"e{l}2"] is not None:
                a[f"e{l}2"].copy_(_nhwc_bf16(dev(full)))
            pooled, idx = F.max_pool2d(full, 2, 2, return_indices=True)
            a[f"p{l}"].copy_(_nhwc_bf16(dev(pooled)))
            Wf = full.shape[-1]
            iy, ix = idx // Wf, idx % Wf                                    # flat h * W + w of the (n, c) plane
            code = ((iy % 2) * 2 + (ix % 2)).to(torch.uint8)                # dy * 2 + dx inside the 2 x 2 window
            a[f"i{l}"].copy_(dev(code).permute(0, 2, 3, 1).contiguous())
            c = g.coff[l]
            cat = a[f"cat{l}"]
            cat[..., :C].zero_()
            pb = _nhwc_bf16(dev(pooled))
            if c < 0:
                cat[:, -c:-c + pb.shape[1], -c:-c + pb.shape[2], :C] = pb
            else:
                cat[..., :C] = pb[:, c:c + cat.shape[1], c:c + cat.shape[2]]
            cat[..., C:] = _nhwc_bf16(dev(taps[f"upconv{l}"]))
            a[f"d{l}1"].copy_(_nhwc_bf16(dev(taps[f"conv{l}1e"])))
            a[f"d{l}2"].copy_(_nhwc_bf16(dev(taps[f"conv{l}2e"])))
        a["e51"].copy_(_nhwc_bf16(dev(taps["conv51c"])))
        a["e52"].copy_(_nhwc_bf16(dev(taps["conv52c"])))
        for p in net.parameters():
            p.grad = None
        eng.backward(ws, ologits.grad.to(DEV))
        eng.assign_grads()
    torch.cuda.synchronize()
    for k, p in net.named_parameters():
        err = ((p.grad.double().cpu() - ograds[k].double()).norm() / (ograds[k].double().norm() + 1e-300)).item()
        assert err < 2e-2, (k, err)
