"""First-light diagnostics for the CUDA kernels, run on a B200 through gpurun:

    python tests/gpu_first_light.py all        # every section in its own subprocess
    python tests/gpu_first_light.py conv_fwd   # one section

Unlike the pytest suite this prints error metrics instead of asserting, and isolates every
section in a subprocess so that one faulting kernel (sticky CUDA error) does not hide the rest.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dl-unet_b200"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from b200 import ops  # noqa: E402

DEV = "cuda:0"


def bf(t):
    return t.to(torch.bfloat16)


def nhwc(t):  # NCHW fp32 -> NHWC bf16 contiguous
    return bf(t.permute(0, 2, 3, 1)).contiguous()


def nchw(t):  # NHWC bf16 -> NCHW fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def report(name, got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-30
    rel_l2 = (err.pow(2).sum().sqrt() / (ref.pow(2).sum().sqrt() + 1e-30)).item()
    print(f"  {name:48s} max_abs={err.max().item():.4e} max_rel_to_absmax={err.max().item() / denom:.3e} "
          f"rel_l2={rel_l2:.3e} nan={int(torch.isnan(got).sum())}", flush=True)
    return rel_l2


def sec_probe():
    rows = ops.probe_run()
    print("mode shift pitch base_off  max_abs_err  n_bad")
    for r in rows:
        print(f"  {int(r[0])}    {int(r[1]):2d}    {int(r[2]):2d}     {int(r[3])}      {r[4]:.3e}  {int(r[5])}")


def sec_mma_rate():
    print("   N a_mn b_mn b_layout   M    cycles/MMA   (nominal = N/2 for M=128)")
    for r in ops.probe_run(64, which=1):
        print(f" {int(r[0]):4d}   {int(r[1])}    {int(r[2])}      {int(r[3])}    {int(r[5]):5d}   {r[4]:8.1f}")


CONV_SHAPES = [
    (1, 20, 20, 64, 64), (2, 30, 30, 128, 128), (1, 18, 34, 64, 256), (1, 12, 12, 256, 512),
    (3, 10, 10, 128, 64), (1, 66, 66, 64, 128), (1, 9, 9, 1024, 1024), (2, 140, 140, 64, 64),
]


def sec_conv_fwd():
    torch.manual_seed(0)
    for (n, h, w, cin, cout) in CONV_SHAPES:
        x = torch.randn(n, cin, h, w, device=DEV)
        wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
        b = torch.randn(cout, device=DEV)
        xq, wq = bf(x).float(), bf(wt).float()
        ref = F.relu(F.conv2d(xq, wq, b))
        xh = nhwc(x)
        wpk = ops.pack_weight(wt.contiguous(), ops.PACK_CONV_FWD)
        # packing check
        wref = bf(wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin))
        assert torch.equal(wpk, wref), "pack_weight mode 0 mismatch"
        y = torch.full((n, h - 2, w - 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.conv3x3_fwd(xh, wpk, b, y, relu=True)
        torch.cuda.synchronize()
        report(f"conv_fwd tcgen05 vs torch {n, h, w, cin, cout}", nchw(y), ref)
        yn = ops.naive_conv3x3(xh, wpk, b, True)
        report("  naive vs torch", yn.permute(0, 3, 1, 2), ref)
    # strided views: read a channel slice / spatial window, write into a channel slice of a wider buffer
    n, h, w, cin, cout = 2, 22, 22, 64, 128
    big = torch.randn(n, h + 6, w + 6, 192, device=DEV).to(torch.bfloat16)
    xv = big[:, 3:3 + h, 3:3 + w, 64:128]
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / 24
    b = torch.randn(cout, device=DEV)
    ref = F.relu(F.conv2d(nchw(xv), bf(wt).float(), b))
    out = torch.zeros(n, h - 2, w - 2, 384, dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_fwd(xv, ops.pack_weight(wt, ops.PACK_CONV_FWD), b, out[..., 128:256], relu=True)
    torch.cuda.synchronize()
    report("conv_fwd strided in/out views", nchw(out[..., 128:256]), ref)
    print("  untouched channels stay zero:", bool((out[..., :128] == 0).all() and (out[..., 256:] == 0).all()))


def sec_conv_dgrad():
    torch.manual_seed(1)
    for (n, h, w, cin, cout) in CONV_SHAPES:
        ho, wo = h - 2, w - 2
        dy = torch.randn(n, cout, ho, wo, device=DEV)
        wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cout ** 0.5)
        xsrc = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)  # post-ReLU activations (zeros ~50%)
        ref = F.conv_transpose2d(bf(dy).float(), bf(wt).float())
        refm = ref * (bf(xsrc).float() > 0)
        wd = ops.pack_weight(wt.contiguous(), ops.PACK_CONV_DGRAD)
        dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.conv3x3_dgrad(nhwc(dy), wd, dx, None)
        torch.cuda.synchronize()
        report(f"conv_dgrad {n, h, w, cin, cout}", nchw(dx), ref)
        ops.conv3x3_dgrad(nhwc(dy), wd, dx, nhwc(xsrc))
        torch.cuda.synchronize()
        report("  + fused ReLU mask", nchw(dx), refm)


def sec_conv_wgrad():
    torch.manual_seed(2)
    for (n, h, w, cin, cout) in CONV_SHAPES:
        ho, wo = h - 2, w - 2
        x = torch.randn(n, cin, h, w, device=DEV)
        dy = torch.randn(n, cout, ho, wo, device=DEV)
        xq = bf(x).float().requires_grad_(False)
        wt = torch.zeros(cout, cin, 3, 3, device=DEV, requires_grad=True)
        F.conv2d(xq, wt).backward(bf(dy).float())
        ref = wt.grad
        dwp = ops.wgrad_workspace(ops.PACK_CONV_FWD, nhwc(x), cout)
        ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dwp)
        dw = torch.empty_like(ref)
        ops.unpack_grad(dwp, dw, ops.PACK_CONV_FWD)
        torch.cuda.synchronize()
        report(f"conv_wgrad {n, h, w, cin, cout}", dw, ref)


def sec_wgrad_modes():
    """3x3 wgrad: tap-per-CTA kernel (mode 0) vs haloed-patch kernel without (1) / with (2) fused s taps."""
    torch.manual_seed(2)
    shapes = CONV_SHAPES + [(16, 72, 72, 64, 64), (16, 72, 72, 128, 128), (4, 200, 200, 256, 128), (16, 30, 30, 1024, 1024)]
    for (n, h, w, cin, cout) in shapes:
        ho, wo = h - 2, w - 2
        x = torch.randn(n, cin, h, w, device=DEV)
        dy = torch.randn(n, cout, ho, wo, device=DEV)
        wt = torch.zeros(cout, cin, 3, 3, device=DEV, requires_grad=True)
        F.conv2d(bf(x).float(), wt).backward(bf(dy).float())
        ref = wt.grad
        xh, dyh = nhwc(x), nhwc(dy)
        for mode in (0, 1, 2, 4, 5):
            ops.set_option(1, mode)
            dwp = ops.wgrad_workspace(ops.PACK_CONV_FWD, xh, cout)
            ops.conv3x3_wgrad(xh, dyh, dwp)
            dw = torch.empty_like(ref)
            ops.unpack_grad(dwp, dw, ops.PACK_CONV_FWD)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv3x3_wgrad(xh, dyh, dwp)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 5 * 1e3
            fl = 2.0 * n * ho * wo * cout * cin * 9
            report(f"wgrad mode {mode} {n, h, w, cin, cout} {us:8.1f} us {fl / us / 1e6:7.1f} TF/s", dw, ref)
    ops.set_option(1, 2)


def sec_igemm2():
    """haloed-patch conv kernel (option 2 = 2: wherever legal) vs the tap-per-load kernel (0): parity + time."""
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5)
    shapes = [(1, 20, 20, 64, 64), (2, 30, 30, 128, 128), (1, 18, 34, 64, 256), (3, 10, 10, 128, 64), (1, 66, 66, 64, 128),
              (2, 37, 53, 128, 128), (1, 21, 19, 256, 128), (1, 40, 40, 128, 256), (1, 5, 3, 64, 64),
              (16, 200, 200, 64, 64), (16, 200, 200, 128, 64), (16, 200, 200, 64, 128), (16, 150, 150, 128, 128),
              (8, 200, 200, 256, 128), (8, 200, 200, 128, 256)]
    for (n, h, w, cin, cout) in shapes:
        x = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
        wt = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5)
        b = torch.randn(cout, device=DEV)
        dy = torch.randn(n, cout, h - 2, w - 2, device=DEV)
        big = n * h * w * cin > 4e7
        if not big:
            ref = F.relu(F.conv2d(bf(x).float(), bf(wt).float(), b))
            refd = F.conv_transpose2d(bf(dy).float(), bf(wt).float()) * (bf(x).float() > 0)
        xh, dyh = nhwc(x), nhwc(dy)
        wf, wd = ops.pack_weight(wt, ops.PACK_CONV_FWD), ops.pack_weight(wt, ops.PACK_CONV_DGRAD)
        outs = {}
        for mode in (0, 2):
            ops.set_option(2, mode)
            y = torch.full((n, h - 2, w - 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
            ops.conv3x3_fwd(xh, wf, b, y, relu=True)
            ops.conv3x3_dgrad(dyh, wd, dx, xh)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            for _ in range(3):
                ops.conv3x3_fwd(xh, wf, b, y, relu=True)
            e[1].record()
            for _ in range(3):
                ops.conv3x3_dgrad(dyh, wd, dx, xh)
            e[2].record()
            torch.cuda.synchronize()
            fl = 2.0 * n * (h - 2) * (w - 2) * cin * cout * 9
            tf, td = e[0].elapsed_time(e[1]) / 3 * 1e3, e[1].elapsed_time(e[2]) / 3 * 1e3
            tag = f"mode {mode} {n, h, w, cin, cout} fwd {tf:7.1f}us {fl / tf / 1e6:6.0f}TF/s dgrad {td:7.1f}us {fl / td / 1e6:6.0f}TF/s"
            outs[mode] = (y, dx)
            if not big:
                report(tag + " fwd", nchw(y), ref)
                report("    dgrad+mask", nchw(dx), refd)
            else:
                print("  " + tag, flush=True)
        print("    bit-identical between kernels: fwd", bool(torch.equal(outs[0][0], outs[2][0])), "dgrad", bool(torch.equal(outs[0][1], outs[2][1])))
    ops.set_option(2, 1)


UP_SHAPES = [(1, 8, 8, 128, 64), (2, 14, 14, 1024, 512), (1, 25, 25, 256, 128), (1, 52, 52, 512, 256)]


def sec_upconv():
    torch.manual_seed(3)
    for (n, h, w, cin, cout) in UP_SHAPES:
        x = torch.randn(n, cin, h, w, device=DEV)
        wt = torch.randn(cin, cout, 2, 2, device=DEV) / cin ** 0.5
        b = torch.randn(cout, device=DEV)
        ref = F.conv_transpose2d(bf(x).float(), bf(wt).float(), b, stride=2)
        # write into the upper channel half of a concat buffer
        cat = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device=DEV)
        ops.upconv2x2_fwd(nhwc(x), ops.pack_weight(wt.contiguous(), ops.PACK_UP_FWD), b, cat[..., cout:])
        torch.cuda.synchronize()
        report(f"upconv_fwd {n, h, w, cin, cout}", nchw(cat[..., cout:]), ref)
        print("    skip half untouched:", bool((cat[..., :cout] == 0).all()))
        # dgrad
        dcat = torch.randn(n, 2 * h, 2 * w, 2 * cout, device=DEV).to(torch.bfloat16)
        dyv = dcat[..., cout:]
        refd = F.conv2d(nchw(dyv), bf(wt).float(), stride=2)  # weight [cin, cout, 2, 2] acts as [out, in, 2, 2]
        xsrc = torch.randn(n, cin, h, w, device=DEV).clamp_min(0)
        dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.upconv2x2_dgrad(dyv, ops.pack_weight(wt.contiguous(), ops.PACK_UP_DGRAD), dx, None)
        torch.cuda.synchronize()
        report("  upconv_dgrad", nchw(dx), refd)
        ops.upconv2x2_dgrad(dyv, ops.pack_weight(wt.contiguous(), ops.PACK_UP_DGRAD), dx, nhwc(xsrc))
        torch.cuda.synchronize()
        report("  upconv_dgrad + mask", nchw(dx), refd * (bf(xsrc).float() > 0))
        # wgrad
        wz = torch.zeros(cin, cout, 2, 2, device=DEV, requires_grad=True)
        F.conv_transpose2d(bf(x).float(), wz, stride=2).backward(nchw(dyv))
        dwp = ops.wgrad_workspace(ops.PACK_UP_FWD, nhwc(x), cout)
        ops.upconv2x2_wgrad(nhwc(x), dyv, dwp)
        dw = torch.empty_like(wz)
        ops.unpack_grad(dwp, dw, ops.PACK_UP_FWD)
        torch.cuda.synchronize()
        report("  upconv_wgrad", dw, wz.grad)


def sec_elementwise():
    torch.manual_seed(4)
    # conv1
    n, h, w = 2, 40, 44
    img = torch.rand(n, 1, h, w, device=DEV)
    w1 = torch.randn(64, 1, 3, 3, device=DEV)
    b1 = torch.randn(64, device=DEV)
    y = torch.empty(n, h - 2, w - 2, 64, dtype=torch.bfloat16, device=DEV)
    ops.conv1_fwd(img, w1, b1, y)
    ref = F.relu(F.conv2d(img, w1, b1))
    report("conv1_fwd", nchw(y), ref)
    dy = torch.randn(n, h - 2, w - 2, 64, device=DEV).to(torch.bfloat16)
    wz = torch.zeros_like(w1, requires_grad=True)
    bz = torch.zeros_like(b1, requires_grad=True)
    F.conv2d(img, wz, bz).backward(nchw(dy))
    dw = torch.zeros(64, 9, device=DEV)
    db = torch.zeros(64, device=DEV)
    ops.conv1_wgrad(img, dy, dw, db)
    report("conv1_wgrad dW", dw.view(64, 1, 3, 3), wz.grad)
    report("conv1_wgrad db", db, bz.grad)

    # pool fwd (+ties, + concat write with pad and with crop)
    for (n, h, w, c, hu) in [(2, 16, 16, 64, 12), (1, 8, 8, 128, 8), (2, 20, 20, 64, 8), (1, 24, 24, 256, 20)]:
        x = torch.randn(n, h, w, c, device=DEV).clamp_min(0).to(torch.bfloat16)  # many exact-zero ties
        hp = h // 2
        coff = int((hp - hu) * 0.5)
        yv = torch.empty(n, hp, hp, c, dtype=torch.bfloat16, device=DEV)
        idx = torch.empty(n, hp, hp, c, dtype=torch.uint8, device=DEV)
        cat = torch.zeros(n, hu, hu, 2 * c, dtype=torch.bfloat16, device=DEV)
        ops.pool2x2_fwd(x, yv, idx, cat[..., :c], coff)
        rv, ri = F.max_pool2d(nchw(x), 2, 2, return_indices=True)
        report(f"pool_fwd values {n, h, w, c}", nchw(yv), rv)
        oy = torch.arange(hp, device=DEV).view(1, 1, hp, 1)
        ox = torch.arange(hp, device=DEV).view(1, 1, 1, hp)
        code = idx.permute(0, 3, 1, 2).long()
        flat = (2 * oy + code // 2) * w + (2 * ox + code % 2)
        print("    indices bit-exact:", bool(torch.equal(flat, ri)))
        refcat = F.pad(rv, (-coff,) * 4)
        print(f"    concat write (c_off={coff}) exact:", bool(torch.equal(nchw(cat[..., :c]), refcat)),
              " other half zero:", bool((cat[..., c:] == 0).all()))
        # pool bwd
        dyp = torch.randn(n, hp, hp, c, device=DEV).to(torch.bfloat16)
        dcat = torch.randn(n, hu, hu, 2 * c, device=DEV).to(torch.bfloat16)
        dx = torch.full((n, h, w, c), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.pool2x2_bwd(dyp, dcat[..., :c], coff, idx, yv, dx)
        xr = nchw(x).requires_grad_(True)
        pr = F.max_pool2d(F.relu(xr), 2, 2)
        g = nchw(dyp) + F.pad(nchw(dcat[..., :c]), (coff,) * 4)
        g = bf(g).float()
        pr.backward(g)
        report("    pool_bwd (+skip sum, +relu mask)", nchw(dx), xr.grad)

    # head
    n, h, w = 2, 36, 36
    x = torch.randn(n, h, w, 64, device=DEV).clamp_min(0).to(torch.bfloat16)
    wf = torch.randn(2, 64, 1, 1, device=DEV) / 8
    bfin = torch.randn(2, device=DEV)
    logits = torch.empty(n, 2, h, w, device=DEV)
    am = torch.empty(n, h, w, dtype=torch.uint8, device=DEV)
    ops.head_fwd(x, wf, bfin, logits, am)
    ref = F.conv2d(nchw(x), wf, bfin)
    report("head_fwd logits", logits, ref)
    print("    argmax agreement:", float((am.long() == ref.argmax(1)).float().mean()))
    dl = torch.randn(n, 2, h, w, device=DEV)
    xr = nchw(x).requires_grad_(True)
    wz = wf.clone().requires_grad_(True)
    bz = bfin.clone().requires_grad_(True)
    F.conv2d(F.relu(xr), wz, bz).backward(dl)
    dx = torch.empty_like(x)
    dw = torch.zeros(2, 64, device=DEV)
    db = torch.zeros(2, device=DEV)
    ops.head_bwd(x, wf, dl, dx, dw, db)
    report("head_bwd dx", nchw(dx), xr.grad)
    report("head_bwd dw", dw.view(2, 64, 1, 1), wz.grad)
    report("head_bwd db", db, bz.grad)

    # losses
    for (b, hp, l) in [(1, 36, 36), (3, 40, 36), (2, 388, 388)]:
        z = (torch.randn(b, 2, hp, hp, device=DEV) * 10).requires_grad_(True)
        lab = (torch.rand(b, l, l, device=DEV) > 0.6).long()
        wm = 1 + 10 * torch.rand(b, l, l, device=DEV)
        pad = (hp - l) // 2
        zc = z[:, :, pad:pad + l, pad:pad + l]
        ws = ops.loss_workspace(DEV)
        for kind in (0, 1):
            if kind == 0:
                tgt = torch.stack([1 - lab, lab], 1).float()
                ref = F.binary_cross_entropy_with_logits(zc, tgt, weight=wm.unsqueeze(1))
                scale = 1.0 / (2 * b * l * l)
            else:
                ref = (F.cross_entropy(zc, lab, reduction="none") * wm).mean()
                scale = 1.0 / (b * l * l)
            z.grad = None
            ref.backward()
            out = torch.zeros(1, device=DEV)
            dz = torch.zeros_like(z)
            for lt in (lab, lab.to(torch.uint8)):
                ops.loss_fwd_bwd(kind, z.detach(), lt, wm, dz, out, ws, scale)
                report(f"loss kind={kind} {b, hp, l} {lt.dtype} value", out[0], ref.detach())
                report("    dlogits", dz, z.grad)

    # colsum on a strided channel-slice view
    g = torch.randn(3, 10, 12, 256, device=DEV).to(torch.bfloat16)
    out = torch.zeros(128, device=DEV)
    ops.colsum(g[..., 128:], out)
    report("colsum (strided view)", out, g[..., 128:].float().sum((0, 1, 2)))

    # sgd
    nel = 1000003
    p = torch.randn(nel + 1, device=DEV)[:nel]
    p0 = p.clone()
    g = torch.randn(nel, device=DEV)
    buf = torch.zeros(nel, device=DEV)
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([pr], lr=1e-4, momentum=0.99)
    for step in range(3):
        pr.grad = g.clone() * (step + 1)
        opt.step()
        ops.sgd_momentum(p, g * (step + 1), buf, 1e-4, 0.99, step == 0)
    report("sgd 3 steps p", p, pr.detach())

    # pack modes 1..3 + unpack
    wt = torch.randn(128, 64, 3, 3, device=DEV)
    ref1 = bf(torch.flip(wt, (2, 3)).permute(1, 2, 3, 0).reshape(64, 9 * 128))
    print("pack mode1 exact:", bool(torch.equal(ops.pack_weight(wt, 1), ref1)))
    wu = torch.randn(128, 64, 2, 2, device=DEV)
    ref2 = bf(wu.permute(2, 3, 1, 0).reshape(4 * 64, 128))
    ref3 = bf(wu.permute(0, 2, 3, 1).reshape(128, 4 * 64))
    print("pack mode2 exact:", bool(torch.equal(ops.pack_weight(wu, 2), ref2)))
    print("pack mode3 exact:", bool(torch.equal(ops.pack_weight(wu, 3), ref3)))


SECTIONS = {
    "probe": sec_probe, "conv_fwd": sec_conv_fwd, "conv_dgrad": sec_conv_dgrad, "conv_wgrad": sec_conv_wgrad,
    "upconv": sec_upconv, "elementwise": sec_elementwise, "wgrad_modes": sec_wgrad_modes, "mma_rate": sec_mma_rate, "igemm2": sec_igemm2,
}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        rc = 0
        for name in SECTIONS:
            print(f"===== {name} =====", flush=True)
            t0 = time.time()
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=600)
            print(f"===== {name}: exit {p.returncode} in {time.time() - t0:.1f}s =====", flush=True)
            rc |= p.returncode
        sys.exit(rc)
    SECTIONS[which]()
    torch.cuda.synchronize()
    rc, code = ops.device_error()
    print(f"device_error rc={rc} code={code:#x}")
