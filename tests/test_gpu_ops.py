"""GPU parity tests of every libstp kernel against the CPU oracle, called through the C ABI.

Tolerances (stated per north_star: 1e-3 relative fp32; index work bit exact):
  * f32 outputs (logits, dW, loss scalars): rel L2 error <= 1e-4 against fp32 math on the same bf16 operands
  * bf16 outputs: rel L2 error <= 3e-3 (one bf16 rounding = 2^-9 max per element)
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from segmentation_training_pipeline_b200 import lib
from tests.util import T, bf16_round, conv_ref, max_abs, rand_bf16, ref, rel_err, stream

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4
TOL_BF16 = 3e-3


def _ws(nbytes, dev):
    return torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad, up
    (2, 16, 16, 64, 64, 3, 1, 1, 1),
    (2, 16, 16, 64, 128, 3, 2, 1, 1),
    (2, 16, 16, 64, 128, 1, 2, 0, 1),
    (1, 32, 32, 8, 64, 7, 2, 3, 1),
    (1, 24, 40, 32, 16, 3, 1, 1, 1),
    (2, 8, 8, 128, 32, 3, 1, 1, 1),
    (1, 8, 8, 256, 256, 3, 1, 1, 1),
    (1, 9, 7, 16, 24, 3, 1, 1, 1),     # odd sizes, Cout not a tile multiple
    (1, 8, 8, 64, 64, 1, 1, 0, 1),
    (1, 8, 8, 32, 16, 4, 1, 2, 2),     # transposed-conv style zero insertion (k4 s2 'same': pad' = 2)
    # shapes aimed at the tcgen05 path: wide N tiles, partial pixel rectangles, 64B / 32B swizzle K blocks
    (1, 12, 24, 128, 256, 3, 1, 1, 1),
    (1, 9, 40, 32, 32, 3, 1, 1, 1),
    (2, 10, 136, 16, 16, 3, 1, 1, 1),
    (1, 16, 16, 192, 64, 3, 1, 1, 1),
    (1, 8, 8, 64, 384, 3, 1, 1, 1),
    (2, 32, 32, 64, 128, 1, 1, 0, 1),
    # stride 2 on the tcgen05 path (TMA element strides): odd sizes, a 128-pixel-wide tile row, 1x1 shortcut
    (1, 17, 23, 64, 64, 3, 2, 1, 1),
    (1, 6, 260, 64, 32, 3, 2, 1, 1),
    (2, 16, 16, 128, 256, 1, 2, 0, 1),
    (1, 16, 24, 32, 64, 4, 1, 2, 1),   # the space-to-depth stem shape: 4x4, Cin 32, Cout 64
    # MobileNetV2 widths (DeepLabV3): Cin a multiple of 32 / 16 but not of 64 (generic kernel; routing them to the first-generation
    # tcgen05 kernel with 32- / 16-channel K blocks measured SLOWER on the DeepLabV3 step: 11.6 vs 10.4 ms)
    (2, 20, 20, 96, 576, 1, 1, 0, 1),
    (1, 20, 20, 160, 960, 1, 1, 0, 1),
    (1, 16, 16, 144, 32, 1, 1, 0, 1),
    (1, 16, 16, 48, 64, 3, 1, 1, 1),
]
TC_WGRAD_CASES = {0, 1, 2, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19}  # ... and whose wgrad must take the tcgen05 wgrad kernel
TC_FWD_CASES = {0, 1, 2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19}  # indices of CONV_CASES whose forward must take the tcgen05 kernel


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("tc", [0, 1])
def test_conv_fwd_dgrad_wgrad(stp, cuda, case, tc):
    n, h, w, cin, cout, k, stride, pad, up = case
    stp.set_tc_enabled(tc)
    try:
        g = torch.Generator().manual_seed(hash(case) % 1000)
        x = rand_bf16((n, h, w, cin), g)
        wt = rand_bf16((cout, k, k, cin), g, scale=1.0 / math.sqrt(k * k * cin))
        if up > 1:
            ho, wo = h * up, w * up
        else:
            ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        res = rand_bf16((n, ho, wo, cout), g)
        desc = lib.ConvDesc(k, k, stride, pad, pad, up, 0)
        # ---- forward (bf16 out, with residual) ----
        y = torch.zeros((n, ho, wo, cout), dtype=torch.bfloat16, device=cuda)
        xs, ys, rs = T(x), T(y), T(res)
        tc0 = stp.tc_launch_count()
        stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(rs), ref(ys), None, 0, stream())
        used_tc = stp.tc_launch_count() - tc0
        # (4 launches: a zero-insertion forward -- case 9, the transposed-conv form -- runs as one halo-kernel launch per output parity class)
        assert used_tc == ((4 if up == 2 else 1) if (tc and CONV_CASES.index(case) in TC_FWD_CASES) else 0), used_tc
        yr = conv_ref(x, wt, stride, pad, up, (ho, wo)) + res.float().cpu()
        assert rel_err(y, yr) < TOL_BF16
        # ---- forward f32 out with bias ----
        yf = torch.zeros((n, ho, wo, cout), dtype=torch.float32, device=cuda)
        bias = torch.randn(cout, generator=g).to(cuda)
        yfs = T(yf)
        stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), bias.data_ptr(), None, ref(yfs), None, 0, stream())
        yr2 = conv_ref(x, wt, stride, pad, up, (ho, wo)) + bias.cpu()
        assert rel_err(yf, yr2) < TOL_F32
        # ---- autograd reference for dgrad / wgrad ----
        xr = x.float().cpu().requires_grad_(True)
        wr = wt.float().cpu().requires_grad_(True)
        dy = rand_bf16((n, ho, wo, cout), g)
        out = conv_ref_autograd(xr, wr, stride, pad, up, (ho, wo))
        out.backward(dy.float().cpu())
        # ---- dgrad ----
        wm = wt.float().contiguous()
        wf = torch.zeros_like(wt)
        wd = torch.zeros_like(wt)
        stp.weight_prep(wm.data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, k, k, cin, stream())
        assert torch.equal(wf, wt)
        dx = torch.zeros_like(x)
        dys, dxs = T(dy), T(dx)
        stp.conv_dgrad(C.byref(desc), ref(dys), wd.data_ptr(), None, ref(dxs), None, 0, stream())
        assert rel_err(dx, xr.grad) < TOL_BF16
        # accumulate form (residual == output)
        dx2 = x.clone()
        dx2s = T(dx2)
        stp.conv_dgrad(C.byref(desc), ref(dys), wd.data_ptr(), ref(dx2s), ref(dx2s), None, 0, stream())
        assert rel_err(dx2, xr.grad + x.float().cpu()) < TOL_BF16
        # ---- wgrad ----
        nws = stp.conv_wgrad_workspace(C.byref(desc), ref(xs), ref(dys))
        ws = _ws(nws, cuda)
        dw = torch.zeros((cout, k, k, cin), dtype=torch.float32, device=cuda)
        tc0 = stp.tc_launch_count()
        stp.conv_wgrad(C.byref(desc), ref(xs), ref(dys), dw.data_ptr(), ws.data_ptr(), ws.numel(), stream())
        used_tc = stp.tc_launch_count() - tc0
        assert used_tc == (1 if (tc and CONV_CASES.index(case) in TC_WGRAD_CASES) else 0), used_tc
        assert rel_err(dw, wr.grad) < TOL_F32
        torch.cuda.synchronize()
    finally:
        stp.set_tc_enabled(1)


def conv_ref_autograd(x_nhwc, w_krsc, stride, pad, up, out_hw):
    x = x_nhwc.permute(0, 3, 1, 2)
    w = w_krsc.permute(0, 3, 1, 2)
    if up > 1:
        n, c, h, ww = x.shape
        z = torch.zeros(n, c, (h - 1) * up + 1, (ww - 1) * up + 1, dtype=x.dtype)
        z[:, :, ::up, ::up] = x
        x = z
    R, S = w.shape[2], w.shape[3]
    ho, wo = out_hw
    pb = max((ho - 1) * stride + R - x.shape[2] - pad, 0)
    pr = max((wo - 1) * stride + S - x.shape[3] - pad, 0)
    x = F.pad(x, (pad, pr, pad, pb))
    y = F.conv2d(x, w, None, stride=stride)[:, :, :ho, :wo]
    return y.permute(0, 2, 3, 1)


def test_conv_strided_views(stp, cuda):
    """input and output as channel slices of wider buffers (zero-copy concat)."""
    g = torch.Generator().manual_seed(5)
    big_in = rand_bf16((2, 12, 12, 96), g)
    big_out = torch.zeros((2, 12, 12, 80), dtype=torch.bfloat16, device=cuda)
    wt = rand_bf16((32, 3, 3, 64), g, scale=0.05)
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
    xs, ys = T(big_in, 32, 64), T(big_out, 16, 32)
    stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), None, 0, stream())
    yr = conv_ref(big_in[..., 32:96], wt, 1, 1)
    assert rel_err(big_out[..., 16:48], yr) < TOL_BF16
    assert float(big_out[..., :16].abs().max()) == 0 and float(big_out[..., 48:].abs().max()) == 0


@pytest.mark.parametrize("shape,up", [((2, 16, 16, 64), 1), ((2, 8, 8, 128), 2), ((1, 12, 20, 16), 1), ((3, 4, 4, 512), 2),
                                      ((2, 40, 24, 192), 1), ((4, 128, 128, 16), 2), ((1, 3, 5, 2048), 1)])
def test_batchnorm_fwd_bwd(stp, cuda, shape, up):
    from oracle import nn as ON
    n, h, w, c = shape
    g = torch.Generator().manual_seed(c + up)
    x = rand_bf16(shape, g, scale=2.0)
    x = (x.float() + torch.randn(c, generator=g).to(cuda) * 0.5).to(torch.bfloat16)
    gamma = (torch.rand(c, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(c, generator=g) * 0.2).to(cuda)
    eps, mom = 2e-5, 0.99
    rows = n * h * w
    nblk = stp.bn_nblk(rows, c)
    partial = torch.zeros(2 * nblk * c, device=cuda)
    coef = torch.zeros(4 * c, device=cuda)
    mm, mv = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
    y = torch.zeros((n, h * up, w * up, c), dtype=torch.bfloat16, device=cuda)
    xs, ys = T(x), T(y)
    stp.bn_stats(ref(xs), partial.data_ptr(), stream())
    stp.bn_finalize(partial.data_ptr(), nblk, c, rows, gamma.data_ptr(), beta.data_ptr(), eps, mom, mm.data_ptr(),
                    mv.data_ptr(), coef.data_ptr(), stream())
    stp.bn_apply(ref(xs), coef.data_ptr(), 1, up, ref(ys), stream())
    # oracle
    xc = x.float().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    gc, bc = gamma.cpu().requires_grad_(True), beta.cpu().requires_grad_(True)
    yo, mean, var = ON.batchnorm_train(xc, gc, bc, eps)
    yo = torch.relu(yo)
    if up == 2:
        yo = ON.upsample_nearest(yo, 2)
    assert max_abs(coef[:c], mean) < 1e-4 * (1 + float(mean.abs().max()))
    assert rel_err(coef[c:2 * c], torch.rsqrt(var + eps)) < 1e-5
    assert rel_err(y, yo.permute(0, 2, 3, 1)) < TOL_BF16
    unb = var * rows / (rows - (1.0 + eps))   # keras: sample_size / (sample_size - (1 + epsilon))
    assert rel_err(mv, 0.99 * torch.ones(c) + 0.01 * unb.detach()) < 1e-5
    assert max_abs(mm, 0.01 * mean.detach()) < 1e-5
    # backward
    dy = rand_bf16((n, h * up, w * up, c), g)
    yo.backward(dy.float().cpu().permute(0, 3, 1, 2))
    bcoef = torch.zeros(3 * c, device=cuda)
    dgamma, dbeta = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
    dx = torch.zeros_like(x)
    dys, dxs = T(dy), T(dx)
    stp.bn_bwd_reduce(ref(dys), ref(xs), coef.data_ptr(), 1, up, partial.data_ptr(), stream())
    stp.bn_bwd_finalize(partial.data_ptr(), nblk, c, rows, coef.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                        bcoef.data_ptr(), stream())
    stp.bn_bwd_apply(ref(dys), ref(xs), coef.data_ptr(), bcoef.data_ptr(), 1, up, None, ref(dxs), stream())
    assert rel_err(dgamma, gc.grad) < 1e-3
    assert rel_err(dbeta, bc.grad) < 1e-3
    assert rel_err(dx, xc.grad.permute(0, 2, 3, 1)) < 5e-3
    # one-launch variants (last block finalises): same numbers, ticket returns to zero so they can be replayed
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * c, dtype=torch.float64, device=cuda)
    for it in range(4):  # deterministic partial reduction (acc NULL) twice, double-atomic accumulation twice
        accp = acc.data_ptr() if it >= 2 else None
        coef2, bcoef2 = torch.zeros(4 * c, device=cuda), torch.zeros(3 * c, device=cuda)
        dg2, db2 = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
        mm2, mv2 = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
        stp.bn_stats_fused(ref(xs), partial.data_ptr(), sync.data_ptr(), accp, gamma.data_ptr(), beta.data_ptr(), eps, mom,
                           mm2.data_ptr(), mv2.data_ptr(), coef2.data_ptr(), stream())
        stp.bn_bwd_reduce_fused(ref(dys), ref(xs), coef.data_ptr(), 1, up, partial.data_ptr(), sync.data_ptr(), accp,
                                dg2.data_ptr(), db2.data_ptr(), bcoef2.data_ptr(), stream())
        assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
        assert max_abs(coef2, coef) <= 1e-6 * (1 + float(coef.abs().max()))
        assert max_abs(mm2, mm) < 1e-7 and rel_err(mv2, mv) < 1e-6
        assert max_abs(bcoef2, bcoef) <= 1e-6 * (1 + float(bcoef.abs().max()))
        assert rel_err(dg2, dgamma) < 1e-6 and rel_err(db2, dbeta) < 1e-6
    # with residual
    r = rand_bf16(shape, g)
    dx2 = torch.zeros_like(x)
    rs, dx2s = T(r), T(dx2)
    stp.bn_bwd_apply(ref(dys), ref(xs), coef.data_ptr(), bcoef.data_ptr(), 1, up, ref(rs), ref(dx2s), stream())
    assert rel_err(dx2, xc.grad.permute(0, 2, 3, 1) + r.float().cpu()) < 5e-3


def test_input_norm_u8(stp, cuda):
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (2, 32, 32, 3), generator=g, dtype=torch.uint8).to(cuda)
    rows, c = 2 * 32 * 32, 3
    nblk = stp.bn_nblk(rows, c)
    partial = torch.zeros(2 * nblk * 8, device=cuda)
    coef = torch.zeros(4 * c, device=cuda)
    beta = torch.tensor([0.1, -0.2, 0.3], device=cuda)
    ims = T(img)
    stp.bn_stats(ref(ims), partial.data_ptr(), stream())
    stp.bn_finalize(partial.data_ptr(), nblk, c, rows, None, beta.data_ptr(), 2e-5, 0.99, None, None, coef.data_ptr(),
                    stream())
    y = torch.zeros((2, 32, 32, 8), dtype=torch.bfloat16, device=cuda)
    ys = T(y)
    stp.stem_prep(img.data_ptr(), 2, 32, 32, 3, coef.data_ptr(), ref(ys), stream())
    x = img.float().cpu()
    mean, var = x.mean(dim=(0, 1, 2)), x.var(dim=(0, 1, 2), unbiased=False)
    yo = (x - mean) * torch.rsqrt(var + 2e-5) + beta.cpu()
    assert rel_err(y[..., :3], yo) < TOL_BF16
    assert float((y[..., 3].float() - 1).abs().max()) == 0
    assert float(y[..., 4:].abs().max()) == 0


@pytest.mark.parametrize("shape,k,s,p", [((2, 16, 16, 64), 3, 2, 1), ((1, 10, 14, 16), 3, 2, 1), ((2, 8, 8, 32), 2, 2, 0),
                                         ((3, 11, 15, 24), 3, 2, 1)])
def test_maxpool(stp, cuda, shape, k, s, p):
    n, h, w, c = shape
    g = torch.Generator().manual_seed(3)
    x = torch.relu(rand_bf16(shape, g))  # post-ReLU data, with exact ties at 0
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    y = torch.zeros((n, ho, wo, c), dtype=torch.bfloat16, device=cuda)
    am = torch.zeros(n * ho * wo * c, dtype=torch.uint8, device=cuda)
    xs, ys = T(x), T(y)
    stp.maxpool_fwd(ref(xs), k, s, p, ref(ys), am.data_ptr(), stream())
    xc = x.float().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    yo = F.max_pool2d(xc, k, s, p)
    assert torch.equal(y.float().cpu(), yo.permute(0, 2, 3, 1).detach())
    dy = rand_bf16((n, ho, wo, c), g)
    yo.backward(dy.float().cpu().permute(0, 3, 1, 2))
    dx = torch.zeros_like(x)
    dys, dxs = T(dy), T(dx)
    stp.maxpool_bwd(ref(dys), am.data_ptr(), k, s, p, None, ref(dxs), stream())
    # gradients routed to zero-valued (tied) inputs are masked by the ReLU upstream: compare where x > 0
    mask = (x.float().cpu() > 0)
    ref_dx = xc.grad.permute(0, 2, 3, 1)
    assert rel_err(dx.float().cpu() * mask, ref_dx * mask) < TOL_BF16
    # accumulate form
    r = rand_bf16(shape, g)
    dx2 = torch.zeros_like(x)
    rs, dx2s = T(r), T(dx2)
    stp.maxpool_bwd(ref(dys), am.data_ptr(), k, s, p, ref(rs), ref(dx2s), stream())
    assert rel_err(dx2.float().cpu() * mask, (ref_dx + r.float().cpu()) * mask) < TOL_BF16


@pytest.mark.parametrize("h,w", [(16, 24), (48, 20), (24, 16)])   # one column strip / three strips per column / per-pixel kernels (24 % 16 != 0)
@pytest.mark.parametrize("classes,cin", [(1, 16), (3, 32)])
def test_head(stp, cuda, classes, cin, h, w):
    g = torch.Generator().manual_seed(7)
    n = 2
    x = rand_bf16((n, h, w, cin), g)
    wt = bf16_round(torch.randn((classes, 3, 3, cin), generator=g) * 0.1).to(cuda)
    bias = torch.randn(classes, generator=g).to(cuda)
    logits = torch.zeros(n * h * w * classes, device=cuda)
    xs = T(x)
    xr = x.float().cpu().requires_grad_(True)
    wr = wt.cpu().requires_grad_(True)
    out = conv_ref_autograd(xr, wr, 1, 1, 1, (h, w)) + bias.cpu()
    # CUDA-core kernel (no workspace) and the tcgen05 path (bf16 weight copy in the workspace) must both match
    stp.head_fwd(ref(xs), wt.data_ptr(), bias.data_ptr(), classes, logits.data_ptr(), None, 0, stream())
    assert rel_err(logits.view(n, h, w, classes), out) < TOL_F32
    logits.zero_()
    fws = _ws(stp.head_fwd_workspace(ref(xs), classes), cuda)
    tc0 = stp.tc_launch_count()
    stp.head_fwd(ref(xs), wt.data_ptr(), bias.data_ptr(), classes, logits.data_ptr(), fws.data_ptr(), fws.numel(), stream())
    assert stp.tc_launch_count() == tc0 + 1
    assert rel_err(logits.view(n, h, w, classes), out) < TOL_F32
    dl = torch.randn((n, h, w, classes), generator=g).to(cuda)
    out.backward(dl.cpu())
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(wt)
    db = torch.zeros(classes, device=cuda)
    ws = _ws(stp.head_bwd_workspace(ref(xs), classes), cuda)
    dxs = T(dx)
    stp.head_bwd(ref(xs), wt.data_ptr(), dl.data_ptr(), classes, ref(dxs), dw.data_ptr(), db.data_ptr(), ws.data_ptr(),
                 ws.numel(), stream())
    assert rel_err(dx, xr.grad) < TOL_BF16
    assert rel_err(dw, wr.grad) < TOL_F32
    assert rel_err(db, dl.cpu().sum(dim=(0, 1, 2))) < TOL_F32


@pytest.mark.parametrize("weights", [(1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.5, 0.1, 0.3), (0.0, 0.0, 0.0, 1.0, 0.0),
                                     (0.0, 0.0, 0.0, 0.0, 1.0), (0.3, 0.2, 0.0, 0.5, 0.7)])
def test_loss(stp, cuda, weights):
    from oracle import losses as OL
    g = torch.Generator().manual_seed(11)
    count = 2 * 40 * 36
    logits = (torch.randn(count, generator=g) * 3).to(cuda)
    # exercise the 1e-7 clip; a logit of exactly 0 is avoided: there autograd frameworks pick a SUBgradient of
    # max(x,0)/|x| (torch 1/0, TF 0/0) while the kernel uses the analytic derivative sigmoid(x)-t
    logits[:5] = torch.tensor([40.0, -40.0, 0.25, 17.0, -17.0])
    mask = (torch.rand(count, generator=g) > 0.7).to(torch.uint8).to(cuda)
    weights = tuple(weights) + (0.0,) * (5 - len(weights))   # (bce, dice, iou, jaccard, focal)
    spec = lib.LossSpec(*weights)
    partial = torch.zeros(stp.loss_partial_floats(), device=cuda)
    result = torch.zeros(16, device=cuda)
    stp.loss_fwd(logits.data_ptr(), mask.data_ptr(), count, C.byref(spec), partial.data_ptr(), result.data_ptr(), stream())
    z = logits.cpu().requires_grad_(True)
    t = mask.float().cpu().view(2, 40, 36, 1)
    p = torch.sigmoid(z).view(2, 40, 36, 1)
    lo = (weights[0] * OL.binary_crossentropy(t, p) + weights[1] * OL.dice_loss(t, p) + weights[2] * OL.iou_loss(t, p) +
          weights[3] * OL.jaccard_loss(t, p) + weights[4] * OL.focal_loss(t, p))
    r = result.cpu()
    assert abs(float(r[lib.L_JACCARD]) - float(OL.jaccard_loss(t, p))) < 1e-5 * max(1.0, float(OL.jaccard_loss(t, p)))
    assert abs(float(r[lib.L_FOCAL]) - float(OL.focal_loss(t, p))) < 1e-5 * max(1.0, float(OL.focal_loss(t, p)))
    assert abs(float(r[lib.L_LOSS]) - float(lo)) <= 1e-5 * max(1.0, abs(float(lo)))
    assert abs(float(r[lib.L_DICE]) - float(OL.dice(t, p))) < 1e-5
    assert abs(float(r[lib.L_IOU]) - float(OL.iou(t, p))) < 1e-5
    assert abs(float(r[lib.L_ACC]) - float(OL.binary_accuracy(t, p))) < 1e-6
    assert abs(float(r[lib.L_IOT]) - float(OL.iot(t, p))) < 1e-5
    lo.backward()
    dl = torch.zeros(count, device=cuda)
    stp.loss_bwd(logits.data_ptr(), mask.data_ptr(), count, C.byref(spec), result.data_ptr(), dl.data_ptr(), stream())
    assert rel_err(dl, z.grad) < 1e-4


def test_loss_known_answers(stp, cuda):
    """closed forms: BCE(p=0.5) = ln 2; dice of identical hard masks -> loss ~ 0."""
    count = 4096
    spec = lib.LossSpec(1.0, 0.0, 0.0)
    partial = torch.zeros(stp.loss_partial_floats(), device=cuda)
    result = torch.zeros(16, device=cuda)
    logits = torch.zeros(count, device=cuda)
    mask = (torch.arange(count) % 2).to(torch.uint8).to(cuda)
    stp.loss_fwd(logits.data_ptr(), mask.data_ptr(), count, C.byref(spec), partial.data_ptr(), result.data_ptr(), stream())
    assert abs(float(result[lib.L_LOSS]) - math.log(2.0)) < 1e-6
    spec = lib.LossSpec(0.0, 1.0, 0.0)
    logits = (mask.float() * 2 - 1) * 50.0
    stp.loss_fwd(logits.data_ptr(), mask.data_ptr(), count, C.byref(spec), partial.data_ptr(), result.data_ptr(), stream())
    assert abs(float(result[lib.L_LOSS])) < 1e-6 and abs(float(result[lib.L_DICE]) - 1.0) < 1e-6


@pytest.mark.parametrize("opt", ["adam", "sgd", "rmsprop", "nadam"])
def test_optimizers(stp, cuda, opt):
    from oracle import optim as OO
    g = torch.Generator().manual_seed(13)
    n = 4096 + 8
    p0 = torch.randn(n, generator=g)
    params = {"p": p0.clone()}
    if opt == "adam":
        o = OO.Adam(params, lr=1e-3, clipnorm=1.0)
    elif opt == "sgd":
        o = OO.SGD(params, lr=0.01, momentum=0.9, nesterov=True, clipvalue=0.5)
    elif opt == "nadam":
        o = OO.Nadam(params, lr=0.002)
    else:
        o = OO.RMSprop(params, lr=1e-3)
    sched = torch.tensor([1.0, 0, 0, 0, 0], device=cuda)
    p = p0.clone().to(cuda)
    m, v = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    d_step = torch.zeros(1, dtype=torch.int64, device=cuda)
    sumsq, part = torch.zeros(1, device=cuda), torch.zeros(1024, device=cuda)
    for it in range(5):
        gr = torch.randn(n, generator=g) * (3.0 if it % 2 else 0.01)
        o.step({"p": gr})
        gd = gr.to(cuda)
        if opt == "adam":
            stp.sumsq(gd.data_ptr(), n, part.data_ptr(), sumsq.data_ptr(), stream())
            gx = lib.GradXform(1.0, 1.0, 0.0, sumsq.data_ptr())
            stp.adam(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-7, C.byref(gx),
                     d_step.data_ptr(), stream())
        elif opt == "nadam":
            gx = lib.GradXform(1.0, 0.0, 0.0, None)
            stp.nadam(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), sched.data_ptr(), n, 0.002, 0.9, 0.999, 1e-7,
                      0.004, C.byref(gx), d_step.data_ptr(), stream())
        elif opt == "sgd":
            gx = lib.GradXform(1.0, 0.0, 0.5, None)
            stp.sgd(p.data_ptr(), gd.data_ptr(), m.data_ptr(), n, 0.01, 0.9, 1, C.byref(gx), stream())
        else:
            gx = lib.GradXform(1.0, 0.0, 0.0, None)
            stp.rmsprop(p.data_ptr(), gd.data_ptr(), m.data_ptr(), n, 1e-3, 0.9, 1e-7, C.byref(gx), stream())
        stp.step_advance(d_step.data_ptr(), stream())
        assert max_abs(p, params["p"]) < 2e-6
    assert int(d_step.item()) == 5


def test_adam_first_step_known_answer(stp, cuda):
    """Keras Adam, t=1: m=(1-b1)g, v=(1-b2)g^2, lr_t=lr*sqrt(1-b2)/(1-b1) -> dp = lr*g/(|g| + eps*sqrt(1-b2))...
    i.e. |dp| ~= lr for |g| >> eps."""
    n = 16
    p = torch.zeros(n, device=cuda)
    gvals = torch.tensor([1.0, -1.0, 1e-3, -1e-3] * 4, device=cuda)
    m, v = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    d_step = torch.zeros(1, dtype=torch.int64, device=cuda)
    gx = lib.GradXform(1.0, 0.0, 0.0, None)
    stp.adam(p.data_ptr(), gvals.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-7, C.byref(gx),
             d_step.data_ptr(), stream())
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expect = -lr_t * (0.1 * gvals.cpu()) / (torch.sqrt(0.001 * gvals.cpu() ** 2) + 1e-7)
    assert max_abs(p, expect) < 1e-8
    assert abs(float(p[0]) + 1e-3) < 1e-6


@pytest.mark.parametrize("mt", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(1, 72, 24, 64, 64), (1, 40, 40, 16, 16), (2, 70, 16, 32, 32), (1, 40, 16, 128, 128),
                                   (1, 33, 8, 64, 128), (1, 36, 20, 192, 32), (1, 20, 20, 32, 16)])
def test_conv_tc2_strip_heights(stp, cuda, shape, mt):
    """second-generation (halo) tcgen05 conv: every strip height MT, partial strips/rectangles, all swizzle widths,
    fused bias / residual / ReLU epilogue; and the stride-1 dgrad through the same kernel."""
    n, h, w, cin, cout = shape
    stp.set_option(b"tc2_force_mt", mt)
    try:
        g = torch.Generator().manual_seed(n * 1000 + h + cin)
        x = rand_bf16((n, h, w, cin), g)
        wt = rand_bf16((cout, 3, 3, cin), g, scale=1.0 / math.sqrt(9 * cin))
        res = rand_bf16((n, h, w, cout), g)
        bias = torch.randn(cout, generator=g).to(cuda)
        desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, lib.CONV_RELU)
        y = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
        xs, ys, rs = T(x), T(y), T(res)
        tc0 = stp.tc_launch_count()
        stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), bias.data_ptr(), ref(rs), ref(ys), None, 0, stream())
        assert stp.tc_launch_count() - tc0 == 1
        yr = torch.relu(conv_ref(x, wt, 1, 1, 1, (h, w)) + res.float().cpu() + bias.cpu())
        assert rel_err(y, yr) < TOL_BF16
        # f32 output, no epilogue extras
        desc0 = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
        yf = torch.zeros((n, h, w, cout), dtype=torch.float32, device=cuda)
        yfs = T(yf)
        stp.conv_fwd(C.byref(desc0), ref(xs), wt.data_ptr(), None, None, ref(yfs), None, 0, stream())
        assert rel_err(yf, conv_ref(x, wt, 1, 1, 1, (h, w))) < TOL_F32
        # dgrad
        wd = torch.zeros((cin, 3, 3, cout), dtype=torch.bfloat16, device=cuda)
        wf = torch.zeros_like(wt)
        stp.weight_prep(wt.float().contiguous().data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, 3, 3, cin, stream())
        dy = rand_bf16((n, h, w, cout), g)
        dx = torch.zeros_like(x)
        dys, dxs = T(dy), T(dx)
        stp.conv_dgrad(C.byref(desc0), ref(dys), wd.data_ptr(), None, ref(dxs), None, 0, stream())
        xr = x.float().cpu().requires_grad_(True)
        conv_ref_autograd(xr, wt.float().cpu(), 1, 1, 1, (h, w)).backward(dy.float().cpu())
        assert rel_err(dx, xr.grad) < TOL_BF16
        torch.cuda.synchronize()
    finally:
        stp.set_option(b"tc2_force_mt", 0)


def test_weight_prep_batched_matches_per_layer(stp, cuda):
    """one-launch weight prep over the flat buffers == the per-layer kernel, bit for bit (odd shapes included)"""
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 7, 7, 8, 0), (64, 3, 3, 64, 1), (24, 3, 3, 16, 1), (128, 1, 1, 64, 1), (40, 4, 4, 72, 1)]
    items, off, tile = [], 0, 0
    for co, r, s_, ci, dg in shapes:
        items.append([off, co, r, s_, ci, dg, tile, 0])
        off += (co * r * s_ * ci + 7) // 8 * 8
        tile += ((co + 31) // 32) * ((ci + 31) // 32)
    flat = torch.randn(off, generator=g).to(cuda)
    wf = torch.zeros(off, dtype=torch.bfloat16, device=cuda)
    wd = torch.zeros(off, dtype=torch.bfloat16, device=cuda)
    wf2, wd2 = torch.zeros_like(wf), torch.zeros_like(wd)
    it = torch.tensor(items, dtype=torch.int64, device=cuda)
    stp.weight_prep_batched(flat.data_ptr(), wf.data_ptr(), wd.data_ptr(), it.data_ptr(), len(items), tile, stream())
    for o, co, r, s_, ci, dg, _, _ in items:
        stp.weight_prep(flat.data_ptr() + 4 * o, wf2.data_ptr() + 2 * o, (wd2.data_ptr() + 2 * o) if dg else None, co, r,
                        s_, ci, stream())
        n = co * r * s_ * ci
        assert torch.equal(wf[o:o + n], wf2[o:o + n])
        assert torch.equal(wd[o:o + n], wd2[o:o + n])


def test_stem_space_to_depth_equals_7x7_stride2(stp, cuda):
    """conv0 as run by the engine (4x4/1 over the space-to-depth bn_data tensor, gradient gathered back) == the 7x7
    stride-2 pad-3 convolution of the reference graph, forward and weight gradient."""
    g = torch.Generator().manual_seed(3)
    n, h, w, cout = 2, 32, 48, 64
    img = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8).to(cuda)
    coef = torch.cat([torch.full((3,), 120.0), torch.full((3,), 0.02), torch.tensor([0.02, 0.015, 0.025]),
                      torch.tensor([-2.0, -1.5, -2.5])]).to(cuda)
    x8 = torch.zeros((n, h, w, 8), dtype=torch.bfloat16, device=cuda)
    xs = torch.zeros((n, h // 2, w // 2, 32), dtype=torch.bfloat16, device=cuda)
    x8s, xss = T(x8), T(xs)
    stp.stem_prep(img.data_ptr(), n, h, w, 3, coef.data_ptr(), ref(x8s), stream())
    stp.stem_prep(img.data_ptr(), n, h, w, 3, coef.data_ptr(), ref(xss), stream())
    # space-to-depth layout check
    x8c = x8.float().cpu()
    s2d = x8c.view(n, h // 2, 2, w // 2, 2, 8).permute(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 32)
    assert torch.equal(xs.float().cpu(), s2d)
    wm = torch.zeros((cout, 7, 7, 8))
    wm[..., :4] = bf16_round(torch.randn((cout, 7, 7, 4), generator=g) * 0.08)
    wm = wm.to(cuda)
    w2 = torch.zeros(cout * 16 * 32, dtype=torch.bfloat16, device=cuda)
    stp.stem_weight_s2d(wm.data_ptr(), w2.data_ptr(), cout, stream())
    desc = lib.ConvDesc(4, 4, 1, 2, 2, 1, 0)
    y = torch.zeros((n, h // 2, w // 2, cout), dtype=torch.bfloat16, device=cuda)
    ys = T(y)
    tc0 = stp.tc_launch_count()
    stp.conv_fwd(C.byref(desc), ref(xss), w2.data_ptr(), None, None, ref(ys), None, 0, stream())
    assert stp.tc_launch_count() == tc0 + 1
    xr = x8c.clone().requires_grad_(True)
    wr = wm.cpu().clone().requires_grad_(True)
    yr = conv_ref_autograd(xr, wr, 2, 3, 1, (h // 2, w // 2))
    assert rel_err(y, yr) < TOL_BF16
    dy = rand_bf16((n, h // 2, w // 2, cout), g)
    yr.backward(dy.float().cpu())
    dys = T(dy)
    ws = _ws(stp.conv_wgrad_workspace(C.byref(desc), ref(xss), ref(dys)), cuda)
    dw2 = torch.zeros(cout * 16 * 32, device=cuda)
    dw = torch.zeros((cout, 7, 7, 8), device=cuda)
    stp.conv_wgrad(C.byref(desc), ref(xss), ref(dys), dw2.data_ptr(), ws.data_ptr(), ws.numel(), stream())
    stp.stem_wgrad_s2d_gather(dw2.data_ptr(), dw.data_ptr(), cout, stream())
    assert rel_err(dw, wr.grad) < TOL_F32


@pytest.mark.parametrize("case", [(2, 24, 40, 64, 64, 3, 1, 1), (1, 9, 17, 128, 256, 3, 1, 1), (2, 20, 36, 16, 16, 3, 1, 1),
                                  (1, 12, 24, 32, 32, 3, 1, 1), (2, 16, 16, 64, 128, 3, 2, 1), (3, 8, 8, 256, 512, 3, 1, 1),
                                  # 1x1 stride-1 layers on the halo kernel (plain GEMM over pixel strips): ResNet-50 bottleneck forms
                                  (2, 24, 40, 64, 256, 1, 1, 0), (1, 9, 17, 256, 64, 1, 1, 0), (3, 16, 16, 512, 128, 1, 1, 0)])
def test_conv_fwd_bn_epilogue_statistics(stp, cuda, case):
    """stp_conv_fwd_bn (statistics accumulated in the conv epilogue, or the fallback pass) == conv + stp_bn_stats_fused,
    with residual, partial tiles, several N tiles; accumulators and ticket return to zero (replayable)."""
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(cin + cout)
    x = rand_bf16((n, h, w, cin), g)
    wt = rand_bf16((cout, k, k, cin), g, scale=1.0 / math.sqrt(k * k * cin))
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    res = rand_bf16((n, ho, wo, cout), g)
    desc = lib.ConvDesc(k, k, stride, pad, pad, 1, 0)
    rows = n * ho * wo
    nblk = stp.bn_nblk(rows, cout)
    partial = torch.zeros(2 * nblk * cout, device=cuda)
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * cout, dtype=torch.float64, device=cuda)
    gamma = (torch.rand(cout, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(cout, generator=g) * 0.2).to(cuda)
    xs, rs = T(x), T(res)
    # reference: plain conv, then the one-launch statistics kernel
    y0 = torch.zeros((n, ho, wo, cout), dtype=torch.bfloat16, device=cuda)
    y0s = T(y0)
    coef0 = torch.zeros(4 * cout, device=cuda)
    mm0, mv0 = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
    stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(rs), ref(y0s), None, 0, stream())
    stp.bn_stats_fused(ref(y0s), partial.data_ptr(), sync.data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                       mm0.data_ptr(), mv0.data_ptr(), coef0.data_ptr(), stream())
    for _ in range(2):
        y1 = torch.zeros_like(y0)
        y1s = T(y1)
        coef1 = torch.zeros(4 * cout, device=cuda)
        mm1, mv1 = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
        bn = lib.BnFwd(partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                       mm1.data_ptr(), mv1.data_ptr(), coef1.data_ptr())
        stp.conv_fwd_bn(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(rs), ref(y1s), C.byref(bn), None, 0, stream())
        assert torch.equal(y1, y0)
        assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
        scale = 1 + float(coef0.abs().max())
        assert max_abs(coef1, coef0) <= 2e-6 * scale, max_abs(coef1, coef0)
        assert max_abs(mm1, mm0) < 1e-6 and rel_err(mv1, mv0) < 1e-6


@pytest.mark.parametrize("cl", [2, 4])
@pytest.mark.parametrize("case", [(2, 24, 40, 128, 128), (1, 9, 17, 128, 256), (3, 8, 8, 256, 512), (5, 16, 16, 64, 128),
                                  (16, 32, 32, 256, 256)])
def test_conv_tc2_cluster_multicast(stp, cuda, case, cl):
    """thread-block-cluster variants (weights TMA-multicast to CL CTAs): same results as the single-CTA kernel, incl.
    ranks that idle on an out-of-range image, residual and fused BatchNorm statistics."""
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(cin * 3 + cout + cl)
    x = rand_bf16((n, h, w, cin), g)
    wt = rand_bf16((cout, 3, 3, cin), g, scale=1.0 / math.sqrt(9 * cin))
    res = rand_bf16((n, h, w, cout), g)
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
    rows = n * h * w
    partial = torch.zeros(2 * stp.bn_nblk(rows, cout) * cout, device=cuda)
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * cout, dtype=torch.float64, device=cuda)
    gamma, beta = torch.ones(cout, device=cuda), torch.zeros(cout, device=cuda)
    xs, rs = T(x), T(res)
    outs = []
    stp.set_option(b"tc3", 1)  # this test is about the conv_tc2 cluster variants: keep the CTA-pair kernel out of the way
    try:
        for c in (1, cl):
            stp.set_option(b"tc2_cluster", c)
            y = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
            ys = T(y)
            coef = torch.zeros(4 * cout, device=cuda)
            mm, mv = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
            bn = lib.BnFwd(partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                           mm.data_ptr(), mv.data_ptr(), coef.data_ptr())
            for _ in range(2):
                stp.conv_fwd_bn(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(rs), ref(ys), C.byref(bn), None, 0, stream())
            torch.cuda.synchronize()
            outs.append((y, coef))
    finally:
        stp.set_option(b"tc2_cluster", 0)
        stp.set_option(b"tc3", 0)
    assert torch.equal(outs[0][0], outs[1][0])
    assert max_abs(outs[0][1], outs[1][1]) <= 2e-6 * (1 + float(outs[0][1].abs().max()))
    yr = conv_ref(x, wt, 1, 1) + res.float().cpu()
    assert rel_err(outs[1][0], yr) < TOL_BF16


@pytest.mark.parametrize("act", ["elu", "relu"])
@pytest.mark.parametrize("ties", [False, True])
def test_lovasz_hinge(stp, cuda, act, ties):
    """Lovasz hinge (per image, on logits) vs the oracle evaluated in float64; `ties` quantises the logits so that
    many errors are equal (stable descending order must match torch.sort(stable=True)); image 1 has an empty mask."""
    from oracle import losses as OL
    g = torch.Generator().manual_seed(21)
    n, h, w = 3, 40, 36
    logits = torch.randn(n, h, w, 1, generator=g) * 2.0
    if ties:
        logits = (logits * 2).round() / 2
    mask = (torch.rand(n, h, w, 1, generator=g) > 0.6).to(torch.uint8)
    mask[1] = 0
    lg = logits.to(cuda).contiguous()
    mk = mask.to(cuda).contiguous()
    ws = _ws(stp.lovasz_workspace(n, h * w), cuda)
    result = torch.zeros(16, device=cuda)
    result[lib.L_LOSS] = 0.25
    dl = torch.full((n * h * w,), 0.5, device=cuda)
    stp.lovasz_fwd(lg.data_ptr(), mk.data_ptr(), n, h * w, int(act == "elu"), 2.0, 1, ws.data_ptr(), ws.numel(), result.data_ptr(),
                   stream())
    stp.lovasz_bwd(ws.data_ptr(), ws.numel(), n, h * w, 2.0, 1, dl.data_ptr(), stream())
    z = logits.double().requires_grad_(True)
    lo = OL.lovasz_loss(mask.double(), z, act=act)
    lo.backward()
    assert abs(float(result[lib.L_LOVASZ]) - float(lo)) < 1e-5 * max(1.0, abs(float(lo)))
    assert abs(float(result[lib.L_LOSS]) - (0.25 + 2.0 * float(lo))) < 1e-5 * max(1.0, abs(float(lo)))
    assert rel_err(dl.view(n, h, w, 1) - 0.5, 2.0 * z.grad) < 1e-5


def test_adam_device_lr_scale(stp, cuda):
    """stp_grad_xform.d_lr_scale: the captured optimizer launch follows a learning rate changed on the device."""
    n = 1024
    g = torch.Generator().manual_seed(4)
    grad = torch.randn(n, generator=g).to(cuda)
    step = torch.zeros(1, dtype=torch.int64, device=cuda)
    outs = []
    for scale in (1.0, 0.25):
        p = torch.zeros(n, device=cuda)
        m, v = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
        sc = torch.full((1,), scale, device=cuda)
        gx = lib.GradXform(1.0, 0.0, 0.0, None, sc.data_ptr())
        stp.adam(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-2, 0.9, 0.999, 1e-7, C.byref(gx),
                 step.data_ptr(), stream())
        outs.append(p.clone())
    assert rel_err(outs[1], 0.25 * outs[0]) < 1e-6


@pytest.mark.parametrize("shape,rate", [((2, 4, 4, 16), 8), ((1, 8, 6, 32), 4), ((2, 16, 12, 8), 2), ((1, 5, 7, 8), 3)])
def test_resize_bilinear_bf16(stp, cuda, shape, rate):
    """TF1-legacy bilinear upsampling (FPN branches) into a channel slice of a wider buffer; backward = exact adjoint
    (checked against autograd of the oracle restatement), residual accumulation."""
    from oracle import nn as ON
    g = torch.Generator().manual_seed(31)
    n, h, w, c = shape
    H, W = h * rate, w * rate
    x = rand_bf16(shape, g, device=cuda)
    ybuf = torch.zeros((n, H, W, c + 16), dtype=torch.bfloat16, device=cuda)
    stp.resize_bilinear_fwd(ref(T(x)), ref(T(ybuf, 8, c)), stream())
    xo = x.float().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    yo = ON.resize_bilinear_tf1(xo, H, W)
    got = ybuf[..., 8:8 + c].float().cpu()
    assert rel_err(got, yo.detach().permute(0, 2, 3, 1)) < TOL_BF16  # one bf16 rounding (fma contraction may move the last fp32 bit)
    assert float(ybuf[..., :8].abs().max()) == 0.0 and float(ybuf[..., 8 + c:].abs().max()) == 0.0
    dybuf = torch.zeros_like(ybuf)
    dy = rand_bf16((n, H, W, c), g, device=cuda)
    dybuf[..., 8:8 + c] = dy
    res = rand_bf16(shape, g, device=cuda)
    for use_res in (False, True):
        dx = torch.zeros(shape, dtype=torch.bfloat16, device=cuda)
        stp.resize_bilinear_bwd(ref(T(dybuf, 8, c)), ref(T(res)) if use_res else None, ref(T(dx)), stream())
        (gx,) = torch.autograd.grad(yo, xo, dy.float().cpu().permute(0, 3, 1, 2), retain_graph=True)
        want = gx.permute(0, 2, 3, 1) + (res.float().cpu() if use_res else 0.0)
        assert rel_err(dx, want) < TOL_BF16


def test_resize_bilinear_f32_logits(stp, cuda):
    """x4 `last_upsample` of the padded FPN head logits: f32 [n,h,w,16] -> dense f32 [n,4h,4w,3]; backward f32 -> bf16 with
    the padded channels written as zero."""
    from oracle import nn as ON
    g = torch.Generator().manual_seed(32)
    n, h, w, cp, cls = 2, 8, 6, 16, 3
    x = torch.randn(n, h, w, cp, generator=g).to(cuda)
    y = torch.zeros(n, 4 * h, 4 * w, cls, device=cuda)
    stp.resize_bilinear_fwd(ref(T(x)), ref(T(y)), stream())
    xo = x[..., :cls].cpu().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    yo = ON.resize_bilinear_tf1(xo, 4 * h, 4 * w)
    assert max_abs(y, yo.detach().permute(0, 2, 3, 1)) < 1e-6
    dy = torch.randn(n, 4 * h, 4 * w, cls, generator=g).to(cuda)
    dx = torch.full((n, h, w, cp), 7.0, dtype=torch.bfloat16, device=cuda)
    stp.resize_bilinear_bwd(ref(T(dy)), None, ref(T(dx)), stream())
    (gx,) = torch.autograd.grad(yo, xo, dy.cpu().permute(0, 3, 1, 2))
    assert rel_err(dx[..., :cls], gx.permute(0, 2, 3, 1)) < TOL_BF16
    assert float(dx[..., cls:].float().abs().max()) == 0.0


def test_upsample2x_bwd(stp, cuda):
    g = torch.Generator().manual_seed(33)
    n, h, w, c = 2, 6, 10, 24
    dy = rand_bf16((n, 2 * h, 2 * w, c), g, device=cuda)
    res = rand_bf16((n, h, w, c), g, device=cuda)
    dx = torch.zeros((n, h, w, c), dtype=torch.bfloat16, device=cuda)
    want = dy.float().cpu().view(n, h, 2, w, 2, c).sum(dim=(2, 4))
    stp.upsample2x_bwd(ref(T(dy)), None, ref(T(dx)), stream())
    assert rel_err(dx, want) < TOL_BF16
    stp.upsample2x_bwd(ref(T(dy)), ref(T(res)), ref(T(dx)), stream())
    assert rel_err(dx, want + res.float().cpu()) < TOL_BF16


def test_lovasz_hinge_multiclass(stp, cuda):
    """classes > 1 ([image][pixel][class] logits): one hinge per (image, class), mean over all -- the oracle's definition
    of the case the reference leaves undefined."""
    from oracle import losses as OL
    g = torch.Generator().manual_seed(22)
    n, h, w, cls = 2, 24, 20, 3
    logits = ((torch.randn(n, h, w, cls, generator=g) * 2.0) * 2).round() / 2
    mask = (torch.rand(n, h, w, cls, generator=g) > 0.6).to(torch.uint8)
    mask[1, :, :, 2] = 0
    lg, mk = logits.to(cuda).contiguous(), mask.to(cuda).contiguous()
    ws = _ws(stp.lovasz_workspace(n * cls, h * w), cuda)
    result = torch.zeros(16, device=cuda)
    dl = torch.zeros(n * h * w * cls, device=cuda)
    stp.lovasz_fwd_mc(lg.data_ptr(), mk.data_ptr(), n, h * w, cls, 1, 1.0, 0, ws.data_ptr(), ws.numel(), result.data_ptr(), stream())
    stp.lovasz_bwd(ws.data_ptr(), ws.numel(), n * cls, h * w, 1.0, 0, dl.data_ptr(), stream())
    z = logits.double().requires_grad_(True)
    lo = OL.lovasz_loss(mask.double(), z, act="elu")
    lo.backward()
    assert abs(float(result[lib.L_LOVASZ]) - float(lo)) < 1e-5 * max(1.0, abs(float(lo)))
    assert rel_err(dl.view(n, h, w, cls), z.grad) < 1e-5


@pytest.mark.parametrize("classes", [2, 3, 4])
def test_softmax_categorical_crossentropy(stp, cuda, classes):
    """`activation: softmax` + `loss: categorical_crossentropy` (schema segmentation.raml:12-21, 62-63): keras
    categorical_crossentropy on probabilities (renormalise, clip 1e-7, -sum t log p, mean) fused with the softmax; value,
    categorical accuracy and dL/dlogits against autograd of the oracle formula, incl. saturated logits (clip region) and the
    accumulate form."""
    from oracle import losses as OL
    g = torch.Generator().manual_seed(100 + classes)
    n, h, w = 2, 20, 28
    logits = torch.randn(n, h, w, classes, generator=g) * 3
    logits[0, 0, 0] = torch.tensor([30.0] + [-30.0] * (classes - 1))      # p saturates: clipped, zero gradient there
    logits[0, 0, 1] = torch.tensor([-30.0] * (classes - 1) + [30.0])
    cls = torch.randint(0, classes, (n, h, w), generator=g)
    mask = torch.nn.functional.one_hot(cls, classes).to(torch.uint8)
    mask[1, 3, 3] = 0                                                     # an unlabeled pixel (all-zero target row)
    lg, mk = logits.to(cuda).contiguous(), mask.to(cuda).contiguous()
    partial = torch.zeros(stp.loss_partial_floats(), device=cuda)
    result = torch.zeros(16, device=cuda)
    result[lib.L_LOSS] = 0.25
    wgt = 0.7
    stp.softmax_cce_fwd(lg.data_ptr(), mk.data_ptr(), n * h * w, classes, wgt, 1, partial.data_ptr(), result.data_ptr(), stream())
    z = logits.double().requires_grad_(True)
    lo = OL.categorical_crossentropy(mask.double(), torch.softmax(z, dim=-1))
    lo.backward()
    assert abs(float(result[lib.L_CCE]) - float(lo)) < 1e-5 * max(1.0, abs(float(lo)))
    assert abs(float(result[lib.L_LOSS]) - (0.25 + wgt * float(lo))) < 1e-5 * max(1.0, abs(float(lo)))
    acc = float((logits.argmax(-1) == mask.argmax(-1)).float().mean())
    assert abs(float(result[lib.L_CACC]) - acc) < 1e-6
    dl = torch.ones(n * h * w * classes, device=cuda)
    stp.softmax_cce_bwd(lg.data_ptr(), mk.data_ptr(), n * h * w, classes, wgt, 1, dl.data_ptr(), stream())
    assert rel_err(dl.view(n, h, w, classes) - 1.0, (wgt * z.grad).float()) < 1e-4
    stp.softmax_cce_bwd(lg.data_ptr(), mk.data_ptr(), n * h * w, classes, 1.0, 0, dl.data_ptr(), stream())
    assert rel_err(dl.view(n, h, w, classes), z.grad.float()) < 1e-4


@pytest.mark.parametrize("halo", [0, 1])
@pytest.mark.parametrize("force", [(0, 0), (128, 1), (128, 2), (256, 1)])
@pytest.mark.parametrize("case", [(2, 24, 40, 128, 128), (1, 9, 17, 128, 256), (3, 8, 8, 256, 512), (5, 16, 16, 64, 128),
                                  (16, 32, 32, 256, 256), (1, 16, 16, 192, 256), (3, 24, 24, 64, 64), (2, 16, 40, 128, 192)])
def test_conv_tc3_cta_pair(stp, cuda, case, force, halo):
    """tcgen05 cta_group::2 CTA-pair kernel (conv_tc3.cu) against the single-CTA halo kernel and the fp32 reference:
    odd numbers of pixel tiles (a rank idling on an out-of-range image), partial tiles, residual, fused BatchNorm
    statistics, every (BN, MT) specialisation, dgrad through the same kernel.  halo = 1 (option tc3_halo): ONE haloed A box per
    channel block, the three filter columns are 1-pixel shifted descriptor views; halo = 0 (default): one box per filter column."""
    n, h, w, cin, cout = case
    fbn, fmt = force
    if fbn == 256 and cout % 256:
        pytest.skip("BN=256 needs Cout % 256 == 0")
    if cout % 128:   # Cout = 64 / 192: the N = 64 pair tiles (option tc3_bn64)
        if fbn in (128, 256):
            pytest.skip("Cout % 128 != 0 is served by BN = 64 only")
        stp.set_option(b"tc3_bn64", 1)
    g = torch.Generator().manual_seed(cin * 5 + cout + fbn + fmt)
    x = rand_bf16((n, h, w, cin), g)
    wt = rand_bf16((cout, 3, 3, cin), g, scale=1.0 / math.sqrt(9 * cin))
    res = rand_bf16((n, h, w, cout), g)
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
    rows = n * h * w
    partial = torch.zeros(2 * stp.bn_nblk(rows, cout) * cout, device=cuda)
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * cout, dtype=torch.float64, device=cuda)
    gamma, beta = torch.ones(cout, device=cuda), torch.zeros(cout, device=cuda)
    xs, rs = T(x), T(res)
    outs = []
    try:
        for mode in (1, 2):   # 1: single-CTA halo kernel, 2: CTA-pair kernel forced on
            stp.set_option(b"tc3", mode)
            stp.set_option(b"tc3_halo", 1 if halo else 0)
            stp.set_option(b"tc3_force_bn", fbn)
            stp.set_option(b"tc3_force_mt", fmt)
            before = stp.tc_launch_count()
            y = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
            ys = T(y)
            coef = torch.zeros(4 * cout, device=cuda)
            mm, mv = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
            bn = lib.BnFwd(partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                           mm.data_ptr(), mv.data_ptr(), coef.data_ptr())
            for _ in range(2):
                stp.conv_fwd_bn(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(rs), ref(ys), C.byref(bn), None, 0, stream())
            # plain conv (no residual / statistics)
            y2 = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
            stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(T(y2)), None, 0, stream())
            torch.cuda.synchronize()
            assert stp.tc_launch_count() - before == 3
            outs.append((y, coef, y2))
    finally:
        stp.set_option(b"tc3", 0)
        stp.set_option(b"tc3_halo", 0)
        stp.set_option(b"tc3_bn64", 0)
        stp.set_option(b"tc3_force_bn", 0)
        stp.set_option(b"tc3_force_mt", 0)
    yr = conv_ref(x, wt, 1, 1)
    assert rel_err(outs[1][2], yr) < TOL_BF16
    assert rel_err(outs[1][0], yr + res.float().cpu()) < TOL_BF16
    # same K order, fp32 accumulation in TMEM: the pair kernel reproduces the single-CTA kernel up to bf16 rounding ties
    assert rel_err(outs[1][0], outs[0][0].float()) < 1e-3 and rel_err(outs[1][2], outs[0][2].float()) < 1e-3
    assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
    scale = 1 + float(outs[0][1].abs().max())
    assert max_abs(outs[1][1], outs[0][1]) <= 1e-3 * scale


@pytest.mark.parametrize("relu", [1, 0])
@pytest.mark.parametrize("case", [(2, 24, 40, 64, 64), (1, 16, 16, 128, 128), (2, 32, 64, 16, 16), (1, 20, 36, 32, 32),
                                  (3, 9, 17, 64, 128), (2, 16, 16, 256, 64), (1, 40, 24, 32, 16)])
def test_conv_dgrad_with_fused_bn_backward_reduce(stp, cuda, case, relu):
    """stp_conv_dgrad_bn: the BatchNorm-backward reduction of the layer that produced the conv's input, done in the dgrad
    epilogue (dx stored ReLU-masked, sums of g and g*xhat -> dgamma, dbeta, bcoef), against dgrad + the separate reduction
    pass; partial tiles, narrow (register-accumulated) and wide (per-chunk transpose-reduced) channel tiles."""
    n, h, w, cin, cout = case   # forward conv: cin -> cout; the BatchNorm under test has cin channels
    g = torch.Generator().manual_seed(cin * 7 + cout + relu)
    x = rand_bf16((n, h, w, cin), g)                       # BatchNorm input
    dy = rand_bf16((n, h, w, cout), g)                     # gradient of the conv output
    wt = rand_bf16((cout, 3, 3, cin), g, scale=1.0 / math.sqrt(9 * cin))
    wd = torch.zeros_like(wt).view(-1)
    wf = torch.zeros_like(wt).view(-1)
    stp.weight_prep(wt.float().contiguous().data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, 3, 3, cin, stream())
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
    rows = n * h * w
    gamma = (torch.rand(cin, generator=g) + 0.5).to(cuda)
    gamma[::3] *= -1.0                                       # negative scales flip the ReLU mask condition
    beta = (torch.rand(cin, generator=g) - 0.5).to(cuda)
    partial = torch.zeros(2 * stp.bn_nblk(rows, cin) * cin, device=cuda)
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * max(cin, cout), dtype=torch.float64, device=cuda)
    coef = torch.zeros(4 * cin, device=cuda)
    mm, mv = torch.zeros(cin, device=cuda), torch.ones(cin, device=cuda)
    stp.bn_stats_fused(ref(T(x)), partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                       mm.data_ptr(), mv.data_ptr(), coef.data_ptr(), stream())
    outs = []
    xs = T(x)
    for fused in (0, 1):
        dx = torch.zeros((n, h, w, cin), dtype=torch.bfloat16, device=cuda)
        dgam, dbet, bco = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda), torch.zeros(3 * cin, device=cuda)
        if fused:
            bnb = lib.BnBwd(C.pointer(xs), coef.data_ptr(), relu, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                            dgam.data_ptr(), dbet.data_ptr(), bco.data_ptr())
            before = stp.launch_count()
            for _ in range(2):   # twice: the accumulators / ticket must come back to zero
                stp.conv_dgrad_bn(C.byref(desc), ref(T(dy)), wd.data_ptr(), ref(T(dx)), C.byref(bnb), None, 0, stream())
            assert stp.launch_count() - before == 2, "the reduction must run inside the dgrad kernel for these shapes"
        else:
            stp.conv_dgrad(C.byref(desc), ref(T(dy)), wd.data_ptr(), None, ref(T(dx)), None, 0, stream())
            stp.bn_bwd_reduce_fused(ref(T(dx)), ref(xs), coef.data_ptr(), relu, 1, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                                    dgam.data_ptr(), dbet.data_ptr(), bco.data_ptr(), stream())
        torch.cuda.synchronize()
        outs.append((dx, dgam, dbet, bco))
    assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
    c4 = coef.view(4, cin)
    mask = (x.float() * c4[2] + c4[3] > 0) if relu else torch.ones_like(x, dtype=torch.bool)
    assert torch.equal(outs[1][0], torch.where(mask, outs[0][0], torch.zeros_like(outs[0][0])))
    for k in (1, 2, 3):
        scale = 1e-6 + float(outs[0][k].abs().max())
        assert max_abs(outs[1][k], outs[0][k]) <= 2e-4 * scale, (k, max_abs(outs[1][k], outs[0][k]), scale)


G1_CASES = [
    # n, h, w, cin, cout: 1x1 stride-1 convolutions as plain GEMMs (csrc/gemm1x1.cu)
    (2, 20, 20, 96, 576),     # MobileNetV2 expand, K = 3 x 32
    (1, 20, 20, 160, 960),    # N = 7.5 x 128: partial last N tile
    (1, 9, 7, 24, 144),       # K tail (24 = 16 + 8), M = 63 (partial M tile)
    (1, 9, 7, 144, 24),       # N = 24 < 64
    (2, 16, 16, 728, 728),    # Xception middle flow
    (1, 12, 12, 16, 96),      # the tiny-K case the tcgen05 kernel loses
    (2, 32, 32, 64, 128),     # a shape the tcgen05 kernel also serves (option gemm1x1 = 2 forces this kernel)
    (1, 8, 8, 960, 320),
]


@pytest.mark.parametrize("case", G1_CASES)
def test_gemm1x1(stp, cuda, case):
    """forward (+ residual, + bias + ReLU), forward into / out of channel slices (ld > c), and dgrad (plain and accumulate in
    place) of 1x1 stride-1 convolutions on the streaming GEMM kernel, against fp32 math on the same bf16 operands"""
    n, h, w, cin, cout = case
    stp.set_option(b"gemm1x1", 2)
    try:
        g = torch.Generator().manual_seed(sum(case))
        x = rand_bf16((n, h, w, cin), g)
        wt = rand_bf16((cout, 1, 1, cin), g, scale=1.0 / math.sqrt(cin))
        res = rand_bf16((n, h, w, cout), g)
        desc = lib.ConvDesc(1, 1, 1, 0, 0, 1, 0)
        xs, rs = T(x), T(res)
        y = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
        ys = T(y)
        tc0, l0 = stp.tc_launch_count(), stp.launch_count()
        stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(rs), ref(ys), None, 0, stream())
        assert stp.tc_launch_count() == tc0 and stp.launch_count() == l0 + 1
        yr = conv_ref(x, wt, 1, 0, 1, (h, w))
        assert rel_err(y, yr + res.float().cpu()) < TOL_BF16
        # bias + ReLU (keras Conv2D(activation='relu') of the decoders without BatchNorm)
        bias = torch.randn(cout, generator=g).to(cuda)
        drelu = lib.ConvDesc(1, 1, 1, 0, 0, 1, lib.CONV_RELU)
        stp.conv_fwd(C.byref(drelu), ref(xs), wt.data_ptr(), bias.data_ptr(), None, ref(ys), None, 0, stream())
        assert rel_err(y, torch.relu(yr + bias.cpu())) < TOL_BF16
        # channel slices of wider buffers on both sides
        xbig = rand_bf16((n, h, w, cin + 16), g)
        ybig = torch.zeros((n, h, w, cout + 24), dtype=torch.bfloat16, device=cuda)
        xbs, ybs = T(xbig, 8, cin), T(ybig, 16, cout)
        stp.conv_fwd(C.byref(desc), ref(xbs), wt.data_ptr(), None, None, ref(ybs), None, 0, stream())
        assert rel_err(ybig[..., 16:16 + cout], conv_ref(xbig[..., 8:8 + cin], wt, 1, 0, 1, (h, w))) < TOL_BF16
        assert float(ybig[..., :16].abs().max()) == 0 and float(ybig[..., 16 + cout:].abs().max()) == 0
        # dgrad = the same GEMM over dy with the [Cin][Cout] weight copy
        wf, wd = torch.zeros_like(wt), torch.zeros_like(wt)
        stp.weight_prep(wt.float().contiguous().data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, 1, 1, cin, stream())
        dy = rand_bf16((n, h, w, cout), g)
        dx = torch.zeros_like(x)
        dys, dxs = T(dy), T(dx)
        stp.conv_dgrad(C.byref(desc), ref(dys), wd.data_ptr(), None, ref(dxs), None, 0, stream())
        dxr = torch.einsum("nhwo,oi->nhwi", dy.float().cpu(), wt.float().cpu().view(cout, cin))
        assert rel_err(dx, dxr) < TOL_BF16
        dx2 = x.clone()
        dx2s = T(dx2)
        stp.conv_dgrad(C.byref(desc), ref(dys), wd.data_ptr(), ref(dx2s), ref(dx2s), None, 0, stream())
        assert rel_err(dx2, dxr + x.float().cpu()) < TOL_BF16
        # ---- BatchNorm epilogues (forward statistics; fused backward masking + reduction with ReLU / ReLU6 / no activation) ----
        rows = n * h * w
        cmax = max(cin, cout)
        partial = torch.zeros(2 * stp.bn_nblk(rows, cmax) * cmax, device=cuda)
        sync = torch.zeros(4, dtype=torch.int32, device=cuda)
        acc = torch.zeros(2 * cmax, dtype=torch.float64, device=cuda)
        gamma = (torch.rand(cout, generator=g) + 0.5).to(cuda)
        beta = (torch.randn(cout, generator=g) * 0.2).to(cuda)
        y0 = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
        stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(T(y0)), None, 0, stream())
        coef0 = torch.zeros(4 * cout, device=cuda)
        mm0, mv0 = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
        stp.bn_stats_fused(ref(T(y0)), partial.data_ptr(), sync.data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                           mm0.data_ptr(), mv0.data_ptr(), coef0.data_ptr(), stream())
        for _ in range(2):   # twice: accumulators / ticket return to zero
            y1 = torch.zeros_like(y0)
            coef1 = torch.zeros(4 * cout, device=cuda)
            mm1, mv1 = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
            bn = lib.BnFwd(partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                           mm1.data_ptr(), mv1.data_ptr(), coef1.data_ptr())
            l0 = stp.launch_count()
            stp.conv_fwd_bn(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(T(y1)), C.byref(bn), None, 0, stream())
            # one launch (tcgen05 halo kernel where it tiles the shape) or GEMM + its 1-block finalize -- never a pass over y
            assert stp.launch_count() - l0 in (1, 2)
            assert rel_err(y1, y0.float()) < 1e-3   # (same math; fp32 summation order differs between the two kernels)
            assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
            scale = 1 + float(coef0.abs().max())
            assert max_abs(coef1, coef0) <= 1e-3 * scale, max_abs(coef1, coef0)
            assert max_abs(mm1, mm0) < 1e-5 and rel_err(mv1, mv0) < 1e-4
        # BatchNorm(+activation) over x (cin channels) feeding this conv: its backward reduction inside the dgrad GEMM
        gam = (torch.rand(cin, generator=g) + 0.5).to(cuda)
        gam[::3] *= -1.0
        bet = (torch.rand(cin, generator=g) - 0.5).to(cuda)
        coef = torch.zeros(4 * cin, device=cuda)
        mm, mv = torch.zeros(cin, device=cuda), torch.ones(cin, device=cuda)
        stp.bn_stats_fused(ref(xs), partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1e-3, 0.99,
                           mm.data_ptr(), mv.data_ptr(), coef.data_ptr(), stream())
        c4 = coef.view(4, cin)
        tpre = x.float() * c4[2] + c4[3]
        for act in (1, 2, 0):
            d0, b0, c0 = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda), torch.zeros(3 * cin, device=cuda)
            stp.conv_dgrad(C.byref(desc), ref(dys), wd.data_ptr(), None, ref(dxs), None, 0, stream())
            stp.bn_bwd_reduce_fused(ref(dxs), ref(xs), coef.data_ptr(), act, 1, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                                    d0.data_ptr(), b0.data_ptr(), c0.data_ptr(), stream())
            d1, b1, c1 = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda), torch.zeros(3 * cin, device=cuda)
            dxm = torch.zeros_like(x)
            bnb = lib.BnBwd(C.pointer(xs), coef.data_ptr(), act, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                            d1.data_ptr(), b1.data_ptr(), c1.data_ptr())
            l0 = stp.launch_count()
            for _ in range(2):
                stp.conv_dgrad_bn(C.byref(desc), ref(dys), wd.data_ptr(), ref(T(dxm)), C.byref(bnb), None, 0, stream())
            assert stp.launch_count() - l0 in (2, 4), "the reduction must run inside the dgrad kernel (+ its 1-block finalize)"
            torch.cuda.synchronize()
            assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
            mask = (tpre > 0) if act == 1 else ((tpre > 0) & (tpre <= 6)) if act == 2 else torch.ones_like(tpre, dtype=torch.bool)
            assert rel_err(dxm, torch.where(mask, dx, torch.zeros_like(dx)).float()) < 1e-3
            for got, want in ((d1, d0), (b1, b0), (c1, c0)):
                scale = 1e-6 + float(want.abs().max())
                assert max_abs(got, want) <= 1e-3 * scale, (act, max_abs(got, want), scale)
    finally:
        stp.set_option(b"gemm1x1", 0)


NC_CASES = [
    # n, h, w, cin, cout: 3x3 stride-1 'same' convolutions of the decoder tail on the narrow-channel kernel (csrc/conv_narrow.cu)
    (4, 100, 300, 16, 16),    # 520 tiles > 2 x 148 CTAs: the persistent loop, both ring stages, partial right / bottom tiles
    (2, 37, 70, 32, 16),
    (2, 40, 33, 16, 32),
    (3, 64, 96, 32, 32),
    (1, 5, 9, 16, 16),        # smaller than one tile
]


@pytest.mark.parametrize("case", NC_CASES)
def test_conv_narrow(stp, cuda, case):
    """forward (plain, bias + ReLU, into / out of channel slices), forward with the BatchNorm-statistics epilogue and dgrad on the
    narrow-channel mma.sync kernel (opt-in, option nconv = 1: measured slower than the tcgen05 halo kernel on every shape,
    profiles/r2_s9_nconv_bench.txt): against fp32 math on the same bf16 operands, and against the halo kernel (nconv = 0)"""
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(sum(case))
    x = rand_bf16((n, h, w, cin), g)
    wt = rand_bf16((cout, 3, 3, cin), g, scale=1.0 / math.sqrt(9 * cin))
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
    xs = T(x)
    yr = conv_ref(x, wt, 1, 1)
    ys_by_path = []
    for on in (1, 0):
        stp.set_option(b"nconv", on)
        y = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
        tc0, l0 = stp.tc_launch_count(), stp.launch_count()
        stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(T(y)), None, 0, stream())
        assert stp.launch_count() == l0 + 1 and stp.tc_launch_count() - tc0 == 1 - on   # which kernel served it
        assert rel_err(y, yr) < TOL_BF16
        ys_by_path.append(y)
    assert rel_err(ys_by_path[0], ys_by_path[1].float()) < 1e-3
    stp.set_option(b"nconv", 1)
    try:
        _conv_narrow_rest(stp, cuda, case, g, x, wt, desc, xs, yr, ys_by_path[0])
    finally:
        stp.set_option(b"nconv", 0)


def _conv_narrow_rest(stp, cuda, case, g, x, wt, desc, xs, yr, y_narrow):
    n, h, w, cin, cout = case
    # bias + ReLU
    bias = torch.randn(cout, generator=g).to(cuda)
    drelu = lib.ConvDesc(3, 3, 1, 1, 1, 1, lib.CONV_RELU)
    y = torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=cuda)
    stp.conv_fwd(C.byref(drelu), ref(xs), wt.data_ptr(), bias.data_ptr(), None, ref(T(y)), None, 0, stream())
    assert rel_err(y, torch.relu(yr + bias.cpu())) < TOL_BF16
    # channel slices of wider buffers on both sides (zero-copy concat)
    xbig = rand_bf16((n, h, w, cin + 24), g)
    ybig = torch.zeros((n, h, w, cout + 16), dtype=torch.bfloat16, device=cuda)
    l0 = stp.launch_count()
    stp.conv_fwd(C.byref(desc), ref(T(xbig, 16, cin)), wt.data_ptr(), None, None, ref(T(ybig, 8, cout)), None, 0, stream())
    assert stp.launch_count() == l0 + 1
    assert rel_err(ybig[..., 8:8 + cout], conv_ref(xbig[..., 16:16 + cin], wt, 1, 1)) < TOL_BF16
    assert float(ybig[..., :8].abs().max()) == 0 and float(ybig[..., 8 + cout:].abs().max()) == 0
    # BatchNorm statistics in the epilogue == conv + the one-launch statistics kernel (twice: accumulators / ticket return to zero)
    rows = n * h * w
    partial = torch.zeros(2 * stp.bn_nblk(rows, cout) * cout, device=cuda)
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * cout, dtype=torch.float64, device=cuda)
    gamma = (torch.rand(cout, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(cout, generator=g) * 0.2).to(cuda)
    y0 = y_narrow
    coef0 = torch.zeros(4 * cout, device=cuda)
    mm0, mv0 = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
    stp.bn_stats_fused(ref(T(y0)), partial.data_ptr(), sync.data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                       mm0.data_ptr(), mv0.data_ptr(), coef0.data_ptr(), stream())
    for _ in range(2):
        y1 = torch.zeros_like(y0)
        coef1 = torch.zeros(4 * cout, device=cuda)
        mm1, mv1 = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
        bn = lib.BnFwd(partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                       mm1.data_ptr(), mv1.data_ptr(), coef1.data_ptr())
        l0 = stp.launch_count()
        stp.conv_fwd_bn(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(T(y1)), C.byref(bn), None, 0, stream())
        assert stp.launch_count() == l0 + 1, "the statistics must come out of the conv kernel's epilogue"
        assert torch.equal(y1, y0)
        assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
        scale = 1 + float(coef0.abs().max())
        assert max_abs(coef1, coef0) <= 2e-6 * scale, max_abs(coef1, coef0)
        assert max_abs(mm1, mm0) < 1e-6 and rel_err(mv1, mv0) < 1e-6
    # dgrad = the same kernel over dy with the tap-flipped [Cin][3][3][Cout] weight copy
    wf, wd = torch.zeros_like(wt), torch.zeros_like(wt)
    stp.weight_prep(wt.float().contiguous().data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, 3, 3, cin, stream())
    dy = rand_bf16((n, h, w, cout), g)
    dx = torch.zeros_like(x)
    tc0 = stp.tc_launch_count()
    stp.conv_dgrad(C.byref(desc), ref(T(dy)), wd.data_ptr(), None, ref(T(dx)), None, 0, stream())
    assert stp.tc_launch_count() == tc0
    xr = x.float().cpu().requires_grad_(True)
    conv_ref_autograd(xr, wt.float().cpu(), 1, 1, 1, (h, w)).backward(dy.float().cpu())
    assert rel_err(dx, xr.grad) < TOL_BF16
    # fused BatchNorm-backward masking + reduction in the dgrad epilogue == dgrad, then the separate reduction pass
    gam = (torch.rand(cin, generator=g) + 0.5).to(cuda)
    gam[::3] *= -1.0
    bet = (torch.rand(cin, generator=g) - 0.5).to(cuda)
    rows = n * h * w
    partial = torch.zeros(2 * stp.bn_nblk(rows, max(cin, cout)) * max(cin, cout), device=cuda)
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * max(cin, cout), dtype=torch.float64, device=cuda)
    coef = torch.zeros(4 * cin, device=cuda)
    mm, mv = torch.zeros(cin, device=cuda), torch.ones(cin, device=cuda)
    stp.bn_stats_fused(ref(xs), partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1e-3, 0.99,
                       mm.data_ptr(), mv.data_ptr(), coef.data_ptr(), stream())
    d0, b0, c0 = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda), torch.zeros(3 * cin, device=cuda)
    stp.bn_bwd_reduce_fused(ref(T(dx)), ref(xs), coef.data_ptr(), 1, 1, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                            d0.data_ptr(), b0.data_ptr(), c0.data_ptr(), stream())
    d1, b1, c1 = torch.zeros(cin, device=cuda), torch.zeros(cin, device=cuda), torch.zeros(3 * cin, device=cuda)
    dxm = torch.zeros_like(x)
    bnb = lib.BnBwd(C.pointer(xs), coef.data_ptr(), 1, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                    d1.data_ptr(), b1.data_ptr(), c1.data_ptr())
    l0, tc0 = stp.launch_count(), stp.tc_launch_count()
    for _ in range(2):
        stp.conv_dgrad_bn(C.byref(desc), ref(T(dy)), wd.data_ptr(), ref(T(dxm)), C.byref(bnb), None, 0, stream())
    assert stp.launch_count() == l0 + 2 and stp.tc_launch_count() == tc0
    torch.cuda.synchronize()
    assert int(sync[0]) == 0 and float(acc.abs().max()) == 0.0
    c4 = coef.view(4, cin)
    mask = x.float() * c4[2] + c4[3] > 0
    assert torch.equal(dxm, torch.where(mask, dx, torch.zeros_like(dx)))
    for got, want in ((d1, d0), (b1, b0), (c1, c0)):
        scale = 1e-6 + float(want.abs().max())
        assert max_abs(got, want) <= 2e-4 * scale, (max_abs(got, want), scale)


@pytest.mark.parametrize("case", [(4, 80, 80, 24, 144), (4, 80, 80, 144, 24), (2, 33, 47, 96, 576), (1, 20, 20, 960, 160),
                                  (3, 40, 40, 192, 32), (1, 7, 9, 8, 8), (2, 64, 64, 728, 728), (16, 1, 1, 320, 256)])
def test_wgrad1x1(stp, cuda, case):
    """weight gradient of 1x1 stride-1 convolutions at MobileNetV2 / Xception widths on the pixel-reduction GEMM (csrc/wgrad1x1.cu):
    many pixel splits, partial channel tiles, a pixel count that is not a multiple of the stage depth, channel slices of wider
    buffers -- against fp32 math on the same bf16 operands; deterministic (two runs bit-identical)"""
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(sum(case))
    xbig = rand_bf16((n, h, w, cin + 8), g)
    dybig = rand_bf16((n, h, w, cout + 16), g)
    xs, dys = T(xbig, 8, cin), T(dybig, 0, cout)
    desc = lib.ConvDesc(1, 1, 1, 0, 0, 1, 0)
    ws = _ws(stp.conv_wgrad_workspace(C.byref(desc), ref(xs), ref(dys)), cuda)
    outs = []
    for _ in range(2):
        flat = torch.zeros(cout * cin + 1, dtype=torch.float32, device=cuda)
        dw = flat[1:].view(cout, 1, 1, cin)     # a 4-byte aligned destination, as inside a flat gradient buffer
        tc0, l0 = stp.tc_launch_count(), stp.launch_count()
        stp.conv_wgrad(C.byref(desc), ref(xs), ref(dys), dw.data_ptr(), ws.data_ptr(), ws.numel(), stream())
        torch.cuda.synchronize()
        assert float(flat[0]) == 0.0
        outs.append(dw)
    want = torch.einsum("nhwo,nhwi->oi", dybig[..., :cout].float().cpu().double(), xbig[..., 8:].float().cpu().double())
    if stp.tc_launch_count() == tc0:      # (shapes the tcgen05 wgrad kernel tiles keep it: nothing to check here)
        assert stp.launch_count() - l0 <= 2
    assert rel_err(outs[0].view(cout, cin), want.float()) < TOL_F32
    assert torch.equal(outs[0], outs[1])
