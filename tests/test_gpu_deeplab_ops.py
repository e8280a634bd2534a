"""GPU parity tests of the kernels the DeepLabV3+ / MobileNetV2 graph adds (reference impl/deeplab/model.py): depthwise conv
with stride / atrous rate, ReLU6 BatchNorm, whole-map mean + broadcast, Dropout, and the probability head (activation at 1/8
resolution followed by the align_corners bilinear resize).  References: fp32 torch ops on the same bf16 operands.
Tolerances as in test_gpu_ops.py: bf16 outputs 3e-3 relative L2, f32 outputs 1e-4."""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.philox import dropout_keep_mask
from segmentation_training_pipeline_b200 import lib
from tests.util import T, rand_bf16, ref, rel_err, stream

pytestmark = pytest.mark.gpu
TOL_F32, TOL_BF16 = 1e-4, 3e-3


def _same_pad(size, k, stride, dil):
    """TF 'same': (pad_before, out)"""
    out = -(-size // stride)
    ke = (k - 1) * dil + 1
    total = max((out - 1) * stride + ke - size, 0)
    return total // 2, out


DW_CASES = [
    # n, h, w, c, k, stride, dilation
    (2, 16, 16, 32, 3, 1, 1),
    (2, 16, 16, 96, 3, 2, 1),      # even size, stride 2: pad 0 before / 1 after
    (1, 15, 17, 144, 3, 2, 1),     # odd sizes, stride 2: pad 1 / 1
    (2, 12, 12, 192, 3, 1, 2),
    (1, 10, 14, 576, 3, 1, 4),
    (1, 9, 9, 960, 3, 1, 4),       # the 1/8 map of a small image, rate 4
    (1, 8, 8, 24, 5, 1, 1),
]


@pytest.mark.parametrize("case", DW_CASES)
def test_dwconv(stp, cuda, case):
    n, h, w, c, k, stride, dil = case
    g = torch.Generator().manual_seed(sum(case))
    (ph, ho), (pw, wo) = _same_pad(h, k, stride, dil), _same_pad(w, k, stride, dil)
    x = rand_bf16((n, h, w, c), g)
    wt = (torch.randn((k, k, c), generator=g) / k).to(torch.bfloat16).to(cuda)   # bf16 [k][k][C]
    wb = wt.float()
    dy = rand_bf16((n, ho, wo, c), g)
    res = rand_bf16((n, h, w, c), g)
    desc = lib.DwConvDesc(k, stride, dil, ph, pw)
    y = torch.zeros((n, ho, wo, c), dtype=torch.bfloat16, device=cuda)
    xs, ys, dys = T(x), T(y), T(dy)
    stp.dwconv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), ref(ys), stream())

    xr = x.float().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wb.cpu().permute(2, 0, 1).unsqueeze(1).contiguous().requires_grad_(True)  # (C,1,k,k)
    ke = (k - 1) * dil + 1
    pb, pr = max((ho - 1) * stride + ke - h - ph, 0), max((wo - 1) * stride + ke - w - pw, 0)
    yr = F.conv2d(F.pad(xr, (pw, pr, ph, pb)), wr, None, stride=stride, dilation=dil, groups=c)
    assert yr.shape[2:] == (ho, wo)
    assert rel_err(y, yr.detach().permute(0, 2, 3, 1)) < TOL_BF16
    yr.backward(dy.float().cpu().permute(0, 3, 1, 2))

    dx = torch.zeros((n, h, w, c), dtype=torch.bfloat16, device=cuda)
    dxs, rs = T(dx), T(res)
    stp.dwconv_dgrad(C.byref(desc), ref(dys), wt.data_ptr(), None, ref(dxs), stream())
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < TOL_BF16
    stp.dwconv_dgrad(C.byref(desc), ref(dys), wt.data_ptr(), ref(rs), ref(dxs), stream())
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1) + res.float().cpu()) < TOL_BF16

    nbytes = stp.dwconv_wgrad_workspace(C.byref(desc), ref(xs), ref(dys))
    ws = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=cuda)
    dw = torch.zeros((k, k, c), dtype=torch.float32, device=cuda)
    stp.dwconv_wgrad(C.byref(desc), ref(xs), ref(dys), dw.data_ptr(), ws.data_ptr(), ws.numel(), stream())
    assert rel_err(dw, wr.grad.squeeze(1).permute(1, 2, 0)) < TOL_F32
    dw2 = torch.zeros_like(dw)
    stp.dwconv_wgrad(C.byref(desc), ref(xs), ref(dys), dw2.data_ptr(), ws.data_ptr(), ws.numel(), stream())
    assert torch.equal(dw, dw2), "wgrad must be deterministic"


def test_dwconv_concat_slices(stp, cuda):
    """input and output as channel slices of wider buffers (ld > c)"""
    g = torch.Generator().manual_seed(5)
    n, h, w, c = 1, 8, 8, 16
    big = rand_bf16((n, h, w, 48), g)
    out = torch.zeros((n, h, w, 32), dtype=torch.bfloat16, device=cuda)
    wt = (torch.randn((3, 3, c), generator=g) / 3).to(torch.bfloat16).to(cuda)
    desc = lib.DwConvDesc(3, 1, 1, 1, 1)
    xs, ys = T(big, 16, c), T(out, 8, c)
    stp.dwconv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), ref(ys), stream())
    xr = big[..., 16:32].float().cpu().permute(0, 3, 1, 2)
    wr = wt.float().cpu().permute(2, 0, 1).unsqueeze(1)
    yr = F.conv2d(xr, wr, None, padding=1, groups=c).permute(0, 2, 3, 1)
    assert rel_err(out[..., 8:24], yr) < TOL_BF16
    assert float(out[..., :8].abs().max()) == 0 and float(out[..., 24:].abs().max()) == 0


def test_dwconv_rejects_bad_arguments(stp, cuda):
    x = torch.zeros((1, 8, 8, 12), dtype=torch.bfloat16, device=cuda)     # c % 8 != 0
    y = torch.zeros((1, 8, 8, 12), dtype=torch.bfloat16, device=cuda)
    wt = torch.zeros(9 * 16, dtype=torch.bfloat16, device=cuda)
    desc = lib.DwConvDesc(3, 1, 1, 1, 1)
    xs, ys = T(x), T(y)
    with pytest.raises(lib.StpError):
        stp.dwconv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), ref(ys), stream())
    x = torch.zeros((1, 8, 8, 16), dtype=torch.bfloat16, device=cuda)
    y = torch.zeros((1, 8, 8, 16), dtype=torch.bfloat16, device=cuda)
    xs, ys = T(x), T(y)
    bad = lib.DwConvDesc(7, 1, 1, 3, 3)
    with pytest.raises(lib.StpError):
        stp.dwconv_fwd(C.byref(bad), ref(xs), wt.data_ptr(), ref(ys), stream())


@pytest.mark.parametrize("shape", [(2, 16, 16, 32), (1, 9, 7, 96)])
def test_bn_relu6(stp, cuda, shape):
    """activation code 2: y = min(max(bn(x), 0), 6); gradient passes on 0 < t <= 6"""
    n, h, w, c = shape
    g = torch.Generator().manual_seed(c)
    x = rand_bf16(shape, g, scale=3.0)
    dy = rand_bf16(shape, g)
    gamma = (1.0 + 0.3 * torch.randn(c, generator=g)).to(cuda)
    beta = (3.5 + 1.5 * torch.randn(c, generator=g)).to(cuda)
    rows = n * h * w
    nblk = stp.bn_nblk(rows, c)
    partial = torch.zeros(2 * nblk * c, device=cuda)
    sync = torch.zeros(64, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * c, dtype=torch.float64, device=cuda)
    mm, mv = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
    coef = torch.zeros(4 * c, device=cuda)
    xs = T(x)
    stp.bn_stats_fused(ref(xs), partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.999,
                       mm.data_ptr(), mv.data_ptr(), coef.data_ptr(), stream())
    y = torch.zeros(shape, dtype=torch.bfloat16, device=cuda)
    ys, dys = T(y), T(dy)
    stp.bn_apply(ref(xs), coef.data_ptr(), 2, 1, ref(ys), stream())

    xr = x.float().cpu().requires_grad_(True)
    gr, br = gamma.cpu().requires_grad_(True), beta.cpu().requires_grad_(True)
    mean = xr.mean((0, 1, 2))
    var = ((xr - mean) ** 2).mean((0, 1, 2))
    t = (xr - mean) / torch.sqrt(var + 1e-3) * gr + br
    yr = torch.clamp(t, 0.0, 6.0)
    frac6 = float((yr == 6.0).float().mean())
    assert 0.02 < frac6 < 0.9, frac6          # the upper clamp is exercised
    assert rel_err(y, yr.detach()) < TOL_BF16
    assert float(y.float().max()) <= 6.0
    yr.backward(dy.float().cpu())

    bcoef = torch.zeros(3 * c, device=cuda)
    dgamma, dbeta = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
    stp.bn_bwd_reduce_fused(ref(dys), ref(xs), coef.data_ptr(), 2, 1, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                            dgamma.data_ptr(), dbeta.data_ptr(), bcoef.data_ptr(), stream())
    dx = torch.zeros(shape, dtype=torch.bfloat16, device=cuda)
    dxs = T(dx)
    stp.bn_bwd_apply(ref(dys), ref(xs), coef.data_ptr(), bcoef.data_ptr(), 2, 1, None, ref(dxs), stream())
    # elements whose pre-activation sits within bf16 noise of a clamp boundary may flip their mask: compare in L2
    assert rel_err(dgamma, gr.grad) < 2e-3
    assert rel_err(dbeta, br.grad) < 2e-3
    assert rel_err(dx, xr.grad) < 1e-2


@pytest.mark.parametrize("shape", [(2, 8, 8, 64), (3, 5, 7, 320), (1, 40, 40, 256)])
def test_global_avgpool_and_broadcast(stp, cuda, shape):
    n, h, w, c = shape
    g = torch.Generator().manual_seed(h)
    x = rand_bf16(shape, g)
    small = torch.zeros((n, 1, 1, c), dtype=torch.bfloat16, device=cuda)
    xs, ss = T(x), T(small)
    stp.global_avgpool_fwd(ref(xs), ref(ss), stream())
    assert rel_err(small, x.float().mean((1, 2), keepdim=True)) < TOL_BF16
    res = rand_bf16(shape, g)
    dx = torch.zeros(shape, dtype=torch.bfloat16, device=cuda)
    dxs, rs = T(dx), T(res)
    stp.global_avgpool_bwd(ref(ss), ref(rs), ref(dxs), stream())
    assert rel_err(dx, small.float() / (h * w) + res.float()) < TOL_BF16
    stp.broadcast_fwd(ref(ss), ref(dxs), stream())
    assert torch.equal(dx, small.expand(n, h, w, c))
    stp.broadcast_bwd(ref(xs), ref(ss), stream())
    assert rel_err(small, x.float().sum((1, 2), keepdim=True)) < TOL_BF16


@pytest.mark.parametrize("rate", [0.1, 0.5])
def test_dropout_matches_oracle_mask(stp, cuda, rate):
    n, h, w, c = 2, 9, 11, 48
    g = torch.Generator().manual_seed(3)
    x = rand_bf16((n, h, w, c), g)
    y = torch.zeros_like(x)
    step = torch.tensor([12345678901], dtype=torch.int64, device=cuda)
    seed, salt = 0x1234567890ABCDEF, 77
    xs, ys = T(x), T(y)
    stp.dropout(ref(xs), rate, seed, salt, step.data_ptr(), ref(ys), stream())
    keep = dropout_keep_mask(n * h * w, c, rate, seed, salt, int(step.item())).reshape(n, h, w, c)
    want = torch.where(torch.from_numpy(keep), x.float().cpu() * (1.0 / (1.0 - np.float32(rate))), torch.zeros(()))
    assert torch.equal(y.float().cpu(), want.to(torch.bfloat16).float())
    assert abs(keep.mean() - (1 - rate)) < 0.02
    # the same call on the gradient is the backward; another step -> another mask; in place is allowed
    y2 = x.clone()
    y2s = T(y2)
    stp.dropout(ref(y2s), rate, seed, salt, step.data_ptr(), ref(y2s), stream())
    assert torch.equal(y2, y)
    step += 1
    stp.dropout(ref(xs), rate, seed, salt, step.data_ptr(), ref(ys), stream())
    assert not torch.equal(y2, y)


@pytest.mark.parametrize("case", [(2, 5, 5, 40, 40, 1, 1), (1, 6, 9, 41, 70, 3, 1), (2, 5, 5, 40, 40, 4, 2), (1, 1, 1, 8, 8, 1, 1)])
def test_prob_head(stp, cuda, case):
    """activation(logits) == align_corners-resize(activation(z)), and the backward is the chain rule through all three steps"""
    n, h, w, H, W, classes, act = case
    g = torch.Generator().manual_seed(H + classes)
    cpad = 16
    z = torch.zeros((n, h, w, cpad), dtype=torch.float32)
    z[..., :classes] = torch.randn((n, h, w, classes), generator=g) * 2.0
    zc = z.to(cuda)
    logits = torch.zeros((n, H, W, classes), dtype=torch.float32, device=cuda)
    zs, ls = T(zc), T(logits)
    stp.prob_head_fwd(ref(zs), classes, act, ref(ls), stream())

    zr = z[..., :classes].permute(0, 3, 1, 2).clone().requires_grad_(True)
    p_small = torch.sigmoid(zr) if act == 1 else torch.softmax(zr, dim=1)
    p = F.interpolate(p_small, size=(H, W), mode="bilinear", align_corners=True)
    got = torch.sigmoid(logits) if act == 1 else torch.softmax(logits, dim=-1)
    assert float((got.cpu() - p.detach().permute(0, 2, 3, 1)).abs().max()) < 2e-6

    dl = torch.randn((n, H, W, classes), generator=g)
    # the loss kernels hand back dL/dlogit; build it from a probability-space gradient so the reference is plain autograd
    dp = torch.randn((n, classes, H, W), generator=g)
    pd = p.detach()
    if act == 1:
        dl = (dp * pd * (1 - pd)).permute(0, 2, 3, 1).contiguous()
    else:
        dl = (pd * (dp - (dp * pd).sum(1, keepdim=True))).permute(0, 2, 3, 1).contiguous()
        dp = dp - (dp * pd).sum(1, keepdim=True)      # what dl / p reconstructs: the component the softmax keeps
    p.backward(dp)
    dlc = dl.to(cuda)
    dz = torch.full((n, h, w, cpad), 7.0, dtype=torch.bfloat16, device=cuda)
    dls, dzs = T(dlc), T(dz)
    stp.prob_head_bwd(ref(dls), ref(ls), ref(zs), classes, act, ref(dzs), stream())
    assert rel_err(dz[..., :classes], zr.grad.permute(0, 2, 3, 1)) < 4e-3
    assert float(dz[..., classes:].abs().max()) == 0.0
