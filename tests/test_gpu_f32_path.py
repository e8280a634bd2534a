"""Unit parity of the PARITY-MODE kernels (csrc/f32_path.cu: fp32 tensors through the same C ABI entry points) against
fp32 / fp64 torch-CPU math.  Tolerances are fp32-level (1e-5 rel-L2; reductions 1e-6)."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from segmentation_training_pipeline_b200 import lib
from tests.test_gpu_ops import conv_ref_autograd
from tests.util import T, ref, rel_err, stream

pytestmark = pytest.mark.gpu

CASES = [
    # n, h, w, cin, cout, k, stride, pad, up
    (2, 16, 16, 64, 64, 3, 1, 1, 1),
    (2, 16, 16, 64, 128, 3, 2, 1, 1),
    (2, 16, 16, 64, 128, 1, 2, 0, 1),
    (2, 32, 32, 8, 64, 7, 2, 3, 1),
    (1, 9, 7, 16, 24, 3, 1, 1, 1),
    (1, 8, 8, 32, 16, 4, 1, 2, 2),
    (3, 20, 12, 24, 1, 3, 1, 1, 1),
    (1, 40, 24, 192, 64, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_f32_conv_fwd_dgrad_wgrad(stp, cuda, case):
    n, h, w, cin, cout, k, stride, pad, up = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn((n, h, w, cin), generator=g).to(cuda)
    wt = (torch.randn((cout, k, k, cin), generator=g) / math.sqrt(k * k * cin)).to(cuda)
    if up > 1:
        ho, wo = h * up, w * up
    else:
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    res = torch.randn((n, ho, wo, cout), generator=g).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    desc = lib.ConvDesc(k, k, stride, pad, pad, up, 0)
    y = torch.zeros((n, ho, wo, cout), device=cuda)
    xs, ys, rs = T(x), T(y), T(res)
    before = stp.tc_launch_count()
    stp.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), bias.data_ptr(), ref(rs), ref(ys), None, 0, stream())
    xr = x.double().cpu().requires_grad_(True)
    wr = wt.double().cpu().requires_grad_(True)
    out = conv_ref_autograd(xr, wr, stride, pad, up, (ho, wo))
    assert rel_err(y, (out + res.double().cpu() + bias.double().cpu()).float()) < 1e-5
    dy = torch.randn((n, ho, wo, cout), generator=g).to(cuda)
    out.backward(dy.double().cpu())
    dx = torch.zeros_like(x)
    dys, dxs = T(dy), T(dx)
    stp.conv_dgrad(C.byref(desc), ref(dys), wt.data_ptr(), None, ref(dxs), None, 0, stream())   # forward weights: flipped in-kernel
    assert rel_err(dx, xr.grad.float()) < 1e-5
    dx2 = x.clone()
    stp.conv_dgrad(C.byref(desc), ref(dys), wt.data_ptr(), ref(T(dx2)), ref(T(dx2)), None, 0, stream())
    assert rel_err(dx2, (xr.grad + x.double().cpu()).float()) < 1e-5
    dw = torch.zeros_like(wt)
    stp.conv_wgrad(C.byref(desc), ref(xs), ref(dys), dw.data_ptr(), None, 0, stream())
    assert rel_err(dw, wr.grad.float()) < 1e-5
    torch.cuda.synchronize()
    assert stp.tc_launch_count() == before


@pytest.mark.parametrize("up", [1, 2])
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 9, 7, 24), (2, 8, 8, 512), (4, 32, 32, 16)])
def test_f32_batchnorm_fwd_bwd(stp, cuda, shape, up):
    from oracle import nn as ON
    n, h, w, c = shape
    g = torch.Generator().manual_seed(c + up)
    x = (torch.randn(shape, generator=g) * 2.0 + torch.randn(c, generator=g) * 3.0).to(cuda)   # |mean| > std: cancellation-prone
    gamma = (torch.rand(c, generator=g) + 0.5).to(cuda)
    gamma[::3] *= -1.0
    beta = (torch.randn(c, generator=g) * 0.2).to(cuda)
    eps, mom = 2e-5, 0.99
    rows = n * h * w
    partial = torch.zeros(16, device=cuda)
    sync = torch.zeros(4, dtype=torch.int32, device=cuda)
    acc = torch.zeros(2 * c, dtype=torch.float64, device=cuda)
    coef = torch.zeros(4 * c, device=cuda)
    mm, mv = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
    y = torch.zeros((n, h * up, w * up, c), device=cuda)
    xs, ys = T(x), T(y)
    stp.bn_stats_fused(ref(xs), partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, mom,
                       mm.data_ptr(), mv.data_ptr(), coef.data_ptr(), stream())
    stp.bn_apply(ref(xs), coef.data_ptr(), 1, up, ref(ys), stream())
    xc = x.double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    gc, bc = gamma.double().cpu().requires_grad_(True), beta.double().cpu().requires_grad_(True)
    yo, mean, var = ON.batchnorm_train(xc, gc, bc, eps)
    yo = torch.relu(yo)
    if up == 2:
        yo = ON.upsample_nearest(yo, 2)
    assert float((coef[:c].double().cpu() - mean).abs().max()) < 1e-6 * (1 + float(mean.abs().max()))
    assert rel_err(coef[c:2 * c], torch.rsqrt(var + eps).float()) < 1e-6
    assert rel_err(y, yo.permute(0, 2, 3, 1).float()) < 1e-5
    assert float(acc.abs().max()) == 0.0
    dy = torch.randn((n, h * up, w * up, c), generator=g).to(cuda)
    yo.backward(dy.double().cpu().permute(0, 3, 1, 2))
    bcoef = torch.zeros(3 * c, device=cuda)
    dgamma, dbeta = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
    dx = torch.zeros_like(x)
    dys, dxs = T(dy), T(dx)
    stp.bn_bwd_reduce_fused(ref(dys), ref(xs), coef.data_ptr(), 1, up, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                            dgamma.data_ptr(), dbeta.data_ptr(), bcoef.data_ptr(), stream())
    stp.bn_bwd_apply(ref(dys), ref(xs), coef.data_ptr(), bcoef.data_ptr(), 1, up, None, ref(dxs), stream())
    torch.cuda.synchronize()
    print("dgamma", rel_err(dgamma, gc.grad.float()), "dbeta", rel_err(dbeta, bc.grad.float()), "dx", rel_err(dx, xc.grad.permute(0, 2, 3, 1).float()))
    assert rel_err(dgamma, gc.grad.float()) < 1e-5
    assert rel_err(dbeta, bc.grad.float()) < 1e-5
    assert rel_err(dx, xc.grad.permute(0, 2, 3, 1).float()) < 2e-5


def test_f32_maxpool_and_head(stp, cuda):
    g = torch.Generator().manual_seed(3)
    x = torch.relu(torch.randn((2, 16, 16, 64), generator=g)).to(cuda)
    y = torch.zeros((2, 8, 8, 64), device=cuda)
    am = torch.zeros(2 * 8 * 8 * 64, dtype=torch.uint8, device=cuda)
    stp.maxpool_fwd(ref(T(x)), 3, 2, 1, ref(T(y)), am.data_ptr(), stream())
    xc = x.double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    yo = F.max_pool2d(xc, 3, 2, 1)
    assert torch.equal(y.cpu(), yo.permute(0, 2, 3, 1).float())
    dy = torch.randn((2, 8, 8, 64), generator=g).to(cuda)
    yo.backward(dy.double().cpu().permute(0, 3, 1, 2))
    dx = torch.zeros_like(x)
    stp.maxpool_bwd(ref(T(dy)), am.data_ptr(), 3, 2, 1, None, ref(T(dx)), stream())
    # ties between equal maxima (ReLU zeros) route the gradient to the first one in scan order in both implementations only
    # where the value is positive; compare where x > 0
    m = (x > 0).cpu()
    assert rel_err(dx.cpu() * m, xc.grad.permute(0, 2, 3, 1).float() * m) < 1e-6
    # head: 3x3 'same' conv + bias to `classes` logits, backward dX / dW / db
    n, h, w, cin, cls = 2, 24, 16, 16, 1
    xh = torch.randn((n, h, w, cin), generator=g).to(cuda)
    wt = (torch.randn((cls, 3, 3, cin), generator=g) * 0.1).to(cuda)
    b = torch.randn(cls, generator=g).to(cuda)
    logits = torch.zeros(n * h * w * cls, device=cuda)
    ws = torch.zeros(4096, dtype=torch.uint8, device=cuda)
    stp.head_fwd(ref(T(xh)), wt.data_ptr(), b.data_ptr(), cls, logits.data_ptr(), ws.data_ptr(), ws.numel(), stream())
    xr, wr, br = xh.double().cpu().requires_grad_(True), wt.double().cpu().requires_grad_(True), b.double().cpu().requires_grad_(True)
    out = conv_ref_autograd(xr, wr, 1, 1, 1, (h, w)) + br
    assert rel_err(logits.view(n, h, w, cls), out.float()) < 1e-5
    dl = torch.randn((n, h, w, cls), generator=g).to(cuda)
    out.backward(dl.double().cpu())
    dxh, dw, db = torch.zeros_like(xh), torch.zeros_like(wt), torch.zeros_like(b)
    stp.head_bwd(ref(T(xh)), wt.data_ptr(), dl.data_ptr(), cls, ref(T(dxh)), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(),
                 stream())
    assert rel_err(dxh, xr.grad.float()) < 1e-5 and rel_err(dw, wr.grad.float()) < 1e-5 and rel_err(db, br.grad.float()) < 1e-5
