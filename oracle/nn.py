"""Keras-2.2.4 / TF-1.15 layer semantics restated on PyTorch-CPU fp32  (TEST INFRASTRUCTURE, parity unpinned).

Every function cites the reference call site that selects the behaviour and the [DEP] package whose
published algorithm is restated.  Activations are NCHW torch tensors inside the
oracle (torch's native conv layout); parameters are kept in KERAS layouts so that weight dictionaries
can be exchanged with the engine by layer name:
    conv kernel  (kh, kw, Cin, Cout)      bias (Cout)
    BN           gamma, beta, moving_mean, moving_var  (C)

`storage="bf16"` emulates the engine's storage precision: every tensor the engine materialises in HBM
as bf16 is rounded to bf16 at the same point (forward value AND its gradient), accumulation stays fp32.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


class _RoundBF16(torch.autograd.Function):
    """Round-to-nearest-even to bf16 and back, in forward and in backward (engine stores both as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _RoundFwdOnly(torch.autograd.Function):
    """bf16 rounding of a master-weight copy: gradient passes through in fp32 (dW is kept fp32)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def rb(x: torch.Tensor, storage: str) -> torch.Tensor:
    return _RoundBF16.apply(x) if storage == "bf16" else x


def rw(w: torch.Tensor, storage: str) -> torch.Tensor:
    return _RoundFwdOnly.apply(w) if storage == "bf16" else w


# ----------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------
def keras_same_pad(size: int, k: int, s: int):
    """TF 'SAME' padding rule [DEP tensorflow==1.15]: total=max((ceil(H/s)-1)*s+k-H,0), before=total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d(x, kernel_hwio, bias=None, stride=1, padding="valid", storage="fp32"):
    """keras.layers.Conv2D [DEP keras>=2.2.4]; kernel in Keras (kh,kw,Cin,Cout) layout.

    padding: "valid" | "same" | int (explicit symmetric ZeroPadding2D in front, qubvel ResNet style).
    """
    w = rw(kernel_hwio, storage).permute(3, 2, 0, 1)  # -> (Cout, Cin, kh, kw)
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    if padding == "same":
        pt, pb = keras_same_pad(x.shape[2], kh, stride)
        pl, pr = keras_same_pad(x.shape[3], kw, stride)
        x = F.pad(x, (pl, pr, pt, pb))
    elif isinstance(padding, int) and padding > 0:
        x = F.pad(x, (padding,) * 4)
    y = F.conv2d(x, w, None, stride=stride)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def depthwise_conv2d(x, kernel_hwc1, stride=1, dilation=1, storage="fp32", explicit_pad=None):
    """keras.layers.DepthwiseConv2D(padding='same', dilation_rate, use_bias=False) [DEP]; kernel (kh, kw, C, 1).
    TF 'SAME' with the dilated extent (k-1)*rate+1 (reference impl/deeplab/model.py:252-255)."""
    kh, kw, c, _ = kernel_hwc1.shape
    w = rw(kernel_hwc1, storage).permute(2, 3, 0, 1)  # -> (C, 1, kh, kw)
    if explicit_pad is not None:   # ZeroPadding2D((beg, end)) + padding='valid' (SepConv_BN with stride > 1, model.py:125-131)
        pt, pb = explicit_pad
        pl, pr = explicit_pad
    else:
        pt, pb = keras_same_pad(x.shape[2], (kh - 1) * dilation + 1, stride)
        pl, pr = keras_same_pad(x.shape[3], (kw - 1) * dilation + 1, stride)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, None, stride=stride, dilation=dilation, groups=c)


def conv2d_transpose(x, kernel_hwoi, bias=None, stride=2, storage="fp32"):
    """keras.layers.Conv2DTranspose(padding='same') [DEP]; Keras kernel layout (kh,kw,Cout,Cin).

    k=4,s=2: == torch conv_transpose2d(padding=1) (exactly 2H).  k=3,s=2: conv_transpose2d(padding=0)
    cropped to [0:2H] (SURVEY.md Appendix B derivation).
    """
    kh = kernel_hwoi.shape[0]
    w = rw(kernel_hwoi, storage).permute(3, 2, 0, 1)  # torch wants (Cin, Cout, kh, kw)
    H, W = x.shape[2], x.shape[3]
    if kh == 4 and stride == 2:
        y = F.conv_transpose2d(x, w, None, stride=2, padding=1)
    else:
        y = F.conv_transpose2d(x, w, None, stride=stride, padding=0)[:, :, : H * stride, : W * stride]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def batchnorm_train(x, gamma, beta, eps):
    """keras BatchNormalization(axis=3) in training mode [DEP]: biased batch variance over N,H,W."""
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps)
    if gamma is not None:
        y = y * gamma.view(1, -1, 1, 1)
    if beta is not None:
        y = y + beta.view(1, -1, 1, 1)
    return y, mean.flatten(), var.flatten()


def batchnorm_infer(x, gamma, beta, mov_mean, mov_var, eps):
    y = (x - mov_mean.view(1, -1, 1, 1)) * torch.rsqrt(mov_var.view(1, -1, 1, 1) + eps)
    if gamma is not None:
        y = y * gamma.view(1, -1, 1, 1)
    if beta is not None:
        y = y + beta.view(1, -1, 1, 1)
    return y


def maxpool(x, k, s, pad):
    """ZeroPadding2D(pad)+MaxPooling2D(k, s, 'valid') on post-ReLU data == max_pool2d(k,s,pad)."""
    return F.max_pool2d(x, k, s, pad)


def upsample_nearest(x, r=2):
    """keras UpSampling2D(r) [DEP]: element repeat."""
    return x.repeat_interleave(r, dim=2).repeat_interleave(r, dim=3)


def resize_bilinear_tf1(x, out_h, out_w, align_corners=False):
    """TF1 legacy tf.image.resize_bilinear [DEP tensorflow==1.15]: src = dst*in/out (no half pixel).

    align_corners=True is the impl/deeplab/model.py:92-100 variant (scale=(in-1)/(out-1))."""
    N, C, H, W = x.shape

    def axis(inp, out):
        if align_corners and out > 1:
            scale = (inp - 1) / (out - 1)
        else:
            scale = inp / out
        src = torch.arange(out, dtype=torch.float32) * np.float32(scale)
        lo = src.floor().long()
        hi = torch.clamp(lo + 1, max=inp - 1)
        w = src - lo.float()
        return lo, hi, w

    ylo, yhi, wy = axis(H, out_h)
    xlo, xhi, wx = axis(W, out_w)
    # TF's compute_lerp (resize_bilinear_op.cc): top = tl + (tr - tl)*x_lerp; out = top + (bottom - top)*y_lerp
    tl, tr = x[:, :, ylo][:, :, :, xlo], x[:, :, ylo][:, :, :, xhi]
    bl, br = x[:, :, yhi][:, :, :, xlo], x[:, :, yhi][:, :, :, xhi]
    top = tl + (tr - tl) * wx
    bot = bl + (br - bl) * wx
    return top + (bot - top) * wy.view(1, 1, -1, 1)


# ----------------------------------------------------------------------------------------------
# initialisers (keras.initializers [DEP]) on Keras-layout kernels
# ----------------------------------------------------------------------------------------------
def he_uniform(shape_hwio, gen: np.random.Generator):
    fan_in = shape_hwio[0] * shape_hwio[1] * shape_hwio[2]
    lim = math.sqrt(6.0 / fan_in)
    return gen.uniform(-lim, lim, size=shape_hwio).astype(np.float32)


def glorot_uniform(shape_hwio, gen: np.random.Generator):
    rf = shape_hwio[0] * shape_hwio[1]
    fan_in, fan_out = rf * shape_hwio[2], rf * shape_hwio[3]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return gen.uniform(-lim, lim, size=shape_hwio).astype(np.float32)
