"""Helpers for the GPU parity tests: build stp_tensor structs over torch CUDA tensors, fp32 references."""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from segmentation_training_pipeline_b200 import lib

_DT = {torch.bfloat16: lib.BF16, torch.float32: lib.F32, torch.uint8: lib.U8}


def T(t: torch.Tensor, c_off=0, c=None):
    """stp_tensor over an NHWC torch tensor (contiguous storage), optionally a channel slice."""
    assert t.is_contiguous() and t.dim() == 4
    n, h, w, ld = t.shape
    c = ld - c_off if c is None else c
    st = lib.Tensor(t.data_ptr() + c_off * t.element_size(), n, h, w, c, ld, _DT[t.dtype])
    return st


def ref(st):
    return C.byref(st)


def stream():
    return torch.cuda.current_stream().cuda_stream


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def rand_bf16(shape, gen, scale=1.0, device="cuda"):
    x = (torch.randn(shape, generator=gen) * scale).to(torch.bfloat16)
    return x.to(device)


def conv_ref(x_nhwc, w_krsc, stride=1, pad=0, up=1, out_hw=None):
    """fp32 CPU reference of stp_conv_fwd semantics (zero insertion `up`, pad-before, output size given)."""
    x = x_nhwc.float().cpu().permute(0, 3, 1, 2)
    w = w_krsc.float().cpu().permute(0, 3, 1, 2)  # KRSC -> (Cout, Cin, R, S)
    if up > 1:
        n, c, h, ww = x.shape
        z = torch.zeros(n, c, (h - 1) * up + 1, (ww - 1) * up + 1)
        z[:, :, ::up, ::up] = x
        x = z
    R, S = w.shape[2], w.shape[3]
    if out_hw is not None:
        ho, wo = out_hw
        need_h = (ho - 1) * stride + R
        need_w = (wo - 1) * stride + S
        pb = max(need_h - x.shape[2] - pad, 0)
        pr = max(need_w - x.shape[3] - pad, 0)
        x = F.pad(x, (pad, pr, pad, pb))
        y = F.conv2d(x, w, None, stride=stride)[:, :, :ho, :wo]
    else:
        y = F.conv2d(F.pad(x, (pad,) * 4), w, None, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().cpu().double(), b.float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b) -> float:
    return float((a.float().cpu() - b.float().cpu()).abs().max())
