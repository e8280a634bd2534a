"""Launch the depthwise conv kernels at the MobileNetV2 shapes of the DeepLabV3 step (for ncu / event timing)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from segmentation_training_pipeline_b200 import lib  # noqa: E402
from tests.util import T, ref, stream  # noqa: E402

L = lib.Lib()
dev = "cuda:0"
shapes = [(16, 40, 40, 960, 1, 4), (16, 40, 40, 576, 1, 2), (16, 80, 80, 144, 2, 1), (16, 160, 160, 32, 1, 1), (16, 160, 160, 96, 2, 1)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for n, h, w, c, stride, dil in shapes:
    ho, wo = -(-h // stride), -(-w // stride)
    tot = max((ho - 1) * stride + 2 * dil + 1 - h, 0)
    x = torch.randn((n, h, w, c), device=dev).to(torch.bfloat16)
    dy = torch.randn((n, ho, wo, c), device=dev).to(torch.bfloat16)
    y = torch.zeros_like(dy)
    dx = torch.zeros_like(x)
    wt = torch.randn((3, 3, c), device=dev).to(torch.bfloat16)
    dw = torch.zeros((3, 3, c), device=dev)
    desc = lib.DwConvDesc(3, stride, dil, tot // 2, tot // 2)
    xs, ys, dys, dxs = T(x), T(y), T(dy), T(dx)
    ws = torch.zeros(int(L.dwconv_wgrad_workspace(C.byref(desc), ref(xs), ref(dys))) + 16, dtype=torch.uint8, device=dev)
    fns = {
        "fwd": lambda: L.dwconv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), ref(ys), stream()),
        "dgrad": lambda: L.dwconv_dgrad(C.byref(desc), ref(dys), wt.data_ptr(), None, ref(dxs), stream()),
        "wgrad": lambda: L.dwconv_wgrad(C.byref(desc), ref(xs), ref(dys), dw.data_ptr(), ws.data_ptr(), ws.numel(), stream()),
    }
    for name, fn in fns.items():
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / reps * 1e3
        mb = (x.numel() + dy.numel()) * 2 / 1e6
        print("%-6s n%d %dx%d c%d s%d d%d: %8.1f us  (%.0f MB min traffic -> %.2f TB/s)" % (name, n, h, w, c, stride, dil, us, mb, mb / us))
