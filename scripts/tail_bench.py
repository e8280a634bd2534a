"""Isolated timings (CUDA events, L2 flushed) of the full-resolution tail kernels of the U-Net/ResNet-34 bs16 512^2 step: segmentation
head forward / backward, stem max-pool forward / backward -- each against its HBM roofline (algorithmic bytes / measured copy peak)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream

L = lib.Lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, us, nbytes):
    print("%-34s %7.1f us   %6.1f MB algorithmic -> %5.2f TB/s = %.2f of the measured copy peak" %
          (name, us, nbytes / 1e6, nbytes / us / 1e6, nbytes / us / 1e3 / PEAK), flush=True)


n, h, cin, cls = 16, 512, 16, 1
x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
w = (torch.randn(cls, 3, 3, cin, device=dev) * 0.1)
b = torch.zeros(cls, device=dev)
logits = torch.zeros(n * h * h * cls, device=dev)
dl = torch.randn(n * h * h * cls, device=dev)
dx = torch.zeros_like(x)
dw, db = torch.zeros_like(w), torch.zeros_like(b)
xs, dxs = T(x), T(dx)
ws = torch.zeros(max(int(L.head_fwd_workspace(ref(xs), cls)), int(L.head_bwd_workspace(ref(xs), cls)), 16), dtype=torch.uint8, device=dev)
M = n * h * h
report("head_fwd (16->1 @512^2)", timeit(lambda: L.head_fwd(ref(xs), w.data_ptr(), b.data_ptr(), cls, logits.data_ptr(), ws.data_ptr(), ws.numel(), stream())),
       M * cin * 2 + M * 4)
report("head_bwd (dgrad + wgrad)", timeit(lambda: L.head_bwd(ref(xs), w.data_ptr(), dl.data_ptr(), cls, ref(dxs), dw.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                                            ws.numel(), stream())), 2 * M * cin * 2 + 2 * M * 4)
# stem max-pool 3x3/2 pad 1: 16x256x256x64 -> 16x128x128x64
xp = torch.relu(torch.randn(16, 256, 256, 64, device=dev)).to(torch.bfloat16)
yp = torch.zeros(16, 128, 128, 64, dtype=torch.bfloat16, device=dev)
am = torch.zeros(yp.numel(), dtype=torch.uint8, device=dev)
dyp = torch.randn(16, 128, 128, 64, device=dev).to(torch.bfloat16)
dxp = torch.zeros_like(xp)
resp = torch.randn(16, 256, 256, 64, device=dev).to(torch.bfloat16)
report("maxpool_fwd 3x3/2 (64ch @256^2)", timeit(lambda: L.maxpool_fwd(ref(T(xp)), 3, 2, 1, ref(T(yp)), am.data_ptr(), stream())),
       xp.numel() * 2 + yp.numel() * 3)
report("maxpool_bwd 3x3/2 (+residual)", timeit(lambda: L.maxpool_bwd(ref(T(dyp)), am.data_ptr(), 3, 2, 1, ref(T(resp)), ref(T(dxp)), stream())),
       yp.numel() * 3 + 2 * xp.numel() * 2)
