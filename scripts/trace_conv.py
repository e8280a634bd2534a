"""Phase timeline (globaltimer) of conv_tc2_kernel's first and last CTA for the encoder 3x3 shapes."""
import ctypes as C, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream
L = lib.Lib(); dev = torch.device("cuda:0")
trace = torch.zeros(64, dtype=torch.int64, device=dev)
names = ["entry", "prologue done", "pdl_wait done", "first stage full (MMA)", "last MMA committed", "first acc_full (epi)",
         "epilogue loop done", "stores drained", "epi warp at exit", "after final sync"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (n, h, w, cin, cout) in [(16, 128, 128, 64, 64), (16, 64, 64, 128, 128), (16, 32, 32, 256, 256), (16, 16, 16, 512, 512)]:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wt = (torch.randn(cout, 3, 3, cin, device=dev) / math.sqrt(9 * cin)).to(torch.bfloat16)
    y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev)
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0); xs, ys = T(x), T(y)
    for it in range(3):
        flush.zero_(); trace.zero_()
        L.set_trace_buffer(trace.data_ptr())
        L.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), None, 0, stream())
        torch.cuda.synchronize(); L.set_trace_buffer(None)
    t = trace.cpu().tolist()
    print("shape", (n, h, w, cin, cout))
    for base, who in ((0, "CTA 0"), (16, "last CTA")):
        t0 = t[base]
        print("  %s: " % who + ", ".join("%s +%.1fus" % (names[i], (t[base + i] - t0) / 1e3) for i in range(1, 10) if t[base + i]))
    print("  last CTA entry vs CTA0 entry: %.1f us" % ((t[16] - t[0]) / 1e3))
