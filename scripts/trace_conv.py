"""Phase timeline (globaltimer) of conv_tc2_kernel's first and last CTA for the encoder 3x3 shapes."""
import ctypes as C, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream
L = lib.Lib(); dev = torch.device("cuda:0")
trace = torch.zeros(64, dtype=torch.int64, device=dev)
names = ["entry", "prologue done", "pdl_wait done", "first stage full (MMA)", "last MMA committed", "first acc_full (epi)",
         "epilogue loop done", "stores drained", "epi warp at exit", "after final sync", "tile-1 MMAs committed",
         "tile-1 epilogue done"]
variants = [("tc2", {b"tc3": 1}), ("tc2 skipAB", {b"tc3": 1, b"tc2_debug": 3}), ("tc3 skipAB", {b"tc3": 2, b"tc2_debug": 3}),
            ("tc3 bn128 mt1 skipAB", {b"tc3": 2, b"tc3_force_bn": 128, b"tc3_force_mt": 1, b"tc2_debug": 3})]
if len(sys.argv) > 1:
    variants = [("tc3 bn%s mt%s" % (sys.argv[1], sys.argv[2]), {b"tc3": 2, b"tc3_force_bn": int(sys.argv[1]), b"tc3_force_mt": int(sys.argv[2])})]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (n, h, w, cin, cout) in [(16, 64, 64, 128, 128), (16, 32, 32, 256, 256), (16, 16, 16, 512, 512)]:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wt = (torch.randn(cout, 3, 3, cin, device=dev) / math.sqrt(9 * cin)).to(torch.bfloat16)
    y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev)
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0); xs, ys = T(x), T(y)
    for vname, opts in variants:
        for k in (b"tc3", b"tc3_force_bn", b"tc3_force_mt", b"tc2_debug"):
            L.set_option(k, 0)
        for k, v in opts.items():
            L.set_option(k, v)
        for it in range(3):
            flush.zero_(); trace.zero_()
            L.set_trace_buffer(trace.data_ptr())
            L.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), None, 0, stream())
            torch.cuda.synchronize(); L.set_trace_buffer(None)
        t = trace.cpu().tolist()
        print("shape", (n, h, w, cin, cout), vname)
        for base, who in ((0, "CTA 0"),):
            t0 = t[base]
            print("  %s: " % who + ", ".join("%s +%.1fus" % (names[i], (t[base + i] - t0) / 1e3) for i in (3, 10, 11, 4, 6, 9) if t[base + i]))
        print("  last CTA entry vs CTA0 entry: %.1f us" % ((t[16] - t[0]) / 1e3))
