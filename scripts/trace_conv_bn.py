"""Phase timeline of the forward conv WITH the fused BatchNorm-statistics epilogue vs the plain conv (tc2 and tc3)."""
import ctypes as C, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream
L = lib.Lib(); dev = torch.device("cuda:0")
trace = torch.zeros(64, dtype=torch.int64, device=dev)
names = {3: "first data", 10: "tile-1 MMAs", 4: "last MMA", 5: "first acc_full", 11: "tile-1 epi done", 6: "epi loop done", 7: "stores drained", 8: "epi exit", 9: "final sync"}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (n, h, w, cin, cout) in [(16, 64, 64, 128, 128), (16, 128, 128, 64, 64)]:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wt = (torch.randn(cout, 3, 3, cin, device=dev) / math.sqrt(9 * cin)).to(torch.bfloat16)
    y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev)
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0); xs, ys = T(x), T(y)
    rows = n * h * w
    partial = torch.zeros(2 * L.bn_nblk(rows, cout) * cout, device=dev)
    sync = torch.zeros(4, dtype=torch.int32, device=dev)
    acc = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    gamma, beta = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    coef = torch.zeros(4 * cout, device=dev); mm, mv = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
    bn = lib.BnFwd(partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99, mm.data_ptr(), mv.data_ptr(), coef.data_ptr())
    for kern, opt in (("tc2", 1), ("tc3", 2)):
        L.set_option(b"tc3", opt)
        for with_bn in (0, 1):
            ts = []
            for it in range(4):
                flush.zero_(); trace.zero_()
                L.set_trace_buffer(trace.data_ptr())
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if with_bn:
                    L.conv_fwd_bn(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), C.byref(bn), None, 0, stream())
                else:
                    L.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), None, 0, stream())
                b.record(); torch.cuda.synchronize(); L.set_trace_buffer(None)
                ts.append(a.elapsed_time(b) * 1e3)
            t = trace.cpu().tolist()
            for base, who in ((0, "CTA0"), (16, "last")):
                print((n, h, w, cin, cout), kern, "bn" if with_bn else "plain", "event us %.1f" % sorted(ts)[1], who,
                      ", ".join("%s +%.1f" % (names[i], (t[base + i] - t[base]) / 1e3) for i in (3, 10, 4, 11, 6, 7, 8, 9) if t[base + i]), flush=True)
L.set_option(b"tc3", 0)
