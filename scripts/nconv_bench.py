"""Narrow-channel 3x3 convs of the decoder tail (Cin, Cout in {16, 32}) through the C ABI: csrc/conv_narrow.cu (option nconv = 1)
against the tcgen05 halo kernel conv_tc2 (nconv = 0, the default).  Per shape: plain forward, forward + BatchNorm-statistics epilogue, dgrad +
fused BatchNorm-backward reduction; isolated (one launch after a 256 MB L2 flush) and back to back (20 launches over 4 rotating
input / output sets, > 126 MB L2), algorithmic bytes (x read once + y written once [+ BatchNorm input read once]) / time against
the measured copy bandwidth."""
import ctypes as C, json, math, os, sys
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
    PEAK = 6552.6
if isinstance(PEAK, dict):
    PEAK = PEAK.get("burst", 6552.6)
NSET = 4


def bench(n, h, w, cin, cout, what):
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
    sets = []
    for _ in range(NSET):
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev)
        bx = torch.randn(n, h, w, cout, device=dev).to(torch.bfloat16)
        sets.append((x, y, bx, T(x), T(y), T(bx)))
    wt = (torch.randn(cout, 3, 3, cin, device=dev) / math.sqrt(9 * cin)).to(torch.bfloat16)
    rows = n * h * w
    partial = torch.zeros(2 * L.bn_nblk(rows, cout) * cout, device=dev)
    sync = torch.zeros(4, dtype=torch.int32, device=dev)
    acc = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    gamma, beta = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    coef = torch.zeros(4 * cout, device=dev)
    coef[cout:3 * cout] = 1.0
    mm, mv = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
    dg, db, bco = torch.zeros(cout, device=dev), torch.zeros(cout, device=dev), torch.zeros(3 * cout, device=dev)
    bn = lib.BnFwd(partial.data_ptr(), sync.data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                   mm.data_ptr(), mv.data_ptr(), coef.data_ptr())

    def run(i):
        x, y, bx, xs, ys, bxs = sets[i % NSET]
        if what == "fwd":
            L.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), None, 0, stream())
        elif what == "fwd_bn":
            L.conv_fwd_bn(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), C.byref(bn), None, 0, stream())
        else:   # dgrad of a (cout -> cin) forward conv reads "dy" = x (cin ch here) and writes dx = y; BatchNorm input = bx
            bnb = lib.BnBwd(C.pointer(bxs), coef.data_ptr(), 1, partial.data_ptr(), sync.data_ptr(), acc.data_ptr(),
                            dg.data_ptr(), db.data_ptr(), bco.data_ptr())
            dd = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
            L.conv_dgrad_bn(C.byref(dd), ref(xs), wt.data_ptr(), ref(ys), C.byref(bnb), None, 0, stream())
    for i in range(3):
        run(i)
    ts = []
    for i in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(i); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    iso = sorted(ts)[len(ts) // 2] * 1e3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(20):
        run(i)
    b.record(); torch.cuda.synchronize()
    b2b = a.elapsed_time(b) * 1e3 / 20
    byt = rows * 2.0 * (cin + cout + (cout if what == "dgrad_bn" else 0))
    return iso, b2b, byt


if __name__ == "__main__":
    shapes = [(16, 512, 512, 16, 16), (16, 512, 512, 32, 16), (16, 512, 512, 16, 32), (16, 256, 256, 32, 32)]
    print("copy peak %.1f GB/s" % PEAK)
    for shp in shapes:
        for what in ("fwd", "fwd_bn", "dgrad_bn"):
            line = "%-26s %-8s" % (shp, what)
            for off in (0, 1):
                L.set_option(b"nconv", 1 - off)
                iso, b2b, byt = bench(*shp, what)
                line += "  %s: isolated %6.1f us  b2b %6.1f us (%.2f of copy peak)" % ("conv_narrow" if off == 0 else "conv_tc2   ", iso, b2b, byt / b2b / 1e3 / PEAK)
            L.set_option(b"nconv", 0)
            print(line, flush=True)
