"""Micro-benchmark single conv shapes through the C ABI (CUDA events, L2 flushed between launches)."""
import ctypes as C, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream

L = lib.Lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def bench(n, h, w, cin, cout, k=3, reps=10, what="fwd", res=False):
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wt = (torch.randn(cout, k, k, cin, device=dev) / math.sqrt(k * k * cin)).to(torch.bfloat16)
    y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(n, h, w, cout, device=dev).to(torch.bfloat16) if res else None
    desc = lib.ConvDesc(k, k, 1, k // 2, k // 2, 1, 0)
    xs, ys = T(x), T(y)
    rs = T(r) if res else None
    dw = torch.zeros(cout, k, k, cin, device=dev)
    nws = L.conv_wgrad_workspace(C.byref(desc), ref(xs), ref(ys))
    ws = torch.zeros(max(nws, 16), dtype=torch.uint8, device=dev)
    def run():
        if what == "fwd":
            L.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(rs) if res else None, ref(ys), None, 0, stream())
        else:
            L.conv_wgrad(C.byref(desc), ref(xs), ref(ys), dw.data_ptr(), ws.data_ptr(), ws.numel(), stream())
    for _ in range(3): run()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = ts[len(ts) // 2]
    fl = 2.0 * n * h * w * cout * k * k * cin
    return ms * 1e3, fl / ms / 1e9

shapes = [(16, 128, 128, 64, 64), (16, 64, 64, 128, 128), (16, 32, 32, 256, 256), (16, 16, 16, 512, 512), (16, 512, 512, 16, 16), (16, 256, 256, 32, 32)]
if __name__ != "__main__":
    pass
elif len(sys.argv) > 1 and sys.argv[1] == "dbg":
    for shp in shapes[:4]:
        for ver in (1, 0):
            L.set_option(b"tc_conv_version", ver)
            for dbg in ([0] if ver == 1 else [0, 1, 2, 3, 4, 8, 4 | 8, 1 | 2 | 4, 1 | 2 | 8, 15]):
                L.set_option(b"tc2_debug", dbg)
                us, tf = bench(*shp)
                print("shape", shp, "ver", "v1" if ver == 1 else "v2", "dbg", dbg, "us %.1f TF %.0f" % (us, tf), flush=True)
        L.set_option(b"tc2_debug", 0); L.set_option(b"tc_conv_version", 0)
elif len(sys.argv) > 1 and sys.argv[1] == "wdbg":
    for shp in shapes[:3]:
        for dbg in [0, 1, 2, 3, 4, 8, 4 | 8, 1 | 2 | 4, 1 | 2 | 8, 15]:
            L.set_option(b"tc2_debug", dbg)
            us, tf = bench(*shp, what="wgrad")
            print("wgrad shape", shp, "dbg", dbg, "us %.1f TF %.0f" % (us, tf), flush=True)
        L.set_option(b"tc2_debug", 0)
else:
    for shp in shapes:
        for what in ("fwd", "wgrad"):
            us, tf = bench(*shp, what=what)
            print("shape", shp, what, "us %.1f TF %.0f" % (us, tf), flush=True)
