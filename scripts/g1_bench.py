"""1x1 stride-1 convolutions (plain GEMMs) through the C ABI on csrc/gemm1x1.cu: ResNet-50 bottleneck / FPN / MobileNetV2 shapes at
bs16, plain and accumulate-in-place (residual == output, the dgrad form), back to back over rotating > L2 buffer sets; algorithmic
bytes (A read once + Y written once [+ residual read once]) / time against the measured copy bandwidth."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream

L = lib.Lib()
dev = torch.device("cuda:0")
PEAK = 6552.6
NSET = 4


def bench(n, h, w, cin, cout, res):
    desc = lib.ConvDesc(1, 1, 1, 0, 0, 1, 0)
    sets = []
    for _ in range(NSET):
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev)
        sets.append((x, y, T(x), T(y)))
    wt = (torch.randn(cout, 1, 1, cin, device=dev) / math.sqrt(cin)).to(torch.bfloat16)

    def run(i):
        x, y, xs, ys = sets[i % NSET]
        L.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, ref(ys) if res else None, ref(ys), None, 0, stream())
    for i in range(3):
        run(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(20):
        run(i)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20
    rows = n * h * w
    byt = rows * 2.0 * (cin + cout + (cout if res else 0))
    return us, byt / us / 1e3 / PEAK, 2.0 * rows * cin * cout / us / 1e6


if __name__ == "__main__":
    shapes = [(16, 128, 128, 64, 256), (16, 128, 128, 256, 64), (16, 64, 64, 128, 512), (16, 64, 64, 512, 128), (16, 32, 32, 256, 1024),
              (16, 32, 32, 1024, 256), (16, 16, 16, 512, 2048), (16, 16, 16, 2048, 512), (16, 128, 128, 256, 256),
              (16, 80, 80, 144, 24), (16, 80, 80, 24, 144), (16, 40, 40, 192, 32), (16, 20, 20, 960, 160), (16, 20, 20, 160, 960)]
    for o in sys.argv[1:]:
        k, v = o.split("=")
        L.set_option(k.encode(), int(v))
    print("options:", sys.argv[1:])
    for shp in shapes:
        line = "%-28s" % (shp,)
        for res in (0, 1):
            us, frac, tf = bench(*shp, res)
            line += "   %s %7.1f us  %.2f of copy peak  %5.0f TF/s" % ("accumulate" if res else "plain     ", us, frac, tf)
        print(line, flush=True)
