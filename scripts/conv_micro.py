"""Time the encoder/decoder 3x3 conv shapes under tiling options.  Usage: python scripts/conv_micro.py name=value ..."""
import sys, ctypes as C, math, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream
L = lib.Lib()
for kv in sys.argv[1:]:
    k, v = kv.split("="); L.set_option(k.encode(), int(v))
dev = torch.device("cuda:0"); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (n, h, w, cin, cout) in [(16, 128, 128, 64, 64), (16, 64, 64, 128, 128), (16, 32, 32, 256, 256), (16, 16, 16, 512, 512), (16, 32, 32, 768, 256), (16, 64, 64, 384, 128)]:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16); wt = (torch.randn(cout, 3, 3, cin, device=dev) / math.sqrt(9 * cin)).to(torch.bfloat16)
    y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev); desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0); xs, ys = T(x), T(y)
    run = lambda: L.conv_fwd(C.byref(desc), ref(xs), wt.data_ptr(), None, None, ref(ys), None, 0, stream())
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); print(" ".join(sys.argv[1:]), (n, h, w, cin, cout), "us %.1f" % (ts[5] * 1e3), flush=True)
