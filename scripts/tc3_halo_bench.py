"""A/B of conv_tc3's single-halo-box mode (option tc3_halo: 1 on, 0 off): isolated launches (L2 flushed) and 30 launches back
to back over rotating input/output sets larger than L2 (how the layer runs inside the step graph)."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_training_pipeline_b200 import lib
from tests.util import T, ref, stream

L = lib.Lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(n, h, w, cin, cout, sets=10, reps=30):
    xs = [torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16) for _ in range(sets)]
    ys = [torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev) for _ in range(sets)]
    wt = (torch.randn(cout, 3, 3, cin, device=dev) / math.sqrt(9 * cin)).to(torch.bfloat16)
    desc = lib.ConvDesc(3, 3, 1, 1, 1, 1, 0)
    xt, yt = [T(x) for x in xs], [T(y) for y in ys]

    def run(i):
        L.conv_fwd(C.byref(desc), ref(xt[i % sets]), wt.data_ptr(), None, None, ref(yt[i % sets]), None, 0, stream())
    for i in range(3):
        run(i)
    iso = []
    for i in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(i); b.record(); torch.cuda.synchronize()
        iso.append(a.elapsed_time(b))
    iso.sort()
    b2b = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            run(i)
        b.record(); torch.cuda.synchronize()
        b2b.append(a.elapsed_time(b) / reps)
    b2b.sort()
    fl = 2.0 * n * h * w * cout * 9 * cin
    return iso[len(iso) // 2] * 1e3, b2b[len(b2b) // 2] * 1e3, fl


if __name__ == "__main__":
    shapes = [(16, 64, 64, 128, 128), (16, 32, 32, 256, 256), (16, 16, 16, 512, 512), (16, 32, 32, 768, 256), (16, 64, 64, 384, 128),
              (16, 128, 128, 64, 128)]
    for shp in shapes:
        for name, opts in (("tc2", {b"tc3": 1}), ("tc3 3-box mt2", {b"tc3": 2, b"tc3_halo": 0, b"tc3_force_mt": 2}),
                           ("tc3 halo mt1", {b"tc3": 2, b"tc3_halo": 1, b"tc3_force_mt": 1}), ("tc3 halo mt2", {b"tc3": 2, b"tc3_halo": 1, b"tc3_force_mt": 2}),
                           ("tc3 halo bn256", {b"tc3": 2, b"tc3_halo": 1, b"tc3_force_bn": 256})):
            for k in (b"tc3", b"tc3_force_bn", b"tc3_force_mt", b"tc3_halo"):
                L.set_option(k, 0)
            for k, v in opts.items():
                L.set_option(k, v)
            iso, b2b, fl = bench(*shp)
            print("shape %-24s %-16s isolated %.1f us (%.0f TF/s)  back-to-back %.1f us (%.0f TF/s)" %
                  (shp, name, iso, fl / iso / 1e6, b2b, fl / b2b / 1e6), flush=True)
    for k in (b"tc3", b"tc3_force_bn", b"tc3_force_mt", b"tc3_halo"):
        L.set_option(k, 0)
