"""A/B micro-benchmark of the CTA-pair conv kernel (conv_tc3) against the single-CTA halo kernel (conv_tc2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.conv_bench import L, bench  # noqa: E402  (prints its default table on import: suppressed via argv)

shapes = [(16, 64, 64, 128, 128), (16, 32, 32, 256, 256), (16, 16, 16, 512, 512), (16, 32, 32, 768, 256), (16, 64, 64, 384, 128)]
for shp in shapes:
    for name, opts in (("tc2", {b"tc3": 1}), ("tc3 auto", {b"tc3": 2}), ("tc3 bn128 mt1", {b"tc3": 2, b"tc3_force_bn": 128, b"tc3_force_mt": 1}),
                       ("tc3 bn128 mt2", {b"tc3": 2, b"tc3_force_bn": 128, b"tc3_force_mt": 2}), ("tc3 bn256", {b"tc3": 2, b"tc3_force_bn": 256})):
        for k in (b"tc3", b"tc3_force_bn", b"tc3_force_mt"):
            L.set_option(k, 0)
        for k, v in opts.items():
            L.set_option(k, v)
        us, tf = bench(*shp)
        print("shape", shp, "%-14s us %.1f TF %.0f" % (name, us, tf), flush=True)
