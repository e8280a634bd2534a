"""Sweep the strip height (MT) of the single-CTA halo kernel per layer shape (tc2_force_mt) to check the cost model's choice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.conv_bench import L, bench  # noqa: E402

shapes = [(16, 128, 128, 64, 64), (16, 64, 64, 128, 128), (16, 32, 32, 256, 256), (16, 16, 16, 512, 512), (16, 128, 128, 192, 64),
          (16, 256, 256, 128, 32), (16, 256, 256, 32, 32), (16, 512, 512, 32, 16), (16, 512, 512, 16, 16), (16, 64, 64, 384, 128),
          (16, 32, 32, 768, 256)]
L.set_option(b"tc3", 1)
for shp in shapes:
    row = []
    for mt in (0, 1, 2, 4, 8):
        L.set_option(b"tc2_force_mt", mt)
        try:
            us, tf = bench(*shp, reps=7)
            row.append("mt%d %.1f" % (mt, us))
        except Exception as e:
            row.append("mt%d err" % mt)
    print(shp, "  ".join(row), flush=True)
L.set_option(b"tc2_force_mt", 0)
