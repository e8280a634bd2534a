"""conv_tc3 phase isolation: tc2_debug bits (1 skip A loads, 2 skip B loads, 8 skip MMAs; results invalid) for the 3-box and
the single-halo-box A schemes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.tc3_halo_bench import L, bench  # noqa

shapes = [(16, 64, 64, 128, 128), (16, 64, 64, 384, 128)]
for shp in shapes:
    for name, opts in (("3-box mt2", {b"tc3": 2, b"tc3_halo": 0, b"tc3_force_mt": 2}), ("halo mt2", {b"tc3": 2, b"tc3_halo": 1, b"tc3_force_mt": 2}),
                       ("halo mt1", {b"tc3": 2, b"tc3_halo": 1, b"tc3_force_mt": 1}), ("3-box mt1", {b"tc3": 2, b"tc3_halo": 0, b"tc3_force_mt": 1})):
        for dbg in (0, 3, 1, 2, 8):
            for k in (b"tc3", b"tc3_force_bn", b"tc3_force_mt", b"tc3_halo"):
                L.set_option(k, 0)
            for k, v in opts.items():
                L.set_option(k, v)
            L.set_option(b"tc2_debug", dbg)
            iso, b2b, fl = bench(*shp)
            print("shape %-24s %-10s dbg %d  isolated %.1f us  back-to-back %.1f us" % (shp, name, dbg, iso, b2b), flush=True)
L.set_option(b"tc2_debug", 0)
