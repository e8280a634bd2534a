"""DeepLabV3+/MobileNetV2 (the reference's example experiment: shape 320, batch 16, binary_crossentropy, Adam): step time under
the whole-step CUDA graph and an eager per-op breakdown (CUDA events around every op's fwd / bwd).

    python scripts/deeplab_bench.py [--size 320] [--batch 16] [--steps 30]
"""
import argparse
import collections
import json

import numpy as np
import torch

from segmentation_training_pipeline_b200.models import SegNet
from segmentation_training_pipeline_b200.trainer import Trainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--opt", action="append", default=[], help="libstp option name=value (stp_set_option), repeatable")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--tc", type=int, default=1, help="0: disable the tcgen05 conv kernels (generic mma.sync everywhere)")
    a = ap.parse_args()
    S, B = a.size, a.batch
    net = SegNet("mobilenetv2", classes=1, input_shape=(S, S, 3), batch=B, device="cuda:0", seed=0, loss=(1.0, 0.0, 0.0),
                 architecture="DeepLabV3")
    net.L.set_tc_enabled(a.tc)
    for o in a.opt:
        k, v = o.split("=")
        net.L.set_option(k.encode(), int(v))
    tr = Trainer(net, optimizer="Adam", lr=1e-3)
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (2 * B, S, S, 3), dtype=np.uint8))
    mask = torch.from_numpy((rng.random((2 * B, S, S, 1)) > 0.7).astype(np.uint8))
    tr.set_pool(img, mask)
    tr.capture()
    for _ in range(5):
        tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = net.L.launch_count()
    e0.record()
    for _ in range(a.steps):
        tr.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"workload": "DeepLabV3/mobilenetv2 %dx%d bs%d train step (graph)" % (S, S, B), "ms_per_step": ms,
                      "img_per_s": B / ms * 1e3, "loss": tr.loss_value()}))
    if a.no_breakdown:
        return
    # ---- eager per-op breakdown -------------------------------------------------------------------
    net.overlap_wgrad = False
    st = torch.cuda.current_stream()
    acc = collections.defaultdict(float)
    per = collections.defaultdict(float)
    R = 20   # back-to-back launches per op: the queue stays full, so the events bracket GPU time, not Python call latency
    net.prep_weights()
    net.forward()
    net.backward()
    for op in net.ops:
        for ph, fn in (("fwd", op.fwd), ("bwd", op.bwd)):
            fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(st)
            for _ in range(R):
                fn()
            e.record(st)
            torch.cuda.synchronize()
            t = s.elapsed_time(e) / R
            acc[type(op).__name__ + "." + ph] += t
            nm = getattr(op, "name", None) or getattr(getattr(op, "y", None), "name", "") or type(op).__name__
            shape = ""
            if hasattr(op, "x") and hasattr(op, "y") and hasattr(op.y, "c"):
                shape = " %dx%d %d->%d" % (op.y.h, op.y.w, op.x.c, op.y.c)
            per["%s %s.%s%s" % (nm, type(op).__name__, ph, shape)] += t
    tot = sum(acc.values())
    print("sum of per-op times (each op timed over %d back-to-back launches) %.3f ms" % (R, tot))
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print("  %-22s %7.3f ms %5.1f %%" % (k, v, 100 * v / tot))
    print("top ops:")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:a.top]:
        print("  %-70s %7.3f ms" % (k, v))


if __name__ == "__main__":
    main()
