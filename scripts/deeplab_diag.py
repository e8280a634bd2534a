"""Depth profile of the DeepLabV3 parity: per tapped layer, engine vs bf16-storage oracle and the bf16-vs-fp32 oracle floor."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle.models import SegModel  # noqa: E402
from segmentation_training_pipeline_b200.models import SegNet  # noqa: E402
from segmentation_training_pipeline_b200.trainer import Trainer  # noqa: E402
from tests.test_gpu_model import _data, _perturb  # noqa: E402

n, size = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 64
net = SegNet("mobilenetv2", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(1.0, 0.0, 0.0),
             architecture="DeepLabV3", dropout=0.0)
W = _perturb(net.get_weights())
net.set_weights(W)
tr = Trainer(net)
img, mask = _data(n, size, size)
tr.set_batch(img.cuda(), mask.cuda())
net.prep_weights()
net.forward()
torch.cuda.synchronize()
oms = {}
for st in ("bf16", "fp32"):
    om = SegModel("DeepLabV3", "mobilenetv2", classes=1, input_shape=(size, size, 3), storage=st, update_moving=False)
    om.load_numpy(W)
    with torch.no_grad():
        om(img.float())
    oms[st] = om
for name in oms["bf16"].taps:
    if name not in net.bufs:
        continue
    e = net.bufs[name].torch().float().cpu()
    b = oms["bf16"].taps[name].permute(0, 2, 3, 1)
    f = oms["fp32"].taps[name].permute(0, 2, 3, 1)
    if e.shape != b.shape:
        continue
    print("%-32s engine-vs-bf16-oracle %.4f   bf16-vs-fp32 floor %.4f   engine-vs-fp32 %.4f" %
          (name, float((e - b).norm() / b.norm()), float((b - f).norm() / f.norm()), float((e - f).norm() / f.norm())))
