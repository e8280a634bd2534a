"""Diagnostic (GPU): per-layer engine vs oracle error, to separate bf16 rounding noise from semantic bugs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.models import SegModel
from oracle import losses as OL
from segmentation_training_pipeline_b200.models import SegNet
from segmentation_training_pipeline_b200.trainer import Trainer
from tests.test_gpu_model import _data, _perturb

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))

for backbone, size, n, tc in [("resnet18", 64, 2, 1), ("resnet18", 64, 2, 0), ("resnet18", 128, 4, 1), ("resnet34", 128, 4, 1), ("resnet34", 256, 2, 1)]:
    net = SegNet(backbone, classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(1.0, 1.0, 0.0))
    net.L.set_tc_enabled(tc)
    W = _perturb(net.get_weights()); net.set_weights(W)
    tr = Trainer(net)
    img, mask = _data(n, size, size)
    tr.set_batch(img.cuda(), mask.cuda())
    net.prep_weights(); net.forward(); net.backward(); torch.cuda.synchronize()
    outs = {}
    for storage in ("bf16", "fp32"):
        om = SegModel("Unet", backbone, classes=1, input_shape=(size, size, 3), storage=storage, update_moving=False)
        om.load_numpy(W)
        y = om(img.float())
        outs[storage] = {k: v.detach().clone() for k, v in om.taps.items()}
    print("==", backbone, size, "n", n, "tc", tc)
    for k in ["relu0", "stage1_unit1_relu1", "stage2_unit1_relu1", "stage3_unit1_relu1", "stage4_unit1_relu1", "relu1", "logits"]:
        if k not in outs["bf16"]: continue
        ob, of = outs["bf16"][k], outs["fp32"][k]
        if k == "logits":
            e = net.head.logits.cpu().view(n, size, size, 1).permute(0, 3, 1, 2)
        elif k == "relu1":
            e = net.bufs["relu1_up"].torch().float().cpu().permute(0, 3, 1, 2)[:, :, ::2, ::2]
        else:
            e = net.bufs[k].torch().float().cpu().permute(0, 3, 1, 2)
        print("  %-22s engine-vs-orc_bf16 %.4f  engine-vs-fp32 %.4f  orc_bf16-vs-fp32 %.4f" % (k, rel(e, ob), rel(e, of), rel(ob, of)))
    net.L.set_tc_enabled(1)
