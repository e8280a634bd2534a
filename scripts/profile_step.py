"""Run N eager training steps of the bench workload (no CUDA graph) so ncu can list every libstp launch of a step.
   ncu --metrics gpu__time_duration.sum --clock-control none -s <launches_per_step*2> -c <launches_per_step> --csv ...
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C2_AUGMENT, synth_pool
from segmentation_training_pipeline_b200.models import SegNet
from segmentation_training_pipeline_b200.trainer import AugmentConfig, Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--backbone", default="resnet34")
ap.add_argument("--config", default="c2", choices=["c2", "c3", "people"], help="c3: FPN/resnet50 3-class Lovasz (BASELINE configs[2]); people: DeepLabV3/mobilenetv2 320x320 (examples/people/people.yaml)")
a = ap.parse_args()
torch.cuda.set_device(0)
if a.config == "people":
    a.size = 320
    net = SegNet("mobilenetv2", classes=1, input_shape=(a.size, a.size, 3), batch=a.batch, device="cuda:0", seed=0, loss=(1.0, 0.0, 0.0),
                 architecture="DeepLabV3")
    tr = Trainer(net, optimizer="Adam", lr=1e-3, augment=AugmentConfig(seed=0, fliplr=0.5))
elif a.config == "c3":
    net = SegNet("resnet50", classes=3, input_shape=(a.size, a.size, 3), batch=a.batch, device="cuda:0", seed=0, loss=(0.0, 0.0, 0.0, 1.0),
                 architecture="FPN")
    tr = Trainer(net, optimizer="Adam", lr=1e-3, augment=AugmentConfig(seed=0, **C2_AUGMENT))
else:
    net = SegNet(a.backbone, classes=1, input_shape=(a.size, a.size, 3), batch=a.batch, device="cuda:0", seed=0, loss=(1.0, 1.0, 0.0))
    tr = Trainer(net, optimizer="Adam", lr=1e-3, augment=AugmentConfig(seed=0, **C2_AUGMENT))
img, mask = synth_pool(a.batch, a.size, a.size, 1234, 4321)
if a.config == "c3":
    import numpy as np
    mask = np.concatenate([mask, np.roll(mask, a.size // 8, axis=2), np.roll(mask, a.size // 4, axis=1)], axis=3)
tr.set_pool(torch.from_numpy(img), torch.from_numpy(mask))
l0 = net.L.launch_count()
for i in range(a.steps):
    tr.step_eager()
    torch.cuda.synchronize()
    if i == 0:
        print("launches_per_step", net.L.launch_count() - l0, "tc", net.L.tc_launch_count())
print("loss", tr.loss_value())
