"""Secondary measurements (NOT the headline bench line): captured training step of the other BASELINE.json configs that fit
one GPU -- configs[2] FPN/ResNet-50 512x512 3-class Lovasz bs16, and Linknet / transposed-decoder U-Net variants -- with a
device-resident synthetic pool, on-device augmentation, CUDA events around K graph replays."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from segmentation_training_pipeline_b200.models import SegNet
from segmentation_training_pipeline_b200.trainer import AugmentConfig, Trainer

AUG = dict(fliplr=0.5, flipud=0.5, affine=True, scale=(0.8, 1.5), translate_x=(-0.2, 0.2), translate_y=(-0.2, 0.2),
           rotate=(-16, 16), shear=(-16, 16), multiply=(0.8, 1.2), add=(-10, 10))
CASES = [("configs[2] FPN/resnet50 3-class lovasz_loss", dict(architecture="FPN", backbone="resnet50", classes=3, loss=(0.0, 0.0, 0.0, 1.0))),
         ("FPN/resnet34 1-class bce+dice", dict(architecture="FPN", backbone="resnet34", classes=1, loss=(1.0, 1.0, 0.0))),
         ("Linknet/resnet34 1-class bce+dice", dict(architecture="Linknet", backbone="resnet34", classes=1, loss=(1.0, 1.0, 0.0))),
         ("Unet/resnet50 1-class bce+dice", dict(architecture="Unet", backbone="resnet50", classes=1, loss=(1.0, 1.0, 0.0)))]
S, B, P, K = 512, 16, 32, 10
for name, kw in CASES:
    net = SegNet(input_shape=(S, S, 3), batch=B, device="cuda:0", seed=0, **kw)
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (P, S, S, 3), generator=g, dtype=torch.uint8)
    mask = (torch.rand(P, S // 32, S // 32, kw["classes"], generator=g) > 0.7).to(torch.uint8)
    mask = mask.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    tr = Trainer(net, optimizer="Adam", lr=1e-3, augment=AugmentConfig(seed=0, **AUG))
    tr.set_pool(img, mask)
    l0 = net.L.launch_count()
    tr.capture()
    lps = (net.L.launch_count() - l0) // 2
    for _ in range(3):
        tr.step()
    first = tr.loss_value()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        tr.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(json.dumps({"config": name, "size": S, "batch": B, "ms_per_step": ms, "img_per_s": B / ms * 1e3, "launches_per_step": lps,
                      "loss_after_3": first, "loss_after_%d" % (3 + K): tr.loss_value(),
                      "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    del tr, net
    torch.cuda.empty_cache()
