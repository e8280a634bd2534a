"""Critical-path knock-out experiment: time the captured U-Net/ResNet-34 bs16 512^2 step with whole kernel categories removed
(STP_SKIP, engine.py; results are numerically invalid) -> what each category costs IN the overlapped step, as opposed to its
serialised share in the ncu launch list.   python scripts/knockout.py > gpurun_out/knockout.txt"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, torch
sys.path.insert(0, %r)
from bench import synth_pool
from segmentation_training_pipeline_b200.models import SegNet
from segmentation_training_pipeline_b200.trainer import AugmentConfig, Trainer
net = SegNet("resnet34", classes=1, input_shape=(512, 512, 3), batch=16, device="cuda:0", seed=0, loss=(1.0, 1.0, 0.0))
img, mask = synth_pool(64, 512, 512, 1234, 4321)
tr = Trainer(net, optimizer="Adam", lr=1e-3, augment=AugmentConfig(fliplr=0.5, flipud=0.5, affine=True, scale=(0.8, 1.5), translate_x=(-0.2, 0.2),
             translate_y=(-0.2, 0.2), rotate=(-16, 16), shear=(-16, 16), multiply=(0.8, 1.2), add=(-10, 10)))
tr.set_pool(torch.from_numpy(img), torch.from_numpy(mask))
tr.capture()
for _ in range(5): tr.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30): tr.step()
b.record(); torch.cuda.synchronize()
print("%%.3f" %% (a.elapsed_time(b) / 30))
''' % ROOT
base = None
for skip in ["", "wgrad", "bn_apply", "bn_bwd_reduce", "bn_bwd_apply", "bn_bwd_reduce,bn_bwd_apply", "dgrad", "conv_fwd",
             "wgrad,dgrad,conv_fwd", "bn_apply,bn_bwd_reduce,bn_bwd_apply", "wgrad,bn_bwd_reduce,bn_bwd_apply"]:
    env = dict(os.environ, STP_SKIP=skip)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    try:
        ms = float(r.stdout.strip().splitlines()[-1])
    except Exception:
        print("skip=%-40s FAILED: %s" % (skip, (r.stderr or r.stdout)[-300:]))
        continue
    if base is None:
        base = ms
    print("skip=%-40s step %.3f ms  (saves %.3f ms)" % (skip or "(nothing)", ms, base - ms), flush=True)
