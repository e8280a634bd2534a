"""Generates tests/golden/augment_golden.npz: inputs + outputs of the augmentation stage computed with the REAL
cv2.warpAffine backend (cv2 4.13 in this container; the backend imgaug's Affine calls in the reference).
Run from the repo root:  python tests/golden/make_augment_golden.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import augment as OA

rng = np.random.default_rng(2024)
N, H, W = 6, 64, 64
images = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
masks = (rng.random((N, H, W, 1)) > 0.65).astype(np.uint8)
spec = OA.AugSpec(fliplr=0.5, flipud=0.5, affine=True, scale=(0.8, 1.5), translate_x=(-0.2, 0.2), translate_y=(-0.2, 0.2),
                  rotate=(-16, 16), shear=(-16, 16), multiply=(0.8, 1.2), add=(-10, 10))
seed, step = 1234, 5
oi, om = OA.augment_batch(images, masks, spec, seed, step, use_cv2=True)
mats = np.stack([OA.draw_params(spec, seed, step, n, H, W).matrix for n in range(N)])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "augment_golden.npz"), images=images, masks=masks,
                    out_images=oi, out_masks=om, matrices=mats, seed=seed, step=step)
print("written", oi.shape, om.shape)
