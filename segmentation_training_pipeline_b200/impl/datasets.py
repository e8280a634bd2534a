"""Dataset protocol of the reference (`from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet,
PredictionItem`; reference impl/datasets.py:1 re-exports musket_core.datasets [DEP]; usage README.md:116-126, 313,
336-337).  A dataset is anything with __len__ and __getitem__(i) -> PredictionItem(id, x: HxWx3 uint8, y: HxWx1 {0,1}),
optionally isPositive(i)."""
from __future__ import annotations

import os
from typing import List

import numpy as np


class PredictionItem:
    def __init__(self, path, x, y, prediction=None):
        self.x = x
        self.y = y
        self.id = path
        self.prediction = prediction

    def original(self):
        return self

    def rootItem(self):
        return self

    def item_id(self):
        return self.id


class SimplePNGMaskDataSet:
    """Images from `path` (jpg/png), masks from `mask_path` (<stem>.png).  x = RGB uint8 (values 0..255, fed to the
    network unscaled exactly as the reference does); y = (png > 0) as HxWx1 uint8 in {0,1} (reference ds_1.yaml:37-38)."""

    EXT = (".jpg", ".jpeg", ".png", ".bmp")

    def __init__(self, path, mask_path, detect_size=False, in_ext="jpg", out_ext="png", generate=False):
        self.path, self.mask_path, self.out_ext = path, mask_path, out_ext
        if not os.path.isdir(path):
            raise FileNotFoundError(path)
        self.ids: List[str] = sorted(f for f in os.listdir(path) if f.lower().endswith(self.EXT))
        if not self.ids:
            raise ValueError("no images in " + path)

    def __len__(self):
        return len(self.ids)

    def _mask_file(self, name):
        stem = os.path.splitext(name)[0]
        for ext in (self.out_ext, "png", "jpg"):
            p = os.path.join(self.mask_path, stem + "." + ext)
            if os.path.exists(p):
                return p
        raise FileNotFoundError("mask for %s in %s" % (name, self.mask_path))

    def __getitem__(self, i) -> PredictionItem:
        import cv2
        name = self.ids[i]
        img = cv2.imread(os.path.join(self.path, name), cv2.IMREAD_COLOR)
        if img is None:
            raise IOError("cannot read " + name)
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        m = cv2.imread(self._mask_file(name), cv2.IMREAD_UNCHANGED)
        if m is None:
            raise IOError("cannot read mask of " + name)
        if m.ndim == 3:
            m = m.sum(axis=2)
        y = (m > 0).astype(np.uint8)[:, :, None]
        return PredictionItem(os.path.splitext(name)[0], img, y)

    def isPositive(self, i) -> bool:
        return bool(self[i].y.any())
