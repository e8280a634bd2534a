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


# --------------------------------------------------------------------------------------------------------------------
# Declarative `datasets:` entries with channel bindings (reference examples/people/ds_2.yaml:43-66, ds_3.yaml:43-76;
# musket_core.datasources [DEP, unpinned]).  An entry is either the short form {input_path, output_path} or
#   inputs / outputs: [ {name, data_type, bindings: [ {reader, path, bind: [source channels], treat: {type, colors}} ]} ]
# Every binding reads <path>/<stem>.* with its reader, applies `treat`, and APPENDS the source channels listed in `bind` to
# the tensor being assembled ("0, 1 channels of image will be written into 0, 1 channels of input tensor", then the next
# binding's [2] becomes channel 2; ds_3.yaml's output `bind: [3]` picks the 4th channel of the treated tensor as output 0).
# --------------------------------------------------------------------------------------------------------------------
_IMG_EXT = (".jpg", ".jpeg", ".png", ".bmp")


def _read(path_dir: str, stem: str, reader: str) -> np.ndarray:
    """reader RGBA -> HxWx4 uint8 (alpha 255 when the file has none); monochrome -> HxWx1 uint8"""
    import cv2
    f = None
    for ext in _IMG_EXT:
        cand = os.path.join(path_dir, stem + ext)
        if os.path.exists(cand):
            f = cand
            break
    if f is None:
        raise FileNotFoundError("no image for '%s' in %s" % (stem, path_dir))
    r = str(reader).lower()
    if r == "monochrome":
        m = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
        if m is None:
            raise IOError("cannot read " + f)
        return m[:, :, None]
    if r == "rgba":
        m = cv2.imread(f, cv2.IMREAD_UNCHANGED)
        if m is None:
            raise IOError("cannot read " + f)
        if m.ndim == 2:
            m = np.repeat(m[:, :, None], 3, axis=2)
        if m.shape[2] == 3:
            m = np.concatenate([m[:, :, ::-1], np.full(m.shape[:2] + (1,), 255, m.dtype)], axis=2)   # BGR -> RGB + opaque alpha
        else:
            m = np.concatenate([m[:, :, 2::-1], m[:, :, 3:4]], axis=2)                                   # BGRA -> RGBA
        if m.dtype != np.uint8:
            m = (m >> 8).astype(np.uint8) if m.dtype == np.uint16 else m.astype(np.uint8)
        return np.ascontiguousarray(m)
    raise NotImplementedError("datasets: reader '%s' is not built (RGBA, monochrome)" % reader)


def _treat(a: np.ndarray, treat) -> np.ndarray:
    t = dict(treat or {})
    kind = t.get("type", t.get('type"', "as_is"))     # the reference's own examples spell the key `type":` (ds_2.yaml:54)
    if kind == "as_is":
        return a
    if kind == "binary_mask":
        colors = t.get("colors")
        if colors:   # (h, w, len(colors)) tensor: 1 where the pixel's RGB equals the listed colour, else 0 (ds_3.yaml:73-76)
            rgb = a[:, :, :3] if a.shape[2] >= 3 else np.repeat(a[:, :, :1], 3, axis=2)
            return np.stack([(rgb == np.asarray(c, dtype=rgb.dtype)[None, None, :]).all(axis=2) for c in colors], axis=2).astype(np.uint8)
        return (a > 0).astype(np.uint8)
    raise NotImplementedError("datasets: treat type '%s' is not built (as_is, binary_mask)" % kind)


class BoundDataSet:
    """A `datasets:` entry written with inputs / outputs / bindings (see the block comment above).  Item ids are the file
    stems of the first input binding's folder; x = uint8 [H, W, sum(len(bind))], y = uint8 [H, W, ...] in {0,1} for
    binary_mask outputs."""

    def __init__(self, spec: dict, base_dir: str = "."):
        ins, outs = spec.get("inputs") or [], spec.get("outputs") or []
        if len(ins) != 1 or len(outs) != 1:
            raise NotImplementedError("datasets: exactly one input and one output tensor are built (the network has one of each)")
        self.in_b = [self._binding(b, base_dir) for b in ins[0].get("bindings") or []]
        self.out_b = [self._binding(b, base_dir) for b in outs[0].get("bindings") or []]
        if not self.in_b or not self.out_b:
            raise ValueError("datasets: an input / output without bindings")
        first = self.in_b[0]["path"]
        if not os.path.isdir(first):
            raise FileNotFoundError(first)
        self.ids: List[str] = sorted({os.path.splitext(f)[0] for f in os.listdir(first) if f.lower().endswith(_IMG_EXT)})
        if not self.ids:
            raise ValueError("no images in " + first)

    @staticmethod
    def _binding(b: dict, base: str) -> dict:
        for k in ("reader", "path", "bind"):
            if k not in b:
                raise ValueError("datasets: binding without '%s'" % k)
        return {"reader": b["reader"], "path": b["path"] if os.path.isabs(b["path"]) else os.path.join(base, b["path"]),
                "bind": [int(i) for i in b["bind"]], "treat": b.get("treat")}

    def __len__(self):
        return len(self.ids)

    def _assemble(self, bindings, stem) -> np.ndarray:
        parts = []
        for b in bindings:
            a = _treat(_read(b["path"], stem, b["reader"]), b["treat"])
            for ch in b["bind"]:
                if not 0 <= ch < a.shape[2]:
                    raise IndexError("datasets: bind channel %d of a %d-channel tensor (%s)" % (ch, a.shape[2], b["path"]))
                parts.append(a[:, :, ch])
        return np.ascontiguousarray(np.stack(parts, axis=2))

    def __getitem__(self, i) -> PredictionItem:
        stem = self.ids[int(i)]
        return PredictionItem(stem, self._assemble(self.in_b, stem), self._assemble(self.out_b, stem))

    def isPositive(self, i) -> bool:
        return bool(self[i].y.any())


def dataset_from_spec(spec: dict, base_dir: str = "."):
    """`datasets:` entry -> dataset object: short form {input_path, output_path} (ds_1.yaml:33-38) or the bindings form."""
    if "inputs" in spec or "outputs" in spec:
        return BoundDataSet(spec, base_dir)
    if "input_path" in spec and "output_path" in spec:
        return SimplePNGMaskDataSet(os.path.join(base_dir, spec["input_path"]), os.path.join(base_dir, spec["output_path"]))
    raise ValueError("datasets: an entry needs input_path + output_path or inputs + outputs")
