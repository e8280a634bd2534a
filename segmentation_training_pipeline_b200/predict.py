"""Forward-only inference behind the reference's prediction verbs (reference segmentation.py:62-91
predict_to_directory / predict_in_directory, :158-191 evaluateAll; README.md:493-534, 745-754): images of a directory are
resized to `shape`, run through the engine graph in inference mode (moving BatchNorm statistics), optionally averaged over
the four flip variants (`ttflips`) and over several folds (`fold` may be a list: ensembling).  Like the reference, callbacks
and writers get the FLOAT probability map scaled back to the image's own size as a SegmentationMapOnImage (`.arr`; users
threshold it themselves, README.md:505-513); `PredictionBatch.segmentation_maps_aug` / evaluateAll's `.results` keep the
0.5-thresholded masks (nearest-neighbour scale-back) for convenience."""
from __future__ import annotations

import os
from typing import Callable, Iterator, List, Optional, Sequence, Union

import numpy as np

IMG_EXT = (".jpg", ".jpeg", ".png", ".bmp")


class SegmentationMapOnImage:
    """What the reference hands to prediction callbacks and stores in its batches: an imgaug.SegmentationMapOnImage built from
    the network's FLOAT output (reference segmentation.py:58-60 `update`, :62-91) -- user code reads `.arr` and thresholds it
    itself (README.md:505-513 `img.arr > threshold`).  `.arr` = float32 probabilities [h, w, classes]; np.asarray(obj) and
    get_arr_int() give the 0.5-thresholded mask."""

    def __init__(self, arr, shape=None):
        a = np.asarray(arr, dtype=np.float32)
        self.arr = a[:, :, None] if a.ndim == 2 else a
        self.shape = tuple(shape) if shape is not None else self.arr.shape

    def get_arr_int(self, threshold=0.5):
        return (self.arr > threshold).astype(np.int32)

    def __array__(self, dtype=None, copy=None):
        m = (self.arr > 0.5).astype(np.uint8)
        return m.astype(dtype) if dtype is not None else m

    def resize(self, sizes, interpolation="cubic"):
        """imgaug 0.3.0 resizes float segmentation maps like heatmaps (cubic, clipped to [0, 1]) [DEP, unpinned]."""
        import cv2
        h, w = int(sizes[0]), int(sizes[1])
        if self.arr.shape[:2] == (h, w):
            return SegmentationMapOnImage(self.arr)
        flag = {"cubic": cv2.INTER_CUBIC, "linear": cv2.INTER_LINEAR, "nearest": cv2.INTER_NEAREST}[interpolation]
        out = cv2.resize(self.arr, (w, h), interpolation=flag)
        return SegmentationMapOnImage(np.clip(out, 0.0, 1.0))


class PredictionBatch:
    """What the reference yields per batch (an imgaug.Batch): ids, original images, probability maps and binary
    segmentation maps at network resolution."""

    def __init__(self, data, images, probs, shape):
        self.data, self.images, self.probabilities = data, images, probs
        self.segmentation_maps_aug = [(p > 0.5).astype(np.uint8) for p in probs]
        self.shape = shape


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def predict_arrays(net, images_u8: np.ndarray, ttflips=False) -> np.ndarray:
    """uint8 [n, H, W, 3] at network resolution -> float32 probabilities [n, H, W, classes]; n <= net.batch."""
    import torch
    n = images_u8.shape[0]
    B = net.batch
    H, W, CI = net.input_shape
    assert n <= B and images_u8.shape[1:] == (H, W, CI), (images_u8.shape, net.input_shape, B)
    variants = [(False, False)] + ([(True, False), (False, True), (True, True)] if ttflips else [])
    acc = np.zeros((n, H, W, net.classes), np.float32)
    was_training = net.training
    net.training = False
    try:
        for fl, fu in variants:
            x = images_u8
            if fl:
                x = x[:, :, ::-1]
            if fu:
                x = x[:, ::-1]
            batch = np.zeros((B, H, W, CI), np.uint8)
            batch[:n] = x
            net.img.storage.copy_(torch.from_numpy(batch).reshape(-1).to(net.device))
            net.prep_weights()
            enabled, net.loss.enabled = net.loss.enabled, False
            net.forward()
            net.loss.enabled = enabled
            z = net.head.logits.detach().cpu().numpy().reshape(B, H, W, net.classes)[:n]
            act = getattr(net, "activation", "sigmoid")
            if act == "softmax":
                e = np.exp(z - z.max(axis=-1, keepdims=True))
                p = e / e.sum(axis=-1, keepdims=True)
            elif act in ("linear", "none"):
                p = z
            else:
                p = _sigmoid(z)
            if fl:
                p = p[:, :, ::-1]
            if fu:
                p = p[:, ::-1]
            acc += p
    finally:
        net.training = was_training
    return acc / len(variants)


def _net_hw(cfg):
    """(H, W) the network runs at: `shape`, or the CELL shape under `crops: N` (reference createNet1, segmentation.py:131-132)."""
    crops = int(getattr(cfg, "crops", 0) or 0)
    H, W = int(cfg.shape[0]), int(cfg.shape[1])
    return (H // crops, W // crops) if crops > 1 else (H, W)


def _list_images(spath) -> List[str]:
    return sorted(f for f in os.listdir(spath) if f.lower().endswith(IMG_EXT))


def predict_on_directory(cfg, spath, fold: Union[int, Sequence[int]] = 0, stage=0, limit=-1, batch_size=32,
                         ttflips=False) -> Iterator[PredictionBatch]:
    import cv2
    folds = list(fold) if isinstance(fold, (list, tuple)) else [fold]
    nets = [cfg.load_model(f, stage) for f in folds]
    B = min(int(batch_size), nets[0].batch)
    H, W = _net_hw(cfg)
    names = _list_images(spath)
    if limit is not None and limit > 0:
        names = names[:limit]
    crops = int(getattr(cfg, "crops", 0) or 0)
    if crops > 1:   # `crops: N`: every image is predicted cell by cell and assembled back at its own size
        from .crops import predict_image_by_cells
        fn = lambda x: sum(predict_arrays(net, x, ttflips) for net in nets) / len(nets)
        for nm in names:
            img = cv2.cvtColor(cv2.imread(os.path.join(spath, nm), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
            yield PredictionBatch([nm], [img], [predict_image_by_cells(fn, img, crops, (H, W), B)], img.shape[:2])
        return
    for s in range(0, len(names), B):
        ids = names[s:s + B]
        origs, xs = [], []
        for nm in ids:
            img = cv2.cvtColor(cv2.imread(os.path.join(spath, nm), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
            origs.append(img)
            xs.append(cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC) if img.shape[:2] != (H, W) else img)
        x = np.stack(xs).astype(np.uint8)
        probs = sum(predict_arrays(net, x, ttflips) for net in nets) / len(nets)   # fold ensembling: mean probability
        yield PredictionBatch(ids, origs, list(probs), (H, W))


def _scale_back(seg: np.ndarray, orig) -> np.ndarray:
    import cv2
    h, w = orig.shape[:2]
    if seg.shape[:2] == (h, w):
        return seg
    out = cv2.resize(seg, (w, h), interpolation=cv2.INTER_NEAREST)
    return out[:, :, None] if out.ndim == 2 else out


def _scaled_map(prob: np.ndarray, orig) -> SegmentationMapOnImage:
    """the probability map of one image as the reference's Scale({"height": orig.h, "width": orig.w}) leaves it"""
    return SegmentationMapOnImage(prob).resize(np.asarray(orig).shape[:2])


def predict_to_directory(cfg, spath, tpath, fold=0, stage=0, limit=-1, batchSize=32, binaryArray=False, ttflips=False):
    """reference segmentation.py:62-79: per image the scaled map's `.arr` (float probabilities) is stored as <stem>.npy
    (binaryArray=True; what ansemblePredictions sums) or as the 8-bit image arr*255."""
    import cv2
    os.makedirs(tpath, exist_ok=True)
    n = 0
    for b in predict_on_directory(cfg, spath, fold, stage, limit, batchSize, ttflips):
        for i, id_ in enumerate(b.data):
            m = _scaled_map(b.probabilities[i], b.images[i]).arr
            stem = id_[0:id_.index(".")]
            if binaryArray:
                np.save(os.path.join(tpath, stem), m)
            else:
                img8 = (m * 255).astype(np.uint8)
                cv2.imwrite(os.path.join(tpath, stem + ".png"), img8[:, :, 0] if img8.shape[2] == 1 else img8)
            n += 1
    return n


def predict_in_directory(cfg, spath, fold, stage, cb: Callable, data, limit=-1, batchSize=32, ttflips=False):
    """reference segmentation.py:81-91: cb(file name, SegmentationMapOnImage scaled to the image's size, data)."""
    for b in predict_on_directory(cfg, spath, fold, stage, limit, batchSize, ttflips):
        for i, id_ in enumerate(b.data):
            cb(id_, _scaled_map(b.probabilities[i], b.images[i]), data)


def ansemble_predictions(source_folder, folders: Sequence[str], cb: Callable, data, weights: Optional[Sequence[float]] = None):
    """segmentation.ansemblePredictions (reference segmentation.py:27 -> musket_core.generic_config [DEP]; README.md:745-754):
    for every image of `source_folder` the <stem>.npy probability arrays that predict_to_directory(..., binaryArray=True)
    stored in each of `folders` are averaged (optionally weighted) and cb(file name, SegmentationMapOnImage, data) is called."""
    ws = [1.0] * len(folders) if weights is None else [float(w) for w in weights]
    if len(ws) != len(folders) or not folders:
        raise ValueError("ansemblePredictions: one weight per folder")
    for name in _list_images(source_folder):
        stem = name[0:name.index(".")]
        acc = None
        for f, w in zip(folders, ws):
            a = np.load(os.path.join(f, stem + ".npy")).astype(np.float32) * w
            acc = a if acc is None else acc + a
        cb(name, SegmentationMapOnImage(acc / sum(ws)), data)


def evaluate_all(cfg, ds, fold=None, stage=-1, negatives="real", ttflips=None, batchSize=32) -> Iterator[PredictionBatch]:
    """Predictions for every item of a dataset (reference evaluateAll, segmentation.py:158-191): yields batches whose
    `.data` are the PredictionItems and whose `.results` are the binary maps scaled back to each item's size."""
    import cv2
    folds = list(range(cfg.folds_count)) if fold is None else (list(fold) if isinstance(fold, (list, tuple)) else [fold])
    nets = [cfg.load_model(f, stage) for f in folds]
    B = min(int(batchSize), nets[0].batch)
    H, W = _net_hw(cfg)
    idx = list(range(len(ds)))
    if negatives == "none" and hasattr(ds, "isPositive"):
        idx = [i for i in idx if ds.isPositive(i)]
    crops = int(getattr(cfg, "crops", 0) or 0)
    if crops > 1:
        from .crops import predict_image_by_cells
        fn = lambda x: sum(predict_arrays(net, x, bool(ttflips)) for net in nets) / len(nets)
        for i in idx:
            it = ds[i]
            img = np.asarray(it.x)
            b = PredictionBatch([it], [img], [predict_image_by_cells(fn, img, crops, (H, W), B)], img.shape[:2])
            b.results = list(b.segmentation_maps_aug)
            b.predicted_maps_aug = [SegmentationMapOnImage(p) for p in b.probabilities]
            b.segmentation_maps = [SegmentationMapOnImage(np.asarray(it.y, dtype=np.float32))]
            yield b
        return
    for s in range(0, len(idx), B):
        items = [ds[i] for i in idx[s:s + B]]
        xs = [cv2.resize(np.asarray(it.x), (W, H), interpolation=cv2.INTER_CUBIC) if np.asarray(it.x).shape[:2] != (H, W)
              else np.asarray(it.x) for it in items]
        x = np.stack(xs).astype(np.uint8)
        probs = sum(predict_arrays(net, x, bool(ttflips)) for net in nets) / len(nets)
        b = PredictionBatch(items, [np.asarray(it.x) for it in items], list(probs), (H, W))
        b.results = [_scale_back(m, np.asarray(it.x)) for m, it in zip(b.segmentation_maps_aug, items)]
        # the reference's batch fields (segmentation.py:182-184): ground-truth maps and the predicted maps scaled to each image
        b.predicted_maps_aug = [_scaled_map(p, np.asarray(it.x)) for p, it in zip(b.probabilities, items)]
        b.segmentation_maps = [SegmentationMapOnImage(np.asarray(it.y, dtype=np.float32)) for it in items]
        yield b


def evaluate(cfg, ds, fold: int, stage: int, negatives="all", limit=16, batchSize=32) -> Iterator[PredictionBatch]:
    """reference PipelineConfig.evaluate (segmentation.py:37-47): heat maps of (at most `limit`) VALIDATION items of one fold at
    network resolution -- yields batches with `.images_aug` (inputs as fed) and `.heatmaps_aug` (float probability maps)."""
    import cv2
    net = cfg.load_model(fold, stage)
    B = min(int(batchSize), net.batch)
    H, W = _net_hw(cfg)
    _, va = cfg.kfold(len(ds))[fold]
    idx = [int(i) for i in va]
    if negatives == "none" and hasattr(ds, "isPositive"):
        idx = [i for i in idx if ds.isPositive(i)]
    if limit is not None and limit > 0:
        idx = idx[:limit]
    for s in range(0, len(idx), B):
        items = [ds[i] for i in idx[s:s + B]]
        xs = [cv2.resize(np.asarray(it.x), (W, H), interpolation=cv2.INTER_CUBIC) if np.asarray(it.x).shape[:2] != (H, W)
              else np.asarray(it.x) for it in items]
        x = np.stack(xs).astype(np.uint8)
        probs = predict_arrays(net, x, False)
        b = PredictionBatch(items, [np.asarray(it.x) for it in items], list(probs), (H, W))
        b.images_aug = list(x)
        b.heatmaps_aug = [SegmentationMapOnImage(p) for p in probs]
        yield b
