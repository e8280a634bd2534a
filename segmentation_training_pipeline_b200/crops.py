"""`crops: N` (reference README.md:471-491, schema key `crops`): every image / mask is split into N x N cells, the model is
trained on the cells (augmentation runs on each cell separately) and at prediction time each cell is predicted on its own and
the results are assembled back into one mask -- "the whole process of cropping is invisible from a consumer perspective".
The split itself lives in musket_core [DEP, unpinned]; cells here are the N x N grid with boundaries round(k * size / N)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .impl.datasets import PredictionItem


def cell_bounds(h: int, w: int, n: int) -> List[Tuple[int, int, int, int]]:
    """(y0, y1, x0, x1) of the n*n cells, row major."""
    ys = np.rint(np.linspace(0, h, n + 1)).astype(int)
    xs = np.rint(np.linspace(0, w, n + 1)).astype(int)
    return [(int(ys[r]), int(ys[r + 1]), int(xs[c]), int(xs[c + 1])) for r in range(n) for c in range(n)]


class CellDataSet:
    """Dataset protocol view whose item k*n*n + c is cell c of item k of the wrapped dataset."""

    def __init__(self, ds, n: int):
        self.ds, self.n, self.n2 = ds, int(n), int(n) * int(n)
        self._last = (None, None)

    def __len__(self):
        return len(self.ds) * self.n2

    def _item(self, k):
        # consecutive cells of one image: decode it once.  `_last` is only a hint shared by the loader's worker threads: the
        # pair is read ONCE into a local and the item returned is the one bound here, never a re-read of the shared slot
        last = self._last
        if last[0] == k:
            return last[1]
        it = self.ds[k]
        self._last = (k, it)
        return it

    def __getitem__(self, i) -> PredictionItem:
        i = int(i)
        k, c = divmod(i, self.n2)
        it = self._item(k)
        x, y = np.asarray(it.x), np.asarray(it.y)
        y0, y1, x0, x1 = cell_bounds(x.shape[0], x.shape[1], self.n)[c]
        return PredictionItem("%s_%d" % (it.id, c), x[y0:y1, x0:x1], y[y0:y1, x0:x1])

    def isPositive(self, i) -> bool:
        return bool(np.asarray(self[i].y).any())

    def expand(self, image_indices) -> np.ndarray:
        """cell indices of the given image indices (the cells of an image stay on the same side of a fold split)."""
        idx = np.asarray(image_indices, dtype=np.int64)
        return (idx[:, None] * self.n2 + np.arange(self.n2)[None, :]).reshape(-1)


def predict_image_by_cells(predict_fn, image: np.ndarray, n: int, shape_hw: Tuple[int, int], batch: int) -> np.ndarray:
    """Probability map [h, w, classes] of one image at ITS OWN size: cells are resized to the network shape (cubic, like the
    training path), predicted `batch` at a time by predict_fn(uint8 [k, H, W, 3]) -> [k, H, W, classes], resized back to the
    cell size (bilinear) and written into place."""
    import cv2
    H, W = shape_hw
    h, w = image.shape[:2]
    cells = cell_bounds(h, w, n)
    xs = []
    for (y0, y1, x0, x1) in cells:
        c = image[y0:y1, x0:x1]
        xs.append(cv2.resize(c, (W, H), interpolation=cv2.INTER_CUBIC) if c.shape[:2] != (H, W) else c)
    probs = []
    for s in range(0, len(xs), batch):
        probs.extend(list(predict_fn(np.stack(xs[s:s + batch]).astype(np.uint8))))
    out = None
    for (y0, y1, x0, x1), p in zip(cells, probs):
        if out is None:
            out = np.zeros((h, w, p.shape[-1]), np.float32)
        q = cv2.resize(p, (x1 - x0, y1 - y0), interpolation=cv2.INTER_LINEAR) if p.shape[:2] != (y1 - y0, x1 - x0) else p
        out[y0:y1, x0:x1] = q if q.ndim == 3 else q[:, :, None]
    return out
