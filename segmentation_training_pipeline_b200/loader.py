"""Host side of the data path: dataset items (decoded jpg/png of any size) -> resized to `shape`
(cv2 cubic for the image, nearest for the mask: imgaug's Resize defaults) -> batches in a ring of PINNED host buffers, filled
by a small thread pool a couple of batches ahead of the training loop, so that decode + resize overlap the device step and
the asynchronous H2D copy (Trainer.step_from_host_pipelined) always finds its batch ready.

Replaces the reference's imgaug BackgroundAugmenter worker processes + multiprocessing queue + np.array stacking
(musket_core.datasets.ImageKFoldedDataSet, selected at reference segmentation.py:54; queue depth FAQ.md:15-22) -- the
augmentation itself runs on the device (csrc/augment.cu), only decode / resize stay on the host."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Iterator, List, Sequence, Tuple

import numpy as np


def resize_pair(x: np.ndarray, y: np.ndarray, shape) -> Tuple[np.ndarray, np.ndarray]:
    import cv2
    H, W = int(shape[0]), int(shape[1])
    if x.shape[0] != H or x.shape[1] != W:
        x = cv2.resize(x, (W, H), interpolation=cv2.INTER_CUBIC)   # imgaug Resize default for images
        y = cv2.resize(y, (W, H), interpolation=cv2.INTER_NEAREST)
    if y.ndim == 2:
        y = y[:, :, None]
    return np.ascontiguousarray(x, dtype=np.uint8), np.ascontiguousarray(y, dtype=np.uint8)


class HostLoader:
    RING = 4      # pinned batch slots
    AHEAD = 2     # batches being filled ahead of the one handed out

    def __init__(self, ds, shape, classes: int, batch: int, workers: int = 4, pin: bool = True):
        import torch
        self.ds, self.shape, self.B = ds, shape, int(batch)
        H, W, C = int(shape[0]), int(shape[1]), int(shape[2])
        mk = lambda c: torch.zeros((self.B, H, W, c), dtype=torch.uint8)
        self.img = [mk(C) for _ in range(self.RING)]
        self.mask = [mk(int(classes)) for _ in range(self.RING)]
        if pin and torch.cuda.is_available():
            self.img = [t.pin_memory() for t in self.img]
            self.mask = [t.pin_memory() for t in self.mask]
        self._img_np = [t.numpy() for t in self.img]      # views of the (pinned) buffers the workers write into
        self._mask_np = [t.numpy() for t in self.mask]
        self.workers = int(workers)
        self.pool = ThreadPoolExecutor(self.workers) if self.workers > 0 else None

    def _fill_one(self, slot: int, j: int, index: int):
        it = self.ds[int(index)]
        x, y = resize_pair(np.asarray(it.x), np.asarray(it.y), self.shape)
        self._img_np[slot][j] = x
        self._mask_np[slot][j] = y

    def _submit(self, k: int, ids: Sequence[int]):
        slot = k % self.RING
        if self.pool is None:
            for j, i in enumerate(ids):
                self._fill_one(slot, j, i)
            return []
        return [self.pool.submit(self._fill_one, slot, j, i) for j, i in enumerate(ids)]

    def iterate(self, batches: List[Sequence[int]]) -> Iterator[Tuple["torch.Tensor", "torch.Tensor"]]:
        """Yields (images, masks) pinned uint8 tensors for every list of dataset indices (each of length <= batch; a shorter
        list leaves the tail rows of the slot as they were).  A yielded pair stays valid until TWO more pairs have been
        requested -- the contract of Trainer.step_from_host_pipelined, whose H2D copy of batch k may still be running while
        batch k+1 is handed out."""
        n = len(batches)
        futs = {}
        for k in range(min(self.AHEAD, n)):
            futs[k] = self._submit(k, batches[k])
        for k in range(n):
            for f in futs.pop(k):
                f.result()                      # re-raises a worker's exception here
            if k + self.AHEAD < n:              # slot (k+2)%4 last held batch k-2, whose successor's step call has returned
                futs[k + self.AHEAD] = self._submit(k + self.AHEAD, batches[k + self.AHEAD])
            yield self.img[k % self.RING], self.mask[k % self.RING]

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None
