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


class RawBatch:
    """One batch of UNRESIZED samples for on-device ingest (`device_resize: true`): the decoded images / masks packed back to
    back in pinned byte arenas plus one stp_resize_item per sample (include/stp.h).  Trainer.step_from_host_pipelined copies
    the arenas to the device and lets stp_resize_u8 (cv2 arithmetic: cubic for images, nearest for masks) write the
    network-shape batch -- the resize leaves the host's decode threads."""

    def __init__(self, capacity_img: int, capacity_mask: int, batch: int, pin: bool):
        import ctypes as C
        import torch
        from . import lib as _lib
        mk = lambda nbytes: (torch.zeros(nbytes, dtype=torch.uint8).pin_memory() if pin else torch.zeros(nbytes, dtype=torch.uint8))
        self.arena_img, self.arena_mask = mk(capacity_img), mk(capacity_mask)
        self.isz = C.sizeof(_lib.ResizeItem)
        self.items_img, self.items_mask = mk(batch * self.isz), mk(batch * self.isz)
        self.n = 0
        self.used_img = self.used_mask = 0
        self._pin = pin

    def pack(self, samples, c_img: int, c_mask: int):
        """samples: list of (x uint8 [h, w, c_img], y uint8 [h, w, c_mask])"""
        import torch
        from . import lib as _lib
        need_i = sum((x.size + 15) // 16 * 16 for x, _ in samples)
        need_m = sum((y.size + 15) // 16 * 16 for _, y in samples)
        if need_i > self.arena_img.numel() or need_m > self.arena_mask.numel():   # grow (rare: sized from the first batches)
            mk = lambda nbytes: (torch.zeros(nbytes, dtype=torch.uint8).pin_memory() if self._pin else torch.zeros(nbytes, dtype=torch.uint8))
            self.arena_img = mk(max(need_i * 3 // 2, self.arena_img.numel()))
            self.arena_mask = mk(max(need_m * 3 // 2, self.arena_mask.numel()))
        ai, am = self.arena_img.numpy(), self.arena_mask.numpy()
        ii = (_lib.ResizeItem * len(samples))()
        im = (_lib.ResizeItem * len(samples))()
        oi = om = 0
        for j, (x, y) in enumerate(samples):
            h, w = x.shape[:2]
            if y.shape[:2] != (h, w):
                raise ValueError("device_resize: image %s and mask %s sizes differ" % (x.shape, y.shape))
            ai[oi:oi + x.size] = x.reshape(-1)
            am[om:om + y.size] = y.reshape(-1)
            ii[j] = _lib.ResizeItem(oi, h, w, 0, 0, h, w)
            im[j] = _lib.ResizeItem(om, h, w, 0, 0, h, w)
            oi += (x.size + 15) // 16 * 16
            om += (y.size + 15) // 16 * 16
        self.items_img.numpy()[:len(samples) * self.isz] = np.frombuffer(bytes(ii), dtype=np.uint8)
        self.items_mask.numpy()[:len(samples) * self.isz] = np.frombuffer(bytes(im), dtype=np.uint8)
        self.n, self.used_img, self.used_mask = len(samples), oi, om


class HostLoader:
    RING = 4      # pinned batch slots
    AHEAD = 2     # batches being filled ahead of the one handed out

    def __init__(self, ds, shape, classes: int, batch: int, workers: int = 4, pin: bool = True, device_resize: bool = False):
        import torch
        self.ds, self.shape, self.B = ds, shape, int(batch)
        self.device_resize = bool(device_resize)
        self.classes = int(classes)
        if self.device_resize:
            H, W, C = int(shape[0]), int(shape[1]), int(shape[2])
            pin = pin and torch.cuda.is_available()
            # initial arena capacity: 4x the network-shape batch (sources larger than `shape` are the reason to resize at all)
            self.raw = [RawBatch(4 * self.B * H * W * C, 4 * self.B * H * W * self.classes, self.B, pin) for _ in range(self.RING)]
            self._raw_items = [[None] * self.B for _ in range(self.RING)]
            self.workers = int(workers)
            self.pool = ThreadPoolExecutor(self.workers) if self.workers > 0 else None
            return
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
        if self.device_resize:      # decode only; the batch is packed (and resized on the device) when it is handed out
            x, y = np.asarray(it.x), np.asarray(it.y)
            if y.ndim == 2:
                y = y[:, :, None]
            self._raw_items[slot][j] = (np.ascontiguousarray(x, dtype=np.uint8), np.ascontiguousarray(y, dtype=np.uint8))
            return
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
            if self.device_resize:
                slot = k % self.RING
                rb = self.raw[slot]
                rb.pack(self._raw_items[slot][:len(batches[k])], int(self.shape[2]), self.classes)
                yield rb, None
            else:
                yield self.img[k % self.RING], self.mask[k % self.RING]

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None
