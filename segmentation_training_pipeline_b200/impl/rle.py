"""Run-length encoding of binary masks in the Kaggle column-major convention the reference uses for submissions
(reference segmentation_pipeline/impl/rle.py:4-54: rle_encode / rle_decode / multi_rle_encode / masks_as_image).
Starts are 1-based, pixels are numbered top-to-bottom then left-to-right."""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np


def rle_encode(img) -> str:
    """binary H x W (or H x W x 1) array -> 'start length start length ...'"""
    a = np.asarray(img)
    if a.ndim == 3:
        a = a[:, :, 0]
    flat = np.concatenate([[0], (a.T.reshape(-1) != 0).astype(np.int8), [0]])
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    starts, ends = change[0::2], change[1::2]
    return " ".join("%d %d" % (s, e - s) for s, e in zip(starts, ends))


def rle_decode(mask_rle: str, shape: Sequence[int]) -> np.ndarray:
    """inverse of rle_encode for SQUARE masks; shape = (height, width).  Bit-for-bit the reference (impl/rle.py:22-35):
    `img.reshape(shape).T`, i.e. for a non-square shape the result is shape[1] x shape[0] with the runs laid out row-major in
    a shape[0] x shape[1] grid before the transpose -- the reference's quirk is kept, not corrected (its Kaggle use,
    768 x 768 Airbus masks, is square).  rle_decode_hw() below is the geometrically correct inverse for any shape."""
    h, w = int(shape[0]), int(shape[1])
    out = np.zeros(h * w, dtype=np.uint8)
    tok = mask_rle.split() if isinstance(mask_rle, str) else []
    for s, n in zip(tok[0::2], tok[1::2]):
        s = int(s) - 1
        out[s:s + int(n)] = 1
    return out.reshape(h, w).T


def rle_decode_hw(mask_rle: str, shape: Sequence[int]) -> np.ndarray:
    """true inverse of rle_encode for any (height, width): returns uint8 {0,1} of exactly that shape."""
    h, w = int(shape[0]), int(shape[1])
    out = np.zeros(h * w, dtype=np.uint8)
    tok = mask_rle.split() if isinstance(mask_rle, str) else []
    for s, n in zip(tok[0::2], tok[1::2]):
        s = int(s) - 1
        out[s:s + int(n)] = 1
    return out.reshape(w, h).T


def multi_rle_encode(img) -> List[str]:
    """one RLE string per 8-connected component of channel 0 (the reference labels with skimage.morphology.label)"""
    import cv2
    a = np.asarray(img)
    if a.ndim == 3:
        a = a[:, :, 0]
    n, labels = cv2.connectedComponents((a != 0).astype(np.uint8), connectivity=8)
    return [rle_encode(labels == k) for k in range(1, n)]


def masks_as_image(in_mask_list: Iterable, shape: Sequence[int]) -> np.ndarray:
    """sum of the decoded masks as H x W x 1 int16"""
    total = np.zeros((int(shape[0]), int(shape[1])), dtype=np.int16)
    for m in in_mask_list:
        if isinstance(m, str):
            total += rle_decode(m, shape)
    return total[:, :, None]


def masks_as_images(in_mask_list: Iterable, shape: Sequence[int]) -> List[np.ndarray]:
    return [rle_decode(m, shape).astype(np.float32) for m in in_mask_list if isinstance(m, str)]
