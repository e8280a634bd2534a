"""cv2.resize restated for uint8 images (TEST INFRASTRUCTURE; product code never imports oracle/).

The reference's input pipeline ends with imgaug `Resize -> shape` = cv2.resize (INTER_CUBIC for images, INTER_NEAREST for
segmentation maps) [DEP imgaug==0.3.0 -> opencv]; this module restates OpenCV's REFERENCE (scalar) arithmetic,
imgproc/src/resize.cpp (resizeGeneric_ / HResizeCubic / VResizeCubic with FixedPtCast<int, uchar, 22>), in numpy.

PINNED: tests/test_cpu_oracle.py compares it with the real cv2 4.13 of this container with IPP switched off
(cv2.ipp.setUseIPP(False)): identical except at rounding ties, where OpenCV's own SIMD vertical pass (fp32 multiply-add
chain, round-half-even) differs from this scalar integer path (round-half-up) -- < 0.05 % of pixels, |difference| = 1.
With Intel IPP on (the pip wheel's default) cv2 calls ippiResizeCubic instead, a different algorithm (+-1 on ~5 % of pixels).
"""
from __future__ import annotations

import numpy as np


def _cubic_coeffs(x: np.float32) -> np.ndarray:
    A = np.float32(-0.75)
    one = np.float32(1.0)
    x = np.float32(x)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def _cubic_axis(src: int, dst: int):
    scale = src / dst
    ofs = np.zeros(dst, np.int64)
    co = np.zeros((dst, 4), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        c = _cubic_coeffs(f)
        co[d] = np.clip(np.rint(c * np.float32(2048)), -32768, 32767).astype(np.int64)
        ofs[d] = s
    return ofs, co


def resize_cubic_u8(img: np.ndarray, H: int, W: int) -> np.ndarray:
    """cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC), scalar path; img uint8 [h, w, C]."""
    if img.ndim == 2:
        return resize_cubic_u8(img[:, :, None], H, W)[:, :, 0]
    h, w, C = img.shape
    if (h, w) == (H, W):
        return img.copy()
    xo, xa = _cubic_axis(w, W)
    yo, ya = _cubic_axis(h, H)
    src = img.astype(np.int64)
    tmp = np.zeros((h, W, C), np.int64)
    for k in range(4):
        tmp += src[:, np.clip(xo - 1 + k, 0, w - 1), :] * xa[:, k][None, :, None]
    out = np.zeros((H, W, C), np.int64)
    for k in range(4):
        out += tmp[np.clip(yo - 1 + k, 0, h - 1), :, :] * ya[:, k][:, None, None]
    return np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)


def resize_nearest_u8(img: np.ndarray, H: int, W: int) -> np.ndarray:
    """cv2.resize(img, (W, H), interpolation=cv2.INTER_NEAREST): sx = min(floor(dx * (1 / (W / w))), w - 1)."""
    if img.ndim == 2:
        return resize_nearest_u8(img[:, :, None], H, W)[:, :, 0]
    h, w = img.shape[:2]
    if (h, w) == (H, W):
        return img.copy()
    ys = np.minimum(np.floor(np.arange(H) * (1.0 / (H / h))).astype(np.int64), h - 1)
    xs = np.minimum(np.floor(np.arange(W) * (1.0 / (W / w))).astype(np.int64), w - 1)
    return img[ys][:, xs]


def window(img: np.ndarray, vy0: int, vx0: int, vh: int, vw: int) -> np.ndarray:
    """the virtual image of stp_resize_item: rows [vy0, vy0+vh) x columns [vx0, vx0+vw) of img, zeros outside (np.pad constant)"""
    if img.ndim == 2:
        return window(img[:, :, None], vy0, vx0, vh, vw)[:, :, 0]
    h, w, C = img.shape
    out = np.zeros((vh, vw, C), img.dtype)
    y0, y1, x0, x1 = max(vy0, 0), min(vy0 + vh, h), max(vx0, 0), min(vx0 + vw, w)
    if y1 > y0 and x1 > x0:
        out[y0 - vy0:y1 - vy0, x0 - vx0:x1 - vx0] = img[y0:y1, x0:x1]
    return out
