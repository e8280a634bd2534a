"""Philox4x32-10 counter-based RNG (Salmon et al. 2011), numpy restatement.  TEST INFRASTRUCTURE.

The reference's imgaug stream (SFC64, reseeded per worker process) is not reproducible run-to-run
(SURVEY.md Appendix B), so the engine defines its own draw: Philox(key=seed, counter=(step, sample,
call, 0)).  This file is the CPU twin of csrc/philox.cuh so oracle and engine draw identical values.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32(counter, key):
    c0, c1, c2, c3 = [int(c) & MASK for c in counter]
    k0, k1 = [int(k) & MASK for k in key]
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u53(hi: int, lo: int) -> float:
    """two uint32 -> uniform double in [0,1) with 53 random bits (exact in fp64)."""
    return float(((hi << 32) | lo) >> 11) * (2.0 ** -53)


def uniforms(seed: int, step: int, sample: int, call: int):
    x = philox4x32((step & MASK, sample & MASK, call & MASK, (step >> 32) & MASK), (seed & MASK, (seed >> 32) & MASK))
    return u53(x[0], x[1]), u53(x[2], x[3])


def philox4x32_vec(c0, c1, c2, c3, k0, k1):
    """vectorised Philox4x32-10: counters are uint64 numpy arrays (values < 2^32) or scalars; returns four uint64 arrays"""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & np.uint64(MASK) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint64(int(k0) & MASK), np.uint64(int(k1) & MASK)
    m = np.uint64(MASK)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m, p1 & m, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m, p0 & m
        k0, k1 = (k0 + np.uint64(W0)) & m, (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def dropout_keep_mask(rows: int, channels: int, rate: float, seed: int, salt: int, step: int) -> np.ndarray:
    """keep mask [rows, channels] (bool) of the engine's Dropout (csrc/dropout.cu): element (row, 8*v + j) keeps iff word j%4 of
    Philox(counter=(step lo, i lo, (i hi << 1) | (j >= 4), step hi), key=(seed lo, seed hi ^ salt)) >= rate * 2^32, i = row*(C/8)+v."""
    assert channels % 8 == 0
    cv = channels // 8
    i = np.arange(rows * cv, dtype=np.uint64)
    thresh = min(int(float(np.float32(rate)) * 4294967296.0), 0xFFFFFFFF)
    k0, k1 = seed & MASK, ((seed >> 32) & MASK) ^ (salt & MASK)
    words = []
    for half in (0, 1):
        w = philox4x32_vec(step & MASK, i & np.uint64(MASK), ((i >> np.uint64(32)) << np.uint64(1)) | np.uint64(half), (step >> 32) & MASK, k0, k1)
        words.extend(w)
    m = np.stack(words, axis=1) >= np.uint64(thresh)        # [rows*cv, 8]
    return m.reshape(rows, channels)
