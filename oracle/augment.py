"""imgaug==0.3.0 augmentation stage restated on numpy + cv2 [DEP]; reference selects augmenters through
schemas/augmenters.raml:43-133 and the YAML `augmentation:` key (README.md:249-268).

TEST INFRASTRUCTURE.  cv2.warpAffine (the backend imgaug's Affine calls) IS available here, so the
pixel/index arithmetic is PINNED bit-exactly by cv2 4.13 itself: `warp_cv2` is the oracle, and
`warp_fixedpoint` is the numpy restatement of cv2's fixed-point rule that the
CUDA kernel implements -- tests check warp_fixedpoint == warp_cv2 == CUDA.
Parameter DRAWS are not pinned (imgaug's RNG is not reproducible); they follow oracle/philox.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import philox

try:  # cv2 is the real backend; present in this image
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


@dataclass
class AugSpec:
    """The subset of augmenters.raml the device kernel fuses (Fliplr, Flipud, Affine, Multiply, Add)."""
    fliplr: float = 0.0
    flipud: float = 0.0
    affine: bool = False
    scale: Tuple[float, float] = (1.0, 1.0)
    translate_x: Tuple[float, float] = (0.0, 0.0)  # translate_percent
    translate_y: Tuple[float, float] = (0.0, 0.0)
    rotate: Tuple[float, float] = (0.0, 0.0)  # degrees
    shear: Tuple[float, float] = (0.0, 0.0)  # degrees
    multiply: Optional[Tuple[float, float]] = None
    add: Optional[Tuple[int, int]] = None
    rot90: bool = False            # musket Rotate90 (reference ds_1.yaml:6) [DEP, unpinned]: np.rot90 by uniform k, applied first
    invert: float = 0.0            # imgaug Invert(p)
    color_order: Tuple[int, int, int] = (0, 1, 2)   # 0 Multiply, 1 Add, 2 Invert in Sequential (YAML) order
    flip_before_rot90: int = 0     # bit 0: Fliplr is listed before Rotate90 (reference examples/people/ds_1.yaml:3-6), bit 1: Flipud


@dataclass
class SampleParams:
    fliplr: bool
    flipud: bool
    matrix: np.ndarray  # forward 2x3 float64 (src->dst), identity if no affine
    has_affine: bool
    mul: float  # float32-representable multiplier (1.0 = off)
    has_mul: bool
    add: int
    rot90_k: int = 0
    invert: bool = False
    color_order: Tuple[int, int, int] = (0, 1, 2)
    flip_before_rot90: int = 0


def affine_matrix(scale, tx_px, ty_px, rot_deg, shear_deg, h, w) -> np.ndarray:
    """imgaug 0.3.0 Affine matrix = to_center * AffineTransform(scale, rot, shear, translation) * to_topleft
    (skimage.transform semantics), centre = (w/2-0.5, h/2-0.5).  Scalar fp64 ops, no fused multiply-add."""
    rot = rot_deg * (math.pi / 180.0)
    sh = shear_deg * (math.pi / 180.0)
    a00 = scale * math.cos(rot)
    a01 = -(scale * math.sin(rot + sh))
    a10 = scale * math.sin(rot)
    a11 = scale * math.cos(rot + sh)
    cx, cy = w / 2.0 - 0.5, h / 2.0 - 0.5
    # A @ T(-c): third column
    b0 = (a00 * -cx + a01 * -cy) + tx_px
    b1 = (a10 * -cx + a11 * -cy) + ty_px
    # T(+c) @ ...
    return np.array([[a00, a01, b0 + cx], [a10, a11, b1 + cy]], dtype=np.float64)


def _lerp(lo, hi, u):
    return lo + u * (hi - lo)


def draw_params(spec: AugSpec, seed: int, step: int, sample: int, h: int, w: int) -> SampleParams:
    u_lr, u_ud = philox.uniforms(seed, step, sample, 0)
    u_sc, u_rot = philox.uniforms(seed, step, sample, 1)
    u_sh, u_tx = philox.uniforms(seed, step, sample, 2)
    u_ty, u_mul = philox.uniforms(seed, step, sample, 3)
    u_add, u_r90 = philox.uniforms(seed, step, sample, 4)
    u_inv, _ = philox.uniforms(seed, step, sample, 5)
    M = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    if spec.affine:
        scale = _lerp(spec.scale[0], spec.scale[1], u_sc)
        rot = _lerp(spec.rotate[0], spec.rotate[1], u_rot)
        sh = _lerp(spec.shear[0], spec.shear[1], u_sh)
        tx = _lerp(spec.translate_x[0], spec.translate_x[1], u_tx)
        ty = _lerp(spec.translate_y[0], spec.translate_y[1], u_ty)
        # imgaug 0.3.0: translate_percent -> px by int(np.round(p*dim)) (round half to even)
        tx_px = float(int(np.round(tx * w)))
        ty_px = float(int(np.round(ty * h)))
        M = affine_matrix(scale, tx_px, ty_px, rot, sh, h, w)
    mul = np.float32(1.0)
    if spec.multiply is not None:
        mul = np.float32(_lerp(spec.multiply[0], spec.multiply[1], u_mul))
    add = 0
    if spec.add is not None:
        lo, hi = int(spec.add[0]), int(spec.add[1])
        add = lo + int(math.floor(u_add * (hi - lo + 1)))
    k90 = min(int(math.floor(u_r90 * 4.0)), 3) if spec.rot90 else 0
    return SampleParams(u_lr < spec.fliplr, u_ud < spec.flipud, M, spec.affine, float(mul),
                        spec.multiply is not None, add, k90, bool(u_inv < spec.invert), tuple(spec.color_order),
                        int(spec.flip_before_rot90))


# ----------------------------------------------------------------------------------------------
# pixel arithmetic
# ----------------------------------------------------------------------------------------------
def invert_affine(M: np.ndarray):
    """cv2.warpAffine's in-place inversion of the forward matrix (imgwarp.cpp), fp64, unfused."""
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    m0, m1, m3, m4 = A11, M[0, 1] * -D, M[1, 0] * -D, A22
    b1 = -m0 * M[0, 2] - m1 * M[1, 2]
    b2 = -m3 * M[0, 2] - m4 * M[1, 2]
    return m0, m1, b1, m3, m4, b2


def warp_fixedpoint(src: np.ndarray, M: np.ndarray, nearest: bool, cval: int = 0) -> np.ndarray:
    """numpy restatement of cv2.warpAffine(BORDER_CONSTANT) fixed-point sampling."""
    H, W = src.shape[:2]
    s = src.reshape(H, W, -1)
    m0, m1, b1, m3, m4, b2 = invert_affine(M)
    xs = np.arange(W, dtype=np.float64)
    ys = np.arange(H, dtype=np.float64)
    adelta = np.rint(m0 * xs * 1024).astype(np.int64)
    bdelta = np.rint(m3 * xs * 1024).astype(np.int64)
    rd = 512 if nearest else 16
    X0 = np.rint((m1 * ys + b1) * 1024).astype(np.int64) + rd
    Y0 = np.rint((m4 * ys + b2) * 1024).astype(np.int64) + rd
    Xf = X0[:, None] + adelta[None, :]
    Yf = Y0[:, None] + bdelta[None, :]

    def tap(sy, sx):
        ok = (sx >= 0) & (sx < W) & (sy >= 0) & (sy < H)
        v = s[np.clip(sy, 0, H - 1), np.clip(sx, 0, W - 1)].astype(np.int64)
        return np.where(ok[..., None], v, cval)

    if nearest:
        out = tap(Yf >> 10, Xf >> 10)
    else:
        X, Y = Xf >> 5, Yf >> 5
        sx, ax, sy, ay = X >> 5, (X & 31)[..., None], Y >> 5, (Y & 31)[..., None]
        w00, w01, w10, w11 = (32 - ay) * (32 - ax) * 32, (32 - ay) * ax * 32, ay * (32 - ax) * 32, ay * ax * 32
        acc = w00 * tap(sy, sx) + w01 * tap(sy, sx + 1) + w10 * tap(sy + 1, sx) + w11 * tap(sy + 1, sx + 1)
        out = (acc + 16384) >> 15
    return out.astype(src.dtype).reshape(src.shape)


def warp_cv2(src: np.ndarray, M: np.ndarray, nearest: bool) -> np.ndarray:
    H, W = src.shape[:2]
    flag = cv2.INTER_NEAREST if nearest else cv2.INTER_LINEAR
    out = cv2.warpAffine(src, np.asarray(M, dtype=np.float64), dsize=(W, H), flags=flag,
                         borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    return out.reshape(src.shape)


def multiply_lut(mul: float, rounding: str = "trunc") -> np.ndarray:
    """imgaug 0.3.0 Multiply on uint8: LUT = clip(arange(256,f32)*mul,0,255).astype(uint8) (truncation).
    rounding='rint' is the alternative reading; flag, [DEP] unpinned."""
    v = np.arange(256, dtype=np.float32) * np.float32(mul)
    v = np.clip(v, 0, 255)
    if rounding == "rint":
        v = np.rint(v)
    return v.astype(np.uint8)


def apply(image: np.ndarray, mask: np.ndarray, p: SampleParams, use_cv2: bool = True, mul_rounding="trunc"):
    """Sequential([Fliplr, Flipud, Affine, Multiply, Add]) on a uint8 HxWx3 image and HxWx1 mask."""
    img, msk = image, mask
    # literal Sequential order: the flips listed before Rotate90, the quarter turns, the remaining flips
    lr_pre, ud_pre = bool(p.flip_before_rot90 & 1), bool(p.flip_before_rot90 & 2)
    if p.fliplr and lr_pre:
        img, msk = img[:, ::-1], msk[:, ::-1]
    if p.flipud and ud_pre:
        img, msk = img[::-1], msk[::-1]
    if p.rot90_k:
        img, msk = np.rot90(img, p.rot90_k), np.rot90(msk, p.rot90_k)
    if p.fliplr and not lr_pre:
        img, msk = img[:, ::-1], msk[:, ::-1]
    if p.flipud and not ud_pre:
        img, msk = img[::-1], msk[::-1]
    img, msk = np.ascontiguousarray(img), np.ascontiguousarray(msk)
    if p.has_affine:
        warp = warp_cv2 if (use_cv2 and cv2 is not None) else warp_fixedpoint
        img = warp(img, p.matrix, False)
        msk = warp(msk, p.matrix, True)
    for op in p.color_order:
        if op == 0 and p.has_mul:
            img = multiply_lut(p.mul, mul_rounding)[img]
        elif op == 1 and p.add != 0:
            img = np.clip(img.astype(np.int32) + p.add, 0, 255).astype(np.uint8)
        elif op == 2 and p.invert:
            img = (255 - img.astype(np.int32)).astype(np.uint8)
    return img, msk


def augment_batch(images: np.ndarray, masks: np.ndarray, spec: AugSpec, seed: int, step: int,
                  sample_ids=None, use_cv2=True):
    N, H, W = images.shape[:3]
    oi, om = np.empty_like(images), np.empty_like(masks)
    for n in range(N):
        sid = n if sample_ids is None else int(sample_ids[n])
        p = draw_params(spec, seed, step, sid, H, W)
        oi[n], om[n] = apply(images[n], masks[n], p, use_cv2)
    return oi, om


# ----------------------------------------------------------------------------------------------
# crop / pad family (schemas/augmenters.raml:72-87, 113-116 -> imgaug 0.3.0 Pad / PadToFixedSize / CropToFixedSize /
# CropAndPad [DEP, recalled -- parity unpinned]): window composition, then the pipeline's final Resize -> shape with cv2
# arithmetic (oracle/resize.py, pinned against the real cv2).  Choices taken where imgaug's source could not be consulted:
# PadToFixedSize puts floor((1 - u) * total) pixels before the image, CropToFixedSize removes floor(u * total) before it
# (position="uniform"); CropAndPad rounds percent * size half-to-even; Pad / CropAndPad keep_size=True.
# ----------------------------------------------------------------------------------------------
def crop_pad_window(ops, seed: int, step: int, sample: int, h: int, w: int):
    """ops: sequence of (kind, ranged, a, b, c, d) as in include/stp.h stp_croppad_op -> (vy0, vx0, vh, vw)"""
    u0, u1 = philox.uniforms(seed, step, sample, 6)
    u2, u3 = philox.uniforms(seed, step, sample, 7)
    u = (u0, u1, u2, u3)
    vy0, vx0, vh, vw = 0, 0, h, w
    for kind, ranged, a, b, c, d in ops:
        if kind == 1:
            vy0 -= int(a); vh += int(a) + int(c)
            vx0 -= int(d); vw += int(d) + int(b)
        elif kind == 2:
            if vw < int(a):
                tot = int(a) - vw
                vx0 -= int(math.floor((1.0 - u[0]) * tot)); vw = int(a)
            if vh < int(b):
                tot = int(b) - vh
                vy0 -= int(math.floor((1.0 - u[1]) * tot)); vh = int(b)
        elif kind == 3:
            if vw > int(a):
                tot = vw - int(a)
                vx0 += int(math.floor(u[2] * tot)); vw = int(a)
            if vh > int(b):
                tot = vh - int(b)
                vy0 += int(math.floor(u[3] * tot)); vh = int(b)
        elif kind == 4:
            f32 = lambda v: float(np.float32(v))     # the C struct carries the percentages as float32
            if ranged:
                lo, hi = f32(a), f32(b)
                pt, pr, pb, pl = (lo + u[k] * (hi - lo) for k in range(4))
            else:
                pt, pr, pb, pl = f32(a), f32(b), f32(c), f32(d)
            t, bb = int(np.rint(pt * vh)), int(np.rint(pb * vh))
            l, r = int(np.rint(pl * vw)), int(np.rint(pr * vw))
            nh, nw = vh + t + bb, vw + l + r
            if nh >= 1 and nw >= 1:
                vy0 -= t; vx0 -= l; vh, vw = nh, nw
        else:
            raise ValueError("unknown crop / pad kind %r" % (kind,))
    return vy0, vx0, vh, vw


def apply_crop_pad(image: np.ndarray, mask: np.ndarray, win, H: int, W: int):
    """window of (image, mask), zeros outside, resized to (H, W): cubic for the image, nearest for the mask"""
    from . import resize as R
    vi, vm = R.window(image, *win), R.window(mask, *win)
    return R.resize_cubic_u8(vi, H, W), R.resize_nearest_u8(vm, H, W)


# ----------------------------------------------------------------------------------------------
# pixel-wise augmenters (schemas/augmenters.raml:43-60, 88-96, 120-122 -> imgaug 0.3.0 [DEP, recalled -- parity unpinned]):
# the CPU twin of csrc/augment.cu augment_pixel_ops_kernel: same Philox counters, same fp32 operation order.
# ops: (kind, per_channel, a, b, group_id, group_size, group_member); kinds per include/stp.h STP_PIX_*.
# ----------------------------------------------------------------------------------------------
def apply_pixel_ops(img: np.ndarray, ops, seed: int, step: int, sid: int, p: SampleParams, mul_rint: bool = False,
                    k_base: int = 0) -> np.ndarray:
    f32 = np.float32
    H, W, CI = img.shape
    v = img.astype(np.int64).reshape(-1, CI)
    pix = np.arange(H * W, dtype=np.uint64)
    key = (seed & philox.MASK, (seed >> 32) & philox.MASK)
    s_lo, s_hi = step & philox.MASK, (step >> 32) & philox.MASK
    u24 = lambda w: ((w >> np.uint64(8)).astype(np.float32) * f32(5.9604644775390625e-08)).astype(np.float32)
    for k, (kind, per_channel, a, b, gid, gsz, gm) in enumerate(ops, start=k_base):   # k = position in the whole colour block
        ri = philox.philox4x32((s_lo, sid, 32 + k, s_hi), key)
        if gsz > 0:
            rg = philox.philox4x32((s_lo, sid, 32 + 16 + gid, s_hi), key)
            pick = min(int(math.floor(philox.u53(rg[0], rg[1]) * gsz)), gsz - 1)
            if pick != gm:
                continue
        pc = philox.u53(ri[0], ri[1]) < float(f32(per_channel))
        a32, b32 = f32(a), f32(b)
        par = f32(a32 + f32(f32(philox.u53(ri[2], ri[3])) * f32(b32 - a32)))
        if kind == 0:
            if p.has_mul:
                fv = np.clip((v.astype(np.float32) * f32(p.mul)).astype(np.float32), 0, 255)
                v = (np.rint(fv) if mul_rint else np.trunc(fv)).astype(np.int64)
        elif kind == 1:
            v = np.clip(v + int(p.add), 0, 255)
        elif kind == 2:
            if p.invert:
                v = 255 - v
        elif kind == 7:
            if CI >= 3:
                vf = v.astype(np.float32)
                g = (f32(0.299) * vf[:, 0] + f32(0.587) * vf[:, 1]).astype(np.float32) + f32(0.114) * vf[:, 2]
                for c in range(3):
                    fv = (f32(1.0) - par) * vf[:, c] + par * g
                    v[:, c] = np.clip(np.rint(fv.astype(np.float32)), 0, 255).astype(np.int64)
        else:
            rp = philox.philox4x32_vec(s_lo, sid, (pix << np.uint64(8)) | np.uint64(64 + k), s_hi, *key)
            rq = philox.philox4x32_vec(s_lo, sid, (pix << np.uint64(8)) | np.uint64(64 + 128 + k), s_hi, *key) if kind == 6 else None
            for c in range(CI):
                wi = (c & 3) if pc else 0
                u = u24(rp[wi])
                if kind == 3:
                    lo, hi = int(a32), int(b32)
                    v[:, c] = np.clip(v[:, c] + lo + np.floor((u * f32(hi - lo + 1)).astype(np.float32)).astype(np.int64), 0, 255)
                elif kind == 4:
                    m = (a32 + (u * f32(b32 - a32)).astype(np.float32)).astype(np.float32)
                    fv = np.clip((v[:, c].astype(np.float32) * m).astype(np.float32), 0, 255)
                    v[:, c] = (np.rint(fv) if mul_rint else np.trunc(fv)).astype(np.int64)
                elif kind == 5:
                    v[:, c] = np.where(u < par, 0, v[:, c])
                elif kind == 6:
                    u2 = u24(rq[wi])
                    z = (np.sqrt((f32(-2.0) * np.log((f32(1.0) - u).astype(np.float32)).astype(np.float32)).astype(np.float32)).astype(np.float32)
                         * np.cos((f32(6.283185307179586) * u2).astype(np.float32)).astype(np.float32)).astype(np.float32)
                    fv = (v[:, c].astype(np.float32) + (z * par).astype(np.float32)).astype(np.float32)
                    v[:, c] = np.clip(np.rint(fv), 0, 255).astype(np.int64)
    return v.reshape(H, W, CI).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# Neighbourhood augmenters (schemas/augmenters.raml:97-112, 117-119 -> imgaug 0.3.0 GaussianBlur / AverageBlur / MedianBlur /
# Sharpen / Emboss / EdgeDetect [DEP, recalled] -> cv2.GaussianBlur / cv2.blur / cv2.medianBlur / cv2.filter2D): the CPU twin of
# csrc/augment_nb.cu.  The cv2 ARITHMETIC restated below is pinned against the real cv2 of this image (tests/test_cpu_oracle.py
# ::test_neighbourhood_arithmetic_is_pinned_against_real_cv2); the imgaug parameter conventions are recalled.
# op: (kind, a, b, c, d, k_index, group_id, group_size, group_member); kinds per include/stp.h STP_NB_*.
# ----------------------------------------------------------------------------------------------
def gaussian_ksize_imgaug(sigma: float) -> int:
    k = 3.3 * sigma if sigma < 3.0 else (2.9 * sigma if sigma < 5.0 else 2.6 * sigma)
    k = int(max(k, 5))
    return k + 1 if k % 2 == 0 else k


def gaussian_kernel_fixed(ksize: int, sigma: float) -> np.ndarray:
    """cv2's 8.8 fixed-point Gaussian kernel for uint8 images: round(k_i * 256) with error diffusion from the border inwards,
    the centre takes the remainder (sum == 256)."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    kd = np.exp((-0.5 / (sigma * sigma)) * x * x)
    kd = kd * (1.0 / kd.sum())
    out = np.zeros(ksize, np.int64)
    err, s2 = 0.0, 0
    for i in range(ksize // 2):
        adj = kd[i] * 256.0 + err
        v = int(math.floor(adj + 0.5))
        err = adj - v
        out[i] = out[ksize - 1 - i] = v
        s2 += 2 * v
    out[ksize // 2] = 256 - s2
    return out


def _window_stack(img2d: np.ndarray, kh: int, kw: int, ay: int, ax: int, mode: str):
    p = np.pad(img2d, ((ay, kh - 1 - ay), (ax, kw - 1 - ax)), mode=mode)
    h, w = img2d.shape
    return [[p[a:a + h, b:b + w] for b in range(kw)] for a in range(kh)]


def gaussian_blur_u8(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    k = gaussian_kernel_fixed(ksize, sigma)
    out = np.empty_like(img)
    for c in range(img.shape[2]):
        win = _window_stack(img[..., c].astype(np.int64), ksize, ksize, ksize // 2, ksize // 2, "reflect")
        acc = sum(k[a] * sum(k[b] * win[a][b] for b in range(ksize)) for a in range(ksize))
        out[..., c] = np.clip((acc + 32768) >> 16, 0, 255)
    return out


def average_blur_round_up_from(k: int) -> int:
    """smallest residue of the window sum modulo k^2 that cv2.blur rounds UP (measured on OpenCV 4.13 for k = 2..17, every pixel
    of random images consistent): round half up, except for k a power of two where the shift path rounds up one residue earlier
    (out = (sum + k^2/2 + 1) >> log2(k^2))"""
    kk = k * k
    return kk // 2 - 1 if (k & (k - 1)) == 0 else (kk + 1) // 2


def average_blur_u8(img: np.ndarray, k: int) -> np.ndarray:
    out = np.empty_like(img)
    for c in range(img.shape[2]):
        win = _window_stack(img[..., c].astype(np.int64), k, k, k // 2, k // 2, "reflect")
        acc = sum(win[a][b] for a in range(k) for b in range(k))
        out[..., c] = acc // (k * k) + (acc % (k * k) >= average_blur_round_up_from(k))
    return out


def median_blur_u8(img: np.ndarray, k: int) -> np.ndarray:
    out = np.empty_like(img)
    for c in range(img.shape[2]):
        win = _window_stack(img[..., c], k, k, k // 2, k // 2, "edge")
        st = np.stack([win[a][b] for a in range(k) for b in range(k)], 0)
        out[..., c] = np.sort(st, axis=0)[(k * k) // 2]
    return out


def filter2d_3x3_u8(img: np.ndarray, mat: np.ndarray) -> np.ndarray:
    """cv2.filter2D(uint8 image, -1, float32 3x3 kernel): acc = fma(k, x, acc) in float32, row-major over the non-zero taps (the
    AVX2 / FMA3 build of OpenCV 4.13 fuses the multiply-add; separate multiply and add differs by 1 on ~0.1 % of the pixels).
    The fused operation is emulated in float64: a float32 product is exact there."""
    f32 = np.float32
    out = np.empty_like(img)
    for c in range(img.shape[2]):
        win = _window_stack(img[..., c].astype(np.float32), 3, 3, 1, 1, "reflect")
        acc = np.zeros(img.shape[:2], np.float32)
        for a in range(3):
            for b in range(3):
                if f32(mat[a, b]) != 0:
                    acc = (acc.astype(np.float64) + np.float64(f32(mat[a, b])) * win[a][b].astype(np.float64)).astype(np.float32)
        out[..., c] = np.clip(np.rint(acc), 0, 255)
    return out


def directed_edge_table() -> np.ndarray:
    """imgaug 0.3.0 DirectedEdgeDetect [DEP, recalled]: effect matrix per integer degree (deg = int(direction * 360) % 360):
    the direction vector is (cos, sin)(rad - pi/2); every non-centre cell (x, y) of the 3x3 matrix gets (1 - angle/180deg)^4 with
    the angle between the cell vector and the direction vector; normalised to sum 1, negated, centre 1.  float32 [360][3][3]."""
    tab = np.zeros((360, 3, 3), np.float32)
    for deg in range(360):
        rad = np.deg2rad(deg)
        dv = np.array([np.cos(rad - 0.5 * np.pi), np.sin(rad - 0.5 * np.pi)])
        m = np.zeros((3, 3), np.float32)
        for x in (-1, 0, 1):
            for y in (-1, 0, 1):
                if (x, y) != (0, 0):
                    cv = np.array([x, y], np.float64)
                    cos_a = np.clip(np.dot(cv / np.linalg.norm(cv), dv / np.linalg.norm(dv)), -1.0, 1.0)
                    distance = np.rad2deg(np.arccos(cos_a)) / 180.0
                    m[y + 1, x + 1] = (1.0 - distance) ** 4
        m = m / np.sum(m)
        m = m * np.float32(-1)
        m[1, 1] = 1
        tab[deg] = m
    return tab


_DIRECTED_TABLE = None


def neighbourhood_params(op, seed: int, step: int, sid: int):
    """(active, kind-specific parameters) of one sample: the draws of csrc/augment_nb.cu nb_prep_kernel"""
    kind, a, b, c, d, k_index, gid, gsz, gm = op
    f32 = np.float32
    key = (seed & philox.MASK, (seed >> 32) & philox.MASK)
    s_lo, s_hi = step & philox.MASK, (step >> 32) & philox.MASK
    if gsz > 0:
        rg = philox.philox4x32((s_lo, sid, 32 + 16 + gid, s_hi), key)
        if min(int(math.floor(philox.u53(rg[0], rg[1]) * gsz)), gsz - 1) != gm:
            return False, None
    ri = philox.philox4x32((s_lo, sid, 32 + k_index, s_hi), key)
    u1, u2 = philox.u53(ri[0], ri[1]), philox.u53(ri[2], ri[3])
    a, b, c, d = float(f32(a)), float(f32(b)), float(f32(c)), float(f32(d))
    p1, p2 = a + u1 * (b - a), c + u2 * (d - c)
    if kind == 0:
        return (p1 >= 1e-3), ("gauss", min(gaussian_ksize_imgaug(p1), 25), p1)
    if kind in (1, 2):
        lo, hi = int(a), int(b)
        k = min(lo + int(math.floor(u1 * (hi - lo + 1))), hi)
        if kind == 2 and k % 2 == 0:
            k += 1
        return (k > 1), ("avg" if kind == 1 else "median", k)
    alpha, c1 = f32(p1), f32(1.0 - p1)
    if kind == 3:
        eff = np.array([[-1, -1, -1], [-1, f32(8.0 + p2), -1], [-1, -1, -1]], np.float32)
    elif kind == 4:
        eff = np.array([[f32(-1.0 - p2), f32(0.0 - p2), 0], [f32(0.0 - p2), 1, f32(0.0 + p2)], [0, f32(0.0 + p2), f32(1.0 + p2)]], np.float32)
    elif kind == 6:   # DirectedEdgeDetect: second parameter = direction in [0, 1] -> integer degree -> table row
        global _DIRECTED_TABLE
        if _DIRECTED_TABLE is None:
            _DIRECTED_TABLE = directed_edge_table()
        eff = _DIRECTED_TABLE[int(p2 * 360.0) % 360]
    else:
        eff = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
    ident = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], np.float32)
    mat = ((c1 * ident).astype(np.float32) + (alpha * eff).astype(np.float32)).astype(np.float32)
    return True, ("filter", mat)


def apply_neighbourhood_op(img: np.ndarray, op, seed: int, step: int, sid: int) -> np.ndarray:
    active, par = neighbourhood_params(op, seed, step, sid)
    if not active:
        return img.copy()
    if par[0] == "gauss":
        return gaussian_blur_u8(img, par[1], par[2])
    if par[0] == "avg":
        return average_blur_u8(img, par[1])
    if par[0] == "median":
        return median_blur_u8(img, par[1])
    return filter2d_3x3_u8(img, par[1])
