"""ctypes binding of libstp.so (include/stp.h).  Fails LOUDLY when the CUDA library is missing: there is no
CPU fallback for the product path (the oracle under /oracle is test infrastructure and is never imported here).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstp.so")

BF16, F32, U8 = 0, 1, 2
OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_WORKSPACE = 0, -1, -2, -3, -4
CONV_RELU, CONV_STATS = 1, 2
L_LOSS, L_BCE, L_DICE, L_IOU, L_ACC, L_IOT, L_SUM_P, L_SUM_T, L_SUM_PT, L_COUNT, L_LOVASZ, L_JACCARD, L_FOCAL, L_CCE, L_CACC = range(15)
BN_MAX_PARTIALS = 1024


class StpError(RuntimeError):
    pass


class Tensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("ld", C.c_int32), ("dtype", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("r", C.c_int32), ("s", C.c_int32), ("stride", C.c_int32), ("pad_h", C.c_int32),
                ("pad_w", C.c_int32), ("up", C.c_int32), ("flags", C.c_int32)]


class DwConvDesc(C.Structure):              # include/stp.h: stp_dwconv_desc
    _fields_ = [("k", C.c_int32), ("stride", C.c_int32), ("dilation", C.c_int32), ("pad_h", C.c_int32), ("pad_w", C.c_int32)]


class AugSpec(C.Structure):
    _fields_ = [("fliplr_p", C.c_double), ("flipud_p", C.c_double), ("affine", C.c_int32),
                ("scale_lo", C.c_double), ("scale_hi", C.c_double),
                ("tx_lo", C.c_double), ("tx_hi", C.c_double), ("ty_lo", C.c_double), ("ty_hi", C.c_double),
                ("rot_lo", C.c_double), ("rot_hi", C.c_double), ("shear_lo", C.c_double), ("shear_hi", C.c_double),
                ("has_mul", C.c_int32), ("mul_lo", C.c_double), ("mul_hi", C.c_double),
                ("has_add", C.c_int32), ("add_lo", C.c_int32), ("add_hi", C.c_int32), ("mul_rint", C.c_int32),
                ("rot90", C.c_int32), ("invert_p", C.c_double), ("color_order", C.c_int32 * 3),
                ("flip_before_rot90", C.c_int32)]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        if len(a) < 23 and "color_order" not in kw:   # default colour order: Multiply, Add, Invert
            self.color_order[0], self.color_order[1], self.color_order[2] = 0, 1, 2


class ResizeItem(C.Structure):             # include/stp.h: stp_resize_item
    _fields_ = [("src_off", C.c_int64), ("sh", C.c_int32), ("sw", C.c_int32), ("vy0", C.c_int32), ("vx0", C.c_int32),
                ("vh", C.c_int32), ("vw", C.c_int32)]


RESIZE_NEAREST, RESIZE_CUBIC = 0, 1
CP_PAD, CP_PAD_TO_FIXED, CP_CROP_TO_FIXED, CP_CROP_AND_PAD = 1, 2, 3, 4


PIX_KINDS = {"Multiply": 0, "Add": 1, "Invert": 2, "AddElementwise": 3, "MultiplyElementwise": 4, "Dropout": 5,
             "AdditiveGaussianNoise": 6, "Grayscale": 7}


class AugPixOp(C.Structure):               # include/stp.h: stp_aug_pix_op
    _fields_ = [("kind", C.c_int32), ("per_channel", C.c_float), ("a", C.c_float), ("b", C.c_float), ("group_id", C.c_int32),
                ("group_size", C.c_int32), ("group_member", C.c_int32)]


class AugPixSpec(C.Structure):             # include/stp.h: stp_aug_pix_spec
    _fields_ = [("n_ops", C.c_int32), ("mul_rint", C.c_int32), ("ops", AugPixOp * 8), ("k_base", C.c_int32)]


class AugNbOp(C.Structure):                # include/stp.h: stp_aug_nb_op
    _fields_ = [("kind", C.c_int32), ("a", C.c_float), ("b", C.c_float), ("c", C.c_float), ("d", C.c_float), ("k_index", C.c_int32),
                ("group_id", C.c_int32), ("group_size", C.c_int32), ("group_member", C.c_int32), ("d_table", C.c_void_p)]


NB_KINDS = {"GaussianBlur": 0, "AverageBlur": 1, "MedianBlur": 2, "Sharpen": 3, "Emboss": 4, "EdgeDetect": 5, "DirectedEdgeDetect": 6}


def directed_edge_table():
    """float32 [360][3][3]: imgaug 0.3.0 DirectedEdgeDetect's effect matrix per integer degree [DEP, recalled] (deg = int(direction *
    360) % 360; direction vector (cos, sin)(rad - pi/2); cell (x, y) != centre gets (1 - angle / 180deg)^4, normalised to sum 1,
    negated, centre 1) -- tabulated on the host because imgaug evaluates it with double-precision numpy trigonometry."""
    import numpy as np
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
                    m[y + 1, x + 1] = (1.0 - np.rad2deg(np.arccos(cos_a)) / 180.0) ** 4
        m = m / np.sum(m)
        m = m * np.float32(-1)
        m[1, 1] = 1
        tab[deg] = m
    return tab


class CropPadOp(C.Structure):              # include/stp.h: stp_croppad_op
    _fields_ = [("kind", C.c_int32), ("ranged", C.c_int32), ("a", C.c_float), ("b", C.c_float), ("c", C.c_float), ("d", C.c_float)]


class CropPadSpec(C.Structure):            # include/stp.h: stp_croppad_spec
    _fields_ = [("n_ops", C.c_int32), ("ops", CropPadOp * 4)]


class AugSample(C.Structure):
    _fields_ = [("m", C.c_double * 6), ("inv", C.c_double * 6), ("fliplr", C.c_int32), ("flipud", C.c_int32),
                ("has_affine", C.c_int32), ("has_mul", C.c_int32), ("mul", C.c_float), ("add", C.c_int32),
                ("src_index", C.c_int32), ("flags2", C.c_int32)]


class BnFwd(C.Structure):
    _fields_ = [("partial", C.c_void_p), ("sync", C.c_void_p), ("acc", C.c_void_p), ("gamma", C.c_void_p),
                ("beta", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float), ("moving_mean", C.c_void_p),
                ("moving_var", C.c_void_p), ("coef", C.c_void_p)]


class BnBwd(C.Structure):
    _fields_ = [("x", C.POINTER(Tensor)), ("coef", C.c_void_p), ("relu", C.c_int32), ("partial", C.c_void_p),
                ("sync", C.c_void_p), ("acc", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("bcoef", C.c_void_p)]


class LossSpec(C.Structure):
    _fields_ = [("w_bce", C.c_float), ("w_dice", C.c_float), ("w_iou", C.c_float), ("w_jaccard", C.c_float),
                ("w_focal", C.c_float)]


class GradXform(C.Structure):
    _fields_ = [("scale", C.c_float), ("clipnorm", C.c_float), ("clipvalue", C.c_float), ("d_sumsq", C.c_void_p),
                ("d_lr_scale", C.c_void_p)]


_P, _I32, _I64, _U64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
_TP, _CDP, _DDP = C.POINTER(Tensor), C.POINTER(ConvDesc), C.POINTER(DwConvDesc)

# name -> (restype, argtypes); every symbol include/stp.h declares
SIGNATURES = {
    "stp_version": (C.c_int, []),
    "stp_last_error": (C.c_char_p, []),
    "stp_launch_count": (_I64, []),
    "stp_tc_enabled": (C.c_int, []),
    "stp_tc_launch_count": (_I64, []),
    "stp_tc3_launch_count": (_I64, []),
    "stp_set_tc_enabled": (None, [C.c_int]),
    "stp_set_option": (C.c_int, [C.c_char_p, _I32]),
    "stp_set_trace_buffer": (None, [_P]),
    "stp_augment_draw": (C.c_int, [C.POINTER(AugSpec), _U64, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "stp_augment_apply": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "stp_conv_fwd": (C.c_int, [_CDP, _TP, _P, _P, _TP, _TP, _P, _SZ, _P]),
    "stp_conv_fwd_bn": (C.c_int, [_CDP, _TP, _P, _P, _TP, _TP, C.POINTER(BnFwd), _P, _SZ, _P]),
    "stp_conv_dgrad": (C.c_int, [_CDP, _TP, _P, _TP, _TP, _P, _SZ, _P]),
    "stp_conv_dgrad_bn": (C.c_int, [_CDP, _TP, _P, _TP, C.POINTER(BnBwd), _P, _SZ, _P]),
    "stp_conv_wgrad": (C.c_int, [_CDP, _TP, _TP, _P, _P, _SZ, _P]),
    "stp_conv_wgrad_workspace": (_SZ, [_CDP, _TP, _TP]),
    "stp_weight_prep": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "stp_weight_prep_batched": (C.c_int, [_P, _P, _P, _P, _I32, _I64, _P]),
    "stp_head_fwd": (C.c_int, [_TP, _P, _P, _I32, _P, _P, _SZ, _P]),
    "stp_head_fwd_workspace": (_SZ, [_TP, _I32]),
    "stp_head_bwd": (C.c_int, [_TP, _P, _P, _I32, _TP, _P, _P, _P, _SZ, _P]),
    "stp_head_bwd_workspace": (_SZ, [_TP, _I32]),
    "stp_bn_nblk": (_I32, [_I64, _I32]),
    "stp_bn_stats": (C.c_int, [_TP, _P, _P]),
    "stp_bn_stats_fused": (C.c_int, [_TP, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P]),
    "stp_bn_bwd_reduce_fused": (C.c_int, [_TP, _TP, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "stp_bias_grad": (C.c_int, [_TP, _P, _P, _P, _P, _P]),
    "stp_bn_finalize": (C.c_int, [_P, _I32, _I32, _I64, _P, _P, _F, _F, _P, _P, _P, _P]),
    "stp_bn_apply": (C.c_int, [_TP, _P, _I32, _I32, _TP, _P]),
    "stp_bn_coef_infer": (C.c_int, [_P, _P, _P, _P, _F, _I32, _P, _P]),
    "stp_bn_bwd_reduce": (C.c_int, [_TP, _TP, _P, _I32, _I32, _P, _P]),
    "stp_bn_bwd_finalize": (C.c_int, [_P, _I32, _I32, _I64, _P, _P, _P, _P, _P]),
    "stp_bn_bwd_apply": (C.c_int, [_TP, _TP, _P, _P, _I32, _I32, _TP, _TP, _P]),
    "stp_relu_bwd": (C.c_int, [_TP, _TP, _I32, _TP, _TP, _P]),
    "stp_stem_prep": (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _TP, _P]),
    "stp_stem_weight_s2d": (C.c_int, [_P, _P, _I32, _P]),
    "stp_stem_wgrad_s2d_gather": (C.c_int, [_P, _P, _I32, _P]),
    "stp_stem_wgrad_post": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "stp_maxpool_fwd": (C.c_int, [_TP, _I32, _I32, _I32, _TP, _P, _P]),
    "stp_maxpool_bwd": (C.c_int, [_TP, _P, _I32, _I32, _I32, _TP, _TP, _P]),
    "stp_avgpool_fwd": (C.c_int, [_TP, _I32, _TP, _P]),
    "stp_avgpool_bwd": (C.c_int, [_TP, _I32, _TP, _TP, _P]),
    "stp_global_avgpool_fwd": (C.c_int, [_TP, _TP, _P]),
    "stp_global_avgpool_bwd": (C.c_int, [_TP, _TP, _TP, _P]),
    "stp_broadcast_fwd": (C.c_int, [_TP, _TP, _P]),
    "stp_broadcast_bwd": (C.c_int, [_TP, _TP, _P]),
    "stp_dropout": (C.c_int, [_TP, _F, _U64, C.c_uint32, _P, _TP, _P]),
    "stp_prob_head_fwd": (C.c_int, [_TP, _I32, _I32, _TP, _P]),
    "stp_prob_head_bwd": (C.c_int, [_TP, _TP, _TP, _I32, _I32, _TP, _P]),
    "stp_dwconv_fwd": (C.c_int, [_DDP, _TP, _P, _TP, _P]),
    "stp_dwconv_dgrad": (C.c_int, [_DDP, _TP, _P, _TP, _TP, _P]),
    "stp_dwconv_wgrad": (C.c_int, [_DDP, _TP, _TP, _P, _P, _SZ, _P]),
    "stp_dwconv_wgrad_workspace": (_SZ, [_DDP, _TP, _TP]),
    "stp_copy_up": (C.c_int, [_TP, _I32, _TP, _P]),
    "stp_add": (C.c_int, [_TP, _TP, _TP, _P]),
    "stp_upsample2x_bwd": (C.c_int, [_TP, _TP, _TP, _P]),
    "stp_resize_bilinear_fwd": (C.c_int, [_TP, _TP, _P]),
    "stp_resize_bilinear_bwd": (C.c_int, [_TP, _TP, _TP, _P]),
    "stp_resize_bilinear_ac_fwd": (C.c_int, [_TP, _TP, _P]),
    "stp_resize_bilinear_ac_bwd": (C.c_int, [_TP, _TP, _TP, _P]),
    "stp_loss_fwd": (C.c_int, [_P, _P, _I64, C.POINTER(LossSpec), _P, _P, _P]),
    "stp_loss_partial_floats": (_SZ, []),
    "stp_loss_bwd": (C.c_int, [_P, _P, _I64, C.POINTER(LossSpec), _P, _P, _P]),
    "stp_augment_pixel_ops": (C.c_int, [_P, _P, C.POINTER(AugPixSpec), C.c_uint64, _P, _I32, _I32, _I32, _I32, _P]),
    "stp_augment_neighbourhood_workspace": (_SZ, [_I32]),
    "stp_augment_neighbourhood": (C.c_int, [_P, _P, _P, C.POINTER(AugNbOp), C.c_uint64, _P, _I32, _I32, _I32, _I32, _P, _SZ, _P]),
    "stp_resize_u8": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _I32, _I32, _P]),
    "stp_croppad_draw": (C.c_int, [C.POINTER(CropPadSpec), C.c_uint64, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "stp_softmax_cce_fwd": (C.c_int, [_P, _P, _I64, _I32, _F, _I32, _P, _P, _P]),
    "stp_softmax_cce_bwd": (C.c_int, [_P, _P, _I64, _I32, _F, _I32, _P, _P]),
    "stp_lovasz_workspace": (_SZ, [_I32, _I64]),
    "stp_lovasz_fwd": (C.c_int, [_P, _P, _I32, _I64, _I32, _F, _I32, _P, _SZ, _P, _P]),
    "stp_lovasz_fwd_mc": (C.c_int, [_P, _P, _I32, _I64, _I32, _I32, _F, _I32, _P, _SZ, _P, _P]),
    "stp_lovasz_bwd": (C.c_int, [_P, _SZ, _I32, _I64, _F, _I32, _P, _P]),
    "stp_adam": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, C.POINTER(GradXform), _P, _P]),
    "stp_sgd": (C.c_int, [_P, _P, _P, _I64, _F, _F, _I32, C.POINTER(GradXform), _P]),
    "stp_rmsprop": (C.c_int, [_P, _P, _P, _I64, _F, _F, _F, C.POINTER(GradXform), _P]),
    "stp_nadam": (C.c_int, [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, C.POINTER(GradXform), _P, _P]),
    "stp_step_advance": (C.c_int, [_P, _P]),
    "stp_comm_unique_id": (C.c_int, [_P]),
    "stp_comm_init": (C.c_int, [_I32, _I32, _P, C.POINTER(C.c_void_p)]),
    "stp_allreduce": (C.c_int, [_P, _P, _I64, _P]),
    "stp_comm_destroy": (C.c_int, [_P]),
    "stp_sumsq": (C.c_int, [_P, _I64, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libstp.so; raise if it was not built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StpError("libstp.so not found at %s -- build it with `python -m segmentation_training_pipeline_b200.build` "
                       "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    # A/B knobs from the environment: STP_OPTIONS="bn_blocks=64,tc3=1" -> stp_set_option for each pair
    for kv in filter(None, os.environ.get("STP_OPTIONS", "").split(",")):
        k, v = kv.split("=")
        if lib.stp_set_option(k.strip().encode(), int(v)) != 0:
            raise StpError("STP_OPTIONS: unknown option " + k)
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().stp_last_error()
        raise StpError("%s failed (%d): %s" % (what or "libstp call", rc, msg.decode() if msg else "?"))


# functions whose int return value is a result, not a status code
_UNCHECKED = ("set_trace_buffer", "version", "tc_enabled", "bn_nblk", "last_error", "launch_count", "tc_launch_count", "tc3_launch_count", "set_tc_enabled",
              "conv_wgrad_workspace", "head_bwd_workspace", "head_fwd_workspace", "loss_partial_floats", "lovasz_workspace")


class Lib:
    """Thin checked call layer: `L.conv_fwd(...)` == check(stp_conv_fwd(...))."""

    def __init__(self):
        self._lib = load()

    def __getattr__(self, name):
        fn = getattr(self._lib, "stp_" + name)
        if name not in _UNCHECKED and fn.restype is C.c_int:
            def wrapped(*a, _fn=fn, _name=name):
                rc = _fn(*a)
                if rc != 0:
                    check(rc, "stp_" + _name)
                return rc
            setattr(self, name, wrapped)
            return wrapped
        setattr(self, name, fn)
        return fn
