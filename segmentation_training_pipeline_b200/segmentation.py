"""Host-side mirror of the reference's experiment API for the training hot path.

    from segmentation_pipeline import segmentation                      # alias package at the repo root
    from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet
    cfg = segmentation.parse("config.yaml"); cfg.fit(SimplePNGMaskDataSet(imgs, masks))

Reference: segmentation_pipeline/segmentation.py -- parse :211-214, PipelineConfig :35-208, createNet1 :96-155,
loss/metric registry :15-22, custom_models :31-33; schema keys schemas/segmentation.raml:26-136; augmenters
schemas/augmenters.raml:43-133; inherited fit/kfold/stages from musket_core.generic_config [DEP] as documented in
README.md:116-205 (5 shuffled folds, random_state, weights/, metrics/, summary.yaml next to the yaml).

Everything numeric runs in libstp (CUDA, no CPU fallback): `fit` raises if the library / a GPU is missing.
Mirrored training controls: k-fold x stage loop, best-weights checkpoint + CSV log, callbacks EarlyStopping / ReduceLROnPlateau /
CyclicLR, freeze_encoder / unfreeze_encoder, negatives / validation_negatives, initial_weights, extra_train_data,
setAllowResume, lr_find, crops.  NOT mirrored (out of the hot-path scope): other callbacks, DrawResults,
the xception DeepLab backbone (these raise NotImplementedError naming the key instead of being silently ignored).
"""
from __future__ import annotations

import ast
import csv
import os
import re
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import yaml

from . import models as _models
from .trainer import AugmentConfig

# architecture plugins: name -> callable(**kwargs) returning a model object (reference segmentation.py:31-33)
custom_models: Dict[str, Callable] = {}
# custom loss/metric registry stand-in for keras.utils.get_custom_objects() (reference README.md:629-634)
custom_objects: Dict[str, Callable] = {}
extra_train: Dict[str, object] = {}


def ansemblePredictions(sourceFolder, folders, cb, data, weights=None):
    """reference segmentation.py:27 / README.md:745-754: average the per-file .npy predictions of several runs."""
    from .predict import ansemble_predictions
    return ansemble_predictions(sourceFolder, folders, cb, data, weights)
dataset_augmenters: Dict[str, Callable] = {}

_LOSS_TERMS = {"binary_crossentropy": 0, "dice_loss": 1, "iou_loss": 2, "lovasz_loss": 3, "jaccard_loss": 4, "focal_loss": 5,
               "categorical_crossentropy": 6}
_UNFUSED_LOSSES = ()
_METRIC_ALIASES = {"binary_accuracy": "binary_accuracy", "dice": "dice", "iou": "iou", "iou_coef": "iou", "iot": "iot",
                   "iot_coef": "iot", "loss": "loss", "binary_crossentropy": "binary_crossentropy",
                   "categorical_crossentropy": "categorical_crossentropy", "categorical_accuracy": "categorical_accuracy",
                   "accuracy": "categorical_accuracy", "acc": "categorical_accuracy"}


def parse_loss(expr: str) -> Tuple[float, ...]:
    """'binary_crossentropy+0.1*dice_loss' (reference README.md:210-214) -> (w_bce, w_dice, w_iou); with lovasz_loss
    (segmentation.py:15-22 registry; computed on logits, the reference strips the final Activation) ->
    (0, 0, 0, w_lovasz) -- mixing it with the probability-based terms is undefined in the reference and rejected."""
    if not isinstance(expr, str) or not expr.strip():
        raise ValueError("loss must be a non-empty string")
    w = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]

    def term(node, scale):
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Add):
            term(node.left, scale)
            term(node.right, scale)
        elif isinstance(node, ast.BinOp) and isinstance(node.op, ast.Sub):
            term(node.left, scale)
            term(node.right, -scale)
        elif isinstance(node, ast.BinOp) and isinstance(node.op, ast.Mult):
            if isinstance(node.left, ast.Constant):
                term(node.right, scale * float(node.left.value))
            elif isinstance(node.right, ast.Constant):
                term(node.left, scale * float(node.right.value))
            else:
                raise ValueError("loss expression: product needs a numeric factor: " + expr)
        elif isinstance(node, ast.Name):
            if node.id in _LOSS_TERMS:
                w[_LOSS_TERMS[node.id]] += scale
            elif node.id in custom_objects:
                raise NotImplementedError("loss '%s' is a Python callable registered in custom_objects: the training step runs as "
                                          "fused device kernels, only the built-in loss names can be trained on" % node.id)
            elif node.id in _UNFUSED_LOSSES:
                raise NotImplementedError("loss '%s' has no fused device kernel yet (DESIGN.md, out of scope rows)" % node.id)
            else:
                raise ValueError("unknown loss '%s'" % node.id)
        else:
            raise ValueError("cannot parse loss expression: " + expr)

    term(ast.parse(expr.strip(), mode="eval").body, 1.0)
    if w[3] != 0.0 and (any(w[:3]) or any(w[4:])):
        raise ValueError("lovasz_loss works on logits and cannot be combined with probability-based losses: " + expr)
    if w[6] != 0.0 and any(w[:6]):
        raise ValueError("categorical_crossentropy works on softmax probabilities and cannot be combined with the sigmoid-based "
                         "losses: " + expr)
    n = 7
    while n > 3 and w[n - 1] == 0.0:   # (w_bce, w_dice, w_iou[, w_lovasz[, w_jaccard[, w_focal[, w_cce]]]])
        n -= 1
    return tuple(w[:n])


def _rng(v, cast=float):
    if isinstance(v, (list, tuple)):
        if len(v) != 2:
            raise ValueError("range needs 2 values: %r" % (v,))
        return cast(v[0]), cast(v[1])
    return cast(v), cast(v)


def parse_augmentation(spec: Optional[dict], seed: int = 0) -> AugmentConfig:
    """YAML `augmentation:` (reference README.md:249-268, augmenters.raml positional convention :24-40) -> the fused
    device augmenter.  Only the augmenters the K1 kernel fuses are accepted; anything else raises."""
    cfg = AugmentConfig(seed=seed)
    if not spec:
        return cfg
    spec, groups = _flatten_augmentation(spec)
    # imgaug Sequential applies the augmenters in YAML order.  The fused kernel runs {Rotate90, Fliplr, Flipud} (ANY order among
    # themselves: they are index permutations, composed exactly -- AugmentConfig.flip_before_rot90) -> Affine -> colour stage
    # (Multiply / Add / Invert in ANY order among themselves); a block in another order would silently compute something
    # else, so it is rejected.
    # The crop / pad family (Pad, PadToFixedSize, CropToFixedSize, CropAndPad) leads the block: the ops are composed into one
    # window per sample that the cv2-arithmetic resize brings back to `shape` (trainer.run_augment).
    rank = {"Pad": 0, "PadToFixedSize": 0, "CropToFixedSize": 0, "CropAndPad": 0,
            "Rotate90": 1, "Fliplr": 1, "Flipud": 1, "Affine": 2, "Multiply": 3, "Add": 3, "Invert": 3,
            "AddElementwise": 3, "MultiplyElementwise": 3, "Dropout": 3, "AdditiveGaussianNoise": 3, "Grayscale": 3,
            "GaussianBlur": 3, "AverageBlur": 3, "MedianBlur": 3, "Sharpen": 3, "Emboss": 3, "EdgeDetect": 3, "DirectedEdgeDetect": 3}
    NB = ("GaussianBlur", "AverageBlur", "MedianBlur", "Sharpen", "Emboss", "EdgeDetect", "DirectedEdgeDetect")
    for name in groups:
        if rank.get(name) != 3:
            raise NotImplementedError("OneOf over '%s': only the colour-block augmenters (Multiply, Add, Invert, AddElementwise, "
                                      "MultiplyElementwise, Dropout, AdditiveGaussianNoise, Grayscale, GaussianBlur, AverageBlur, "
                                      "MedianBlur, Sharpen, Emboss, EdgeDetect) can be OneOf members" % name)
    last, colour = -1, []
    for name in spec:
        if name not in rank:
            raise NotImplementedError("augmenter '%s' is not fused on device (supported: %s)" % (name, ", ".join(rank)))
        if rank[name] < last:
            raise NotImplementedError("augmentation order %s is not the fused kernel's (Pad/PadToFixedSize/CropToFixedSize/"
                                      "CropAndPad, Rotate90/Fliplr/Flipud, Affine, then Multiply/Add/Invert)" % list(spec))
        last = rank[name]
        if rank[name] == 3 and name in ("Multiply", "Add", "Invert"):
            colour.append({"Multiply": 0, "Add": 1, "Invert": 2}[name])
    cfg.color_order = tuple(colour + [o for o in (0, 1, 2) if o not in colour])
    extended = [n for n in spec if rank.get(n) == 3 and n not in ("Multiply", "Add", "Invert")]   # incl. the neighbourhood augmenters
    if extended or groups:
        # pixel-wise colour stage in YAML order (csrc/augment.cu augment_pixel_ops_kernel)
        from . import lib as _lib
        ops, seq = [], []
        for name, val in spec.items():
            if rank.get(name) != 3:
                continue
            v = val if isinstance(val, dict) else {}
            pc = float(v.get("per_channel", 0.0) or 0.0)
            gid, gsz, gm = groups.get(name, (0, 0, 0))
            if name in NB:
                # neighbourhood augmenters (csrc/augment_nb.cu): (kind, a, b, c, d, group...) -- first / second parameter ranges
                one = (lambda key, default: _rng(v.get(key, default) if isinstance(val, dict) else (val if val is not None else default)))
                c = d = 0.0
                if name == "GaussianBlur":
                    a, b = one("sigma", 0.0)
                    if not 0.0 <= a <= b or 2.6 * b >= 25:
                        raise ValueError("GaussianBlur: sigma must lie in [0, 9.6)")
                elif name in ("AverageBlur", "MedianBlur"):
                    a, b = _rng(v.get("k", 1) if isinstance(val, dict) else val, int)
                    if not 0 <= a <= b or b > (7 if name == "MedianBlur" else 31):
                        raise ValueError("%s: k must lie in [0, %d]" % (name, 7 if name == "MedianBlur" else 31))
                    a, b = (max(a, 1), max(b, 1)) if name == "MedianBlur" else (a, b)
                else:
                    a, b = _rng(v.get("alpha", 0.0) if isinstance(val, dict) else val)
                    if not 0.0 <= a <= b <= 1.0:
                        raise ValueError("%s: alpha must lie in [0, 1]" % name)
                    if name == "Sharpen":
                        c, d = _rng(v.get("lightness", 1.0))
                    elif name == "Emboss":
                        c, d = _rng(v.get("strength", 1.0))
                    elif name == "DirectedEdgeDetect":
                        c, d = _rng(v.get("direction", [0.0, 1.0]))
                seq.append(("nb", (_lib.NB_KINDS[name], float(a), float(b), float(c), float(d), gid, gsz, gm)))
                continue
            if name in ("Multiply", "Add", "Invert"):
                a = b = 0.0
            elif name in ("AddElementwise", "MultiplyElementwise"):
                a, b = _rng(v.get("range", val) if isinstance(val, dict) else val, int if name == "AddElementwise" else float)
            elif name == "Dropout":
                a, b = _rng(v.get("p", 0.0) if isinstance(val, dict) else val)
                if not 0.0 <= a <= b <= 1.0:
                    raise ValueError("Dropout: p must lie in [0, 1]")
            elif name == "AdditiveGaussianNoise":
                a, b = _rng(v.get("scale", 0.0) if isinstance(val, dict) else val)
            else:   # Grayscale
                a, b = _rng(v.get("alpha", 1.0) if isinstance(val, dict) else val)
            ops.append((_lib.PIX_KINDS[name], pc, float(a), float(b), gid, gsz, gm))
            seq.append(("pix", ops[-1]))
        if len(ops) > 8 or len(seq) > 16:
            raise NotImplementedError("augmentation: at most 8 pixel-wise and 16 colour-block augmenters in total")
        cfg.pix_ops = tuple(ops)
        if any(t == "nb" for t, _ in seq):
            cfg.colour_seq = tuple(seq)
    names = list(spec)
    if "Rotate90" in names:   # reference examples/people/*.yaml list Fliplr, Flipud, Rotate90
        r = names.index("Rotate90")
        cfg.flip_before_rot90 = (1 if "Fliplr" in names and names.index("Fliplr") < r else 0) | \
                                (2 if "Flipud" in names and names.index("Flipud") < r else 0)
    cp, keep_size_seen = [], False
    for name, val in spec.items():
        if rank.get(name) != 0:
            continue
        if keep_size_seen:
            # imgaug Pad / CropAndPad default to keep_size=True (resample back to the input size); composed with the final
            # Resize that is exact only when nothing else of the family follows
            raise NotImplementedError("augmentation: %s after Pad / CropAndPad (keep_size resampling between crop / pad ops is "
                                      "not built)" % name)
        if len(cp) >= 4:
            raise NotImplementedError("augmentation: at most 4 crop / pad augmenters")
        val = val if isinstance(val, dict) else ({"px": val} if name == "Pad" else {"percent": val} if name == "CropAndPad" else val)
        if name == "Pad":
            px = val.get("px")
            px = [px] * 4 if isinstance(px, int) else list(px or [])
            if len(px) != 4 or any(int(v) < 0 for v in px):
                raise ValueError("Pad: px must be one non-negative integer or [top, right, bottom, left]")
            cp.append((1, 0, int(px[0]), int(px[1]), int(px[2]), int(px[3])))
            keep_size_seen = True
        elif name in ("PadToFixedSize", "CropToFixedSize"):
            if not isinstance(val, dict) or "width" not in val or "height" not in val:
                raise ValueError("%s needs width and height" % name)
            cp.append((2 if name == "PadToFixedSize" else 3, 0, int(val["width"]), int(val["height"]), 0, 0))
        else:
            pc = val.get("percent")
            pc = [pc] if isinstance(pc, (int, float)) else list(pc or [])
            if len(pc) == 1:
                cp.append((4, 0, float(pc[0]), float(pc[0]), float(pc[0]), float(pc[0])))
            elif len(pc) == 2:
                cp.append((4, 1, float(pc[0]), float(pc[1]), 0.0, 0.0))     # a range: one draw per side and sample
            elif len(pc) == 4:
                cp.append((4, 0, float(pc[0]), float(pc[1]), float(pc[2]), float(pc[3])))
            else:
                raise ValueError("CropAndPad: percent must hold 1, 2 (range) or 4 (top, right, bottom, left) numbers")
            if min(pc) <= -0.5:
                raise ValueError("CropAndPad: crops of 50 % or more per side leave no image")
            keep_size_seen = True
    cfg.crop_pad = tuple(cp)
    for name, val in spec.items():
        if name == "Fliplr":
            cfg.fliplr = float(val)
        elif name == "Flipud":
            cfg.flipud = float(val)
        elif name == "Rotate90":
            cfg.rot90 = bool(val)
        elif name == "Invert":
            cfg.invert = float(val["p"] if isinstance(val, dict) else val)
        elif name == "Affine":
            val = val or {}
            cfg.affine = True
            if "scale" in val:
                cfg.scale = _rng(val["scale"])
            tp = val.get("translate_percent")
            if tp is not None:
                if isinstance(tp, dict):
                    cfg.translate_x = _rng(tp.get("x", 0.0))
                    cfg.translate_y = _rng(tp.get("y", 0.0))
                else:
                    cfg.translate_x = cfg.translate_y = _rng(tp)
            if "rotate" in val:
                cfg.rotate = _rng(val["rotate"])
            if "shear" in val:
                cfg.shear = _rng(val["shear"])
            bad = set(val) - {"scale", "translate_percent", "rotate", "shear"}
            if bad:
                raise NotImplementedError("Affine keys not fused on device: %s" % sorted(bad))
        elif name == "Multiply":
            cfg.multiply = _rng(val.get("range", val.get("mul")) if isinstance(val, dict) else val)
        elif name == "Add":
            cfg.add = _rng(val.get("range", val.get("value")) if isinstance(val, dict) else val, int)
    return cfg


def _flatten_augmentation(spec):
    """`Sequential` (schemas/augmenters.raml:56, ControlFlow: a list of augmenters) is spliced into the block in place; `OneOf`
    (:72) likewise, its members recorded as a group (group id, size, member index) so that exactly one of them runs per sample.
    Children are written as a list of one-key mappings or as one mapping.  Returns (ordered dict, {name: group})."""
    out, groups, gid = {}, {}, [0]

    def children(val):
        if isinstance(val, dict):
            return list(val.items())
        items = []
        for c in val or []:
            if not isinstance(c, dict) or len(c) != 1:
                raise ValueError("Sequential / OneOf children must be one-key mappings, got %r" % (c,))
            items.extend(c.items())
        return items

    def walk(items, group=None):
        for name, val in items:
            if name == "Sequential":
                if group is not None:
                    raise NotImplementedError("Sequential inside OneOf is not built")
                walk(children(val))
            elif name == "OneOf":
                if group is not None:
                    raise NotImplementedError("nested OneOf is not built")
                ch = children(val)
                if not ch:
                    continue
                g = gid[0]
                gid[0] += 1
                if g >= 16:
                    raise NotImplementedError("at most 16 OneOf groups")
                walk([(n, v) for n, v in ch], group=(g, len(ch)))
            else:
                if name in out:
                    raise NotImplementedError("augmentation: '%s' appears twice (each augmenter can be used once per block)" % name)
                out[name] = val
                if group is not None:
                    groups[name] = (group[0], group[1], sum(1 for n in groups if groups[n][0] == group[0]))

    walk(list(spec.items()) if isinstance(spec, dict) else children(spec))
    return out, groups


class PipelineConfig:
    """Mirrors the attributes/verbs of the reference PipelineConfig that the training path uses."""

    def __init__(self, **atrs):
        self.architecture = atrs.pop("architecture", None)
        self.backbone = atrs.pop("backbone", "resnet34")
        self.classes = int(atrs.pop("classes", 1))
        self.activation = atrs.pop("activation", "sigmoid")
        self.shape = list(atrs.pop("shape", [512, 512, 3]))
        self.encoder_weights = atrs.pop("encoder_weights", None)
        self.freeze_encoder = bool(atrs.pop("freeze_encoder", False))
        self.augmentation = atrs.pop("augmentation", None) or {}
        self.transforms = atrs.pop("transforms", None)
        self.optimizer = atrs.pop("optimizer", "Adam")
        self.lr = atrs.pop("lr", None)
        self.clipnorm = atrs.pop("clipnorm", None)
        self.clipvalue = atrs.pop("clipvalue", None)
        self.loss = atrs.pop("loss", "binary_crossentropy")
        self.batch = int(atrs.pop("batch", 16))                     # schema default, segmentation.raml:93-97
        self.metrics = list(atrs.pop("metrics", []) or [])
        self.primary_metric = atrs.pop("primary_metric", "val_loss")
        self.primary_metric_mode = atrs.pop("primary_metric_mode", "auto")
        self.stages = list(atrs.pop("stages", [{"epochs": 1}]) or [{"epochs": 1}])
        self.folds_count = int(atrs.pop("folds_count", 5))
        self.random_state = int(atrs.pop("random_state", 33))
        self.testSplit = float(atrs.pop("testSplit", 0.0) or 0.0)
        self.decoder_filters = tuple(atrs.pop("decoder_filters", (256, 128, 64, 32, 16)))
        self.decoder_block_type = atrs.pop("decoder_block_type", "upsampling")   # segmentation.raml:162-165
        self.callbacks = atrs.pop("callbacks", None)
        self.crops = int(atrs.pop("crops", 0) or 0)   # N: train / predict on the N x N cells of every image (README.md:471-491)
        self.datasets = atrs.pop("datasets", None)
        self.fit_with = atrs.pop("fit_with", None)
        self.extra = atrs            # accepted, unused keys (kept so configs round-trip)
        self.path: Optional[str] = None
        self.gpus = 1                # FAQ.md:108-112 `cfg.gpus = N`; data parallel = 1 process / GPU here (see ddp.py)
        self.showDataExamples = False
        self.allowResume = False
        self.device = "cuda:0"
        self.device_augment = True

    # -- reference verbs ---------------------------------------------------------------------------
    def setAllowResume(self, v: bool = True):
        self.allowResume = bool(v)

    def net_shape(self) -> List[int]:
        """Input shape of the network: `shape`, or -- with `crops: N` -- the CELL shape (shape[0]//N, shape[1]//N, C), exactly
        what the reference's createNet1 builds (segmentation.py:131-132): cells are trained and predicted at cell resolution."""
        if self.crops and self.crops > 1:
            return [int(self.shape[0]) // self.crops, int(self.shape[1]) // self.crops, int(self.shape[2])]
        return list(self.shape)

    def createNet(self, batch: Optional[int] = None, loss: Optional[str] = None):
        """YAML keys -> engine graph (reference createNet1, segmentation.py:96-155): unknown names raise the same
        ValueErrors after printing the known lists."""
        arch = self.architecture
        if arch in custom_models:
            return custom_models[arch](backbone=self.backbone, classes=self.classes, input_shape=tuple(self.net_shape()),
                                       activation=self.activation)
        if arch not in _models.KNOWN_ARCHITECTURES:
            print("Unknown architecture:" + str(arch))
            print("Known architectures:", _models.KNOWN_ARCHITECTURES + sorted(custom_models))
            raise ValueError("Unknown architecture")
        bb = str(self.backbone).lower()
        deeplab = arch == "DeepLabV3"   # the reference's in-tree model (custom_models, segmentation.py:31-33; impl/deeplab/model.py)
        known = _models.DEEPLAB_BACKBONES if deeplab else _models.KNOWN_BACKBONES
        if bb not in known:
            print("Unknown backbone:" + bb)
            print("Known backbones:", known)
            raise ValueError("Unknown backbone")
        lw = parse_loss(loss or self.loss)
        pure_lovasz = len(lw) == 4 and lw[3] != 0.0
        pure_cce = len(lw) == 7 and lw[6] != 0.0
        if pure_cce and (self.activation != "softmax" or self.classes < 2):
            raise ValueError("categorical_crossentropy needs `activation: softmax` and classes >= 2 (schema segmentation.raml:12-21, 62-63)")
        if self.activation == "softmax" and not (pure_lovasz or pure_cce):
            # lovasz_loss is computed on LOGITS (the reference strips the final Activation) and categorical_crossentropy has its
            # own fused softmax kernel; the sigmoid-based loss kernels (binary_crossentropy, dice, ...) do not apply to softmax
            raise NotImplementedError("activation: softmax is built for categorical_crossentropy and lovasz_loss (the dice / iou / "
                                      "jaccard / focal / binary_crossentropy kernels fuse a sigmoid)")
        if self.activation not in ("sigmoid", "softmax", None, "none"):
            raise NotImplementedError("activation '%s' is not built" % self.activation)
        if self.classes > 4:
            raise NotImplementedError("classes <= 4 (mask channels of the on-device augmentation / head kernels)")
        if deeplab and pure_lovasz:
            raise NotImplementedError("lovasz_loss with DeepLabV3: the activation sits inside the last Conv2D there, the reference's "
                                      "compile cannot strip it; not built")
        if deeplab and self.activation not in ("sigmoid", "softmax"):
            raise NotImplementedError("DeepLabV3: activation sigmoid or softmax (the head resizes probabilities)")
        enc_file = None
        if self.encoder_weights not in (None, "None", "none"):
            # `encoder_weights: imagenet` downloads a Keras checkpoint in the reference; there is no network here.  A path to a
            # local .npz with Keras-named arrays (conv0/kernel, bn0/gamma, bn0/moving_mean, ...) is accepted instead.
            cand = str(self.encoder_weights)
            base = os.path.dirname(os.path.abspath(self.path)) if self.path else os.getcwd()
            names = [cand, cand + ".npz", os.path.join(base, cand), os.path.join(base, cand + ".npz")]
            if deeplab and cand == "pascal_voc":
                # the reference fetches this file into ~/.keras/models (impl/deeplab/model.py:505-512) and loads it by layer
                # name; the same arrays as .npz (scripts/keras_h5_to_npz.py converts) are looked up in the same places
                stem = "deeplabv3_%s_tf_dim_ordering_tf_kernels.npz" % bb
                names += [os.path.join(base, stem), os.path.join(os.path.expanduser("~"), ".keras", "models", stem)]
            for c in names:
                if os.path.isfile(c):
                    enc_file = c
                    break
            if enc_file is None:
                raise NotImplementedError("encoder_weights: '%s' cannot be downloaded here -- give the path of a local .npz with the "
                                          "encoder's Keras-named arrays" % cand)
        mk = self._model_kwargs(arch)
        net = _models.SegNet(bb, classes=self.classes, input_shape=tuple(self.net_shape()), batch=batch or self.batch,
                              decoder_filters=self.decoder_filters, device=self.device, seed=self.random_state,
                              architecture=arch, decoder_block_type=getattr(self, "decoder_block_type", None) or "upsampling",
                              pyramid_block_filters=int(mk.get("pyramid_block_filters") or 256),
                              segmentation_block_filters=int(mk.get("segmentation_block_filters") or 128),
                              dropout=mk.get("dropout") or None,
                              decoder_use_batchnorm=bool(mk.get("use_batchnorm", True)) if arch != "PSPNet" else True,
                              downsample_factor=int(mk.get("downsample_factor") or 8),
                              psp_conv_filters=int(mk.get("psp_conv_filters") or 512),
                              precision=str(self.extra.get("precision", "bf16")),
                              activation=self.activation, OS=int(mk.get("OS") or 16), loss=lw)
        net.activation = self.activation or "linear"   # what predict applies to the logits
        if enc_file is not None and deeplab:
            # model.load_weights(path, by_name=True): every layer whose name and shapes match, ASPP included; the class layer
            # only when classes == 21 (its name differs otherwise)
            cur = net.get_weights()
            w = {k: v for k, v in dict(np.load(enc_file)).items() if k in cur}
            if not w:
                raise ValueError("encoder_weights: %s holds no array of this model" % enc_file)
            net.set_weights(self._adapt_input_channels(net, w), strict=False)
        elif enc_file is not None:
            layers = {n.rsplit("/", 1)[0] for n in net.encoder_param_names}
            w = {k: v for k, v in dict(np.load(enc_file)).items() if k.rsplit("/", 1)[0] in layers}
            if not w:
                raise ValueError("encoder_weights: %s holds no array of this encoder (%s, ...)" % (enc_file, sorted(layers)[:3]))
            w = self._adapt_input_channels(net, w)
            net.set_weights(w, strict=False)
        return net

    def _adapt_input_channels(self, net, w: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        """More than 3 input channels with pretrained 3-channel encoder weights (reference createNet1, segmentation.py:138-153):
        the reference builds the 3-channel model WITH weights and the N-channel one without, copies every layer across
        (`adaptNet` / `copyWeights`, musket_core [DEP, unpinned]) and caches the result in `<config>.mdl-nchannel`.  Here: every
        array whose shape matches is taken as is; a first-layer kernel (kh, kw, 3, f) is widened to (kh, kw, C, f) -- the RGB
        planes are copied and every extra plane is their mean (config key `nchannel_fill: mean | zero`); the per-channel
        arrays of the input BatchNorm (bn_data) are extended the same way.  The adapted dict is cached as
        `<config>.mdl-nchannel.npz` and reused on the next build."""
        C = int(self.shape[2])
        if C <= 3:
            return w
        cache = (self.path + ".mdl-nchannel.npz") if self.path else None
        if cache and os.path.exists(cache):
            return dict(np.load(cache))
        fill = str(self.extra.get("nchannel_fill", "mean"))
        if fill not in ("mean", "zero"):
            raise ValueError("nchannel_fill must be 'mean' or 'zero'")
        cur = net.get_weights()
        out = {}
        for k, v in w.items():
            tgt = cur.get(k)
            if tgt is None or tgt.shape == v.shape:
                out[k] = v
            elif v.ndim == 4 and tgt.ndim == 4 and v.shape[:2] == tgt.shape[:2] and v.shape[3] == tgt.shape[3] and v.shape[2] == 3:
                extra = v.mean(axis=2, keepdims=True) if fill == "mean" else np.zeros_like(v[:, :, :1])
                out[k] = np.concatenate([v] + [extra] * (tgt.shape[2] - 3), axis=2).astype(np.float32)
            elif v.ndim == 1 and tgt.ndim == 1 and v.shape[0] == 3 and tgt.shape[0] == C:
                ext = np.full(C - 3, v.mean() if fill == "mean" else (1.0 if k.endswith("moving_variance") else 0.0), np.float32)
                out[k] = np.concatenate([v, ext]).astype(np.float32)
            else:
                raise ValueError("encoder_weights: %s has shape %s, the %d-channel encoder needs %s" % (k, v.shape, C, tgt.shape))
        if cache:
            np.savez(cache, **out)
        return out

    def _model_kwargs(self, arch: str) -> Dict[str, object]:
        """The architecture's schema-typed keyword arguments (schemas/segmentation.raml:158-248) as the reference's createNet1
        would forward them (segmentation.py:119-129), each one either honoured by the engine graph or rejected BY NAME:
        nothing the schema types as a model argument is dropped silently."""
        from . import schema as _schema
        given = dict(self.extra)
        given["decoder_filters"], given["decoder_block_type"] = list(self.decoder_filters), self.decoder_block_type
        mk = _schema.resolve(arch, given)

        def only(key, allowed, what):
            v = mk.get(key)
            if v is not None and v not in allowed:
                raise NotImplementedError("%s: %r is not built for %s (%s; schema default %r)" %
                                          (key, v, arch, what, _schema.model_keys(arch)[key][1]))

        if arch == "Unet":
            only("n_upsample_blocks", (5,), "the decoder has one block per encoder stage")
            only("upsample_rates", ([2, 2, 2, 2, 2], (2, 2, 2, 2, 2)), "every decoder block upsamples x2")
            if mk.get("use_batchnorm") is False and self.decoder_block_type == "transpose":
                raise NotImplementedError("use_batchnorm / decoder_use_batchnorm: false is built for decoder_block_type: upsampling only")
        elif arch == "FPN":
            only("upsample_rates", ([2, 2, 2], (2, 2, 2)), "the top-down pathway upsamples x2 per level")
            only("last_upsample", (4,), "logits are upsampled x4")
            only("interpolation", ("bilinear",), "segmentation branches use TF1 bilinear resize")
            only("use_batchnorm", (True,), "the FPN blocks are conv + BatchNorm + ReLU")
            only("dropout", (0, 0.0, None), "SpatialDropout2D is not built")
        elif arch == "PSPNet":
            only("psp_pooling_type", ("avg",), "pyramid levels use average pooling")
            only("use_batchnorm", (True,), "the PSP blocks are conv + BatchNorm + ReLU")
            only("dropout", (0, 0.0, None), "SpatialDropout2D is not built")
            only("final_interpolation", ("bilinear",), "logits are upsampled bilinearly")
            only("downsample_factor", (4, 8, 16), "feature layers at 1/4, 1/8 or 1/16")
        elif arch == "DeepLabV3":
            only("alpha", (1, 1.0), "MobileNetV2 width multiplier 1")   # OS applies to xception only (model.py:339-349, 384-385)
            only("OS", (8, 16), "output stride 8 or 16")
        elif arch == "Linknet":
            only("use_batchnorm", (True,), "the Linknet blocks are conv + BatchNorm + ReLU")
            only("n_upsample_blocks", (5,), "the decoder has one block per encoder stage")
            only("upsample_layer", ("upsampling",), "decoder blocks use UpSampling2D")
            only("upsample_kernel_size", ([3, 3], (3, 3)), "only used by upsample_layer: transpose")
        return mk

    def kfold(self, n: int) -> List[Tuple[np.ndarray, np.ndarray]]:
        """sklearn KFold(folds_count, shuffle=True, random_state) as the reference's ImageKFoldedDataSet [DEP]."""
        from sklearn.model_selection import KFold
        idx = np.arange(n)
        return [(tr, te) for tr, te in KFold(n_splits=self.folds_count, shuffle=True, random_state=self.random_state).split(idx)]

    def _dir(self):
        if self.path is None:
            raise ValueError("cfg.path is not set (use segmentation.parse)")
        return os.path.dirname(os.path.abspath(self.path))

    def _resolve_dataset(self, d):
        if d is not None:
            return d
        if self.fit_with and self.datasets and self.fit_with in self.datasets:
            from .impl.datasets import dataset_from_spec
            spec = self.datasets[self.fit_with]
            if not isinstance(spec, dict):   # e.g. `composite: ["default"]` (ds_2.yaml:41): a list of tensor names, not a dataset
                raise ValueError("datasets: '%s' is not a dataset declaration" % self.fit_with)
            return dataset_from_spec(spec, self._dir())
        raise ValueError("fit() needs a dataset or `datasets:` + `fit_with:` in the config")

    def _check_training_keys(self):
        """keys the training path does not implement must stop the run, not be silently ignored"""
        if self.transforms:
            raise NotImplementedError("transforms: (augmenters applied to every sample, also at validation / prediction time) is "
                                      "not built; resize-to-shape is implicit")
        for k in ("dataset_augmenter", "bgr", "manualResize", "compressPredictionsAsInts"):
            if self.extra.get(k):
                raise NotImplementedError("config key '%s' is not built (DESIGN.md section 7)" % k)

    def fit(self, d=None, subsample=1.0, foldsToExecute: Optional[Sequence[int]] = None, start_from_stage=0):
        from .fit import run_fit
        self._check_training_keys()
        return run_fit(self, self._resolve_dataset(d), subsample, foldsToExecute, start_from_stage)

    def lr_find(self, d=None, start_lr=0.00001, end_lr=1.0, epochs=1, stage=0):
        """Learning-rate range test (reference README.md:455-470): returns an object with lrs / losses / plot_loss /
        plot_loss_change."""
        from .fit import run_lr_find
        return run_lr_find(self, self._resolve_dataset(d), start_lr, end_lr, epochs, stage)

    def load_model(self, fold: int = 0, stage: int = -1):
        """Engine graph with the best weights of (fold, stage) (reference README.md:553)."""
        if stage < 0:
            stage = len(self.stages) - 1
        net = self.createNet()
        p = os.path.join(self._dir(), "weights", "best-%d.%d.weights.npz" % (fold, stage))
        w = dict(np.load(p))
        net.set_weights(w)
        return net

    # -- inference verbs (reference segmentation.py:62-91, 158-191; README.md:493-534) ------------------------------
    def predict_on_directory(self, spath, fold=0, stage=0, limit=-1, batch_size=32, ttflips=False):
        from . import predict as _p
        return _p.predict_on_directory(self, spath, fold, stage, limit, batch_size, ttflips)

    def predict_to_directory(self, spath, tpath, fold=0, stage=0, limit=-1, batchSize=32, binaryArray=False, ttflips=False):
        from . import predict as _p
        return _p.predict_to_directory(self, spath, tpath, fold, stage, limit, batchSize, binaryArray, ttflips)

    def predict_in_directory(self, spath, fold, stage, cb, data, limit=-1, batchSize=32, ttflips=False):
        from . import predict as _p
        return _p.predict_in_directory(self, spath, fold, stage, cb, data, limit, batchSize, ttflips)

    def evaluate(self, d, fold, stage, negatives="all", limit=16):
        from . import predict as _p
        return _p.evaluate(self, d, fold, stage, negatives, limit)

    def evaluateAll(self, ds, fold=None, stage=-1, negatives="real", ttflips=None, batchSize=32):
        from . import predict as _p
        return _p.evaluate_all(self, ds, fold, stage, negatives, ttflips, batchSize)

    def info(self):
        """aggregated best metric per fold/stage from metrics/*.csv (reference FAQ.md:63-70)."""
        out = []
        mdir = os.path.join(self._dir(), "metrics")
        if not os.path.isdir(mdir):
            return out
        for f in sorted(os.listdir(mdir)):
            m = re.match(r"metrics-(\d+)\.(\d+)\.csv$", f)
            if not m:
                continue
            rows = list(csv.DictReader(open(os.path.join(mdir, f))))
            if rows:
                out.append({"fold": int(m.group(1)), "stage": int(m.group(2)), "epochs": len(rows), "last": rows[-1]})
        return out


def parse(path) -> PipelineConfig:
    """reference segmentation.py:211-214."""
    with open(path) as f:
        atrs = yaml.safe_load(f) or {}
    cfg = PipelineConfig(**atrs)
    cfg.path = path
    return cfg
