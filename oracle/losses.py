"""Losses / metrics registered by the reference at segmentation.py:15-22 -> musket_core.losses + Keras [DEP].

TEST INFRASTRUCTURE.  Formulas per SURVEY.md section 8 a-6 / Appendix B; musket_core is absent, so the reference's exact choices
(smoothing constants, focal reduction, the Lovasz activation) are parity unpinned and kept as named flags.  What IS pinned
(tests/test_cpu_oracle_optim.py): the cross-entropies against torch's losses on Keras-clipped probabilities, the focal formula
against torchvision.ops.sigmoid_focal_loss, the Lovasz hinge against the set-based definition of the Lovasz extension.  All take
(y_true, y_pred) as float NHWC tensors like the Keras callables, and return a scalar (Keras reduces
per-sample losses by mean over every remaining axis and over the batch).
"""
from __future__ import annotations

import ast
import re
from typing import Callable, Dict, List, Tuple

import torch

EPS = 1e-7  # keras.backend.epsilon()


def binary_crossentropy(y_true, y_pred):
    """keras.losses.binary_crossentropy on probabilities with the TF backend:
    p<-clip(p,eps,1-eps); x=log(p/(1-p)); l=max(x,0)-x*t+log1p(exp(-|x|)); mean(axis=-1) then mean."""
    p = torch.clamp(y_pred, EPS, 1 - EPS)
    x = torch.log(p / (1 - p))
    l = torch.clamp(x, min=0) - x * y_true + torch.log1p(torch.exp(-torch.abs(x)))
    return l.mean(dim=-1).mean()


def categorical_crossentropy(y_true, y_pred):
    p = y_pred / y_pred.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, EPS, 1 - EPS)
    return (-(y_true * torch.log(p)).sum(dim=-1)).mean()


def dice(y_true, y_pred):
    """musket_core.losses.dice: (2*sum(p*t)+1)/(sum(p)+sum(t)+1) over the whole batch flattened."""
    inter = (y_true * y_pred).sum()
    return (2.0 * inter + 1.0) / (y_true.sum() + y_pred.sum() + 1.0)


def dice_loss(y_true, y_pred):
    return 1.0 - dice(y_true, y_pred)


def iou(y_true, y_pred, smooth=1.0):
    inter = (y_true * y_pred).sum()
    union = y_true.sum() + y_pred.sum() - inter
    return (inter + smooth) / (union + smooth)


def iou_loss(y_true, y_pred):
    return 1.0 - iou(y_true, y_pred)


def iot(y_true, y_pred):
    return iou(y_true, (y_pred > 0.5).float())


def jaccard_loss(y_true, y_pred, smooth=100.0):
    inter = (y_true * y_pred).abs().sum(dim=-1)
    s = (y_true.abs() + y_pred.abs()).sum(dim=-1)
    jac = (inter + smooth) / (s - inter + smooth)
    return ((1 - jac) * smooth).mean()


def focal_loss(y_true, y_pred, gamma=2.0, alpha=0.75, reduction="mean"):
    """musket_core.losses.focal_loss [DEP, unpinned].  reduction: "mean" (TAKEN by the engine and by lookup("focal_loss"):
    Keras reduces a per-pixel loss by mean, consistent with every other loss here) or "sum" (the Keras-RetinaNet-style
    K.sum() form some musket_core versions use: same gradient direction, scaled by the element count) -- the doubtful [DEP]
    choice is this named flag."""
    p = torch.clamp(y_pred, EPS, 1 - EPS)
    pt1 = torch.where(y_true == 1, p, torch.ones_like(p))
    pt0 = torch.where(y_true == 0, p, torch.zeros_like(p))
    l = -(alpha * (1 - pt1) ** gamma * torch.log(pt1)) - ((1 - alpha) * pt0 ** gamma * torch.log(1 - pt0))
    if reduction == "sum":
        return l.sum()
    if reduction != "mean":
        raise ValueError("reduction must be 'mean' or 'sum'")
    return l.mean(dim=-1).mean()


def binary_accuracy(y_true, y_pred):
    """keras.metrics.binary_accuracy: mean(round(p) == t)."""
    return (torch.round(y_pred) == y_true).float().mean()


def _lovasz_grad(gt_sorted):
    gts = gt_sorted.sum()
    inter = gts - gt_sorted.cumsum(0)
    union = gts + (1 - gt_sorted).cumsum(0)
    jac = 1.0 - inter / union
    if gt_sorted.numel() > 1:
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
    return jac


def lovasz_hinge_flat(logits, labels, act="elu"):
    """Berman's lovasz_hinge_flat; act='elu' is the Kaggle-TGS `elu(e)+1` variant musket is believed to
    copy, act='relu' is Berman's original (exposed as a flag, [DEP] unpinned)."""
    signs = 2.0 * labels - 1.0
    errors = 1.0 - logits * signs
    errors_sorted, perm = torch.sort(errors, dim=0, descending=True, stable=True)
    gt_sorted = labels[perm]
    grad = _lovasz_grad(gt_sorted)
    a = torch.nn.functional.elu(errors_sorted) + 1.0 if act == "elu" else torch.relu(errors_sorted)
    return torch.dot(a, grad.detach())


def lovasz_loss(y_true, y_pred_logits, act="elu"):
    """musket_core.losses.lovasz_loss: squeeze(-1), logits, per_image=True, mean over images.
    For C>1 the reference is undefined (K.squeeze fails); defined here as per-class hinge averaged."""
    N, C = y_true.shape[0], y_true.shape[-1]
    tot = 0.0
    for n in range(N):
        for c in range(C):
            tot = tot + lovasz_hinge_flat(y_pred_logits[n, ..., c].reshape(-1), y_true[n, ..., c].reshape(-1), act)
    return tot / (N * C)


REGISTRY: Dict[str, Callable] = {
    "binary_crossentropy": binary_crossentropy, "categorical_crossentropy": categorical_crossentropy,
    "dice": dice, "dice_loss": dice_loss, "iou": iou, "iou_loss": iou_loss, "iot": iot,
    "jaccard_loss": jaccard_loss, "focal_loss": focal_loss, "lovasz_loss": lovasz_loss,
    "binary_accuracy": binary_accuracy,
}


def parse_composite(expr: str) -> List[Tuple[float, str]]:
    """`binary_crossentropy+0.1*dice_loss` (README.md:210-214) -> [(1.0,'binary_crossentropy'),(0.1,'dice_loss')]."""
    out = []
    for term in expr.replace(" ", "").split("+"):
        if "*" in term:
            a, b = term.split("*")
            try:
                out.append((float(a), b))
            except ValueError:
                out.append((float(b), a))
        else:
            out.append((1.0, term))
    return out


def composite(expr: str) -> Callable:
    terms = parse_composite(expr)

    def fn(y_true, y_pred):
        return sum(w * REGISTRY[n](y_true, y_pred) for w, n in terms)

    return fn
