"""K-fold x stage x epoch training loop behind PipelineConfig.fit (reference: inherited
musket_core.generic_config.GenericImageTaskConfig.fit -> Stage.execute -> keras fit_generator [DEP]; behaviour per
README.md:116-205: shuffled K folds with a fixed seed, per-stage epochs/lr/loss overrides, metrics/metrics-<fold>.<stage>.csv,
weights/best-<fold>.<stage>.weights chosen by primary_metric, summary.yaml when done).

Per training step the host only stacks the raw uint8 batch into pinned memory; augmentation, forward, loss, backward
and the optimizer are one CUDA-graph replay of libstp kernels (trainer.Trainer.step_from_host)."""
from __future__ import annotations

import csv
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import yaml

from . import callbacks as _cb
from .segmentation import _METRIC_ALIASES, parse_augmentation, parse_loss


def _resize_pair(x, y, shape):
    import cv2
    H, W = int(shape[0]), int(shape[1])
    if x.shape[0] != H or x.shape[1] != W:
        x = cv2.resize(x, (W, H), interpolation=cv2.INTER_CUBIC)   # imgaug Resize default for images
        y = cv2.resize(y, (W, H), interpolation=cv2.INTER_NEAREST)
    if y.ndim == 2:
        y = y[:, :, None]
    return np.ascontiguousarray(x, dtype=np.uint8), np.ascontiguousarray(y, dtype=np.uint8)


def _stack(ds, ids: Sequence[int], shape, out_img, out_mask):
    for j, i in enumerate(ids):
        it = ds[int(i)]
        x, y = _resize_pair(np.asarray(it.x), np.asarray(it.y), shape)
        out_img[j] = torch_from(x)
        out_mask[j] = torch_from(y)


def torch_from(a):
    import torch
    return torch.from_numpy(a)


def _better(mode: str, name: str, new: float, best: Optional[float]) -> bool:
    if best is None:
        return True
    if mode == "auto":
        mode = "min" if "loss" in name else "max"
    return new < best if mode == "min" else new > best


def run_fit(cfg, ds, subsample=1.0, foldsToExecute=None, start_from_stage=0):
    import torch
    from .trainer import Trainer

    base = cfg._dir()
    summary = os.path.join(base, "summary.yaml")
    if os.path.exists(summary) and not cfg.allowResume:
        raise ValueError("Experiment is already finished!")
    for d in ("weights", "metrics"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    n_all = len(ds)
    rng = np.random.default_rng(cfg.random_state)
    all_idx = np.arange(n_all)
    if cfg.testSplit > 0:
        perm = rng.permutation(n_all)
        n_test = int(round(n_all * cfg.testSplit))
        all_idx = np.sort(perm[n_test:])
    folds = cfg.kfold(len(all_idx))
    B, shape = cfg.batch, cfg.shape
    metric_names = [m for m in (_METRIC_ALIASES.get(m) for m in cfg.metrics) if m]
    results = []
    for fi, (tr, va) in enumerate(folds):
        if foldsToExecute is not None and fi not in foldsToExecute:
            continue
        tr_idx, va_idx = all_idx[tr], all_idx[va]
        if subsample < 1.0:
            tr_idx = tr_idx[: max(1, int(len(tr_idx) * subsample))]
        net = cfg.createNet()
        # two pinned host batches: the loader fills one while the previous one is still being copied / trained on
        himgs = [torch.zeros((B, shape[0], shape[1], shape[2]), dtype=torch.uint8).pin_memory() for _ in range(2)]
        hmasks = [torch.zeros((B, shape[0], shape[1], cfg.classes), dtype=torch.uint8).pin_memory() for _ in range(2)]
        himg, hmask = himgs[0], hmasks[0]
        for si, stage in enumerate(cfg.stages):
            if si < start_from_stage:
                continue
            net.loss.set_weights(*parse_loss(stage.get("loss", cfg.loss)))
            tr_ = Trainer(net, optimizer=cfg.optimizer, lr=stage.get("lr", cfg.lr), clipnorm=cfg.clipnorm,
                          clipvalue=cfg.clipvalue,
                          augment=parse_augmentation(cfg.augmentation, seed=cfg.random_state + 1000 * fi + si))
            tr_.enable_host_feed()
            # stage `callbacks:` replaces the config-level block, `extra_callbacks:` adds to it (StageConfig, segmentation.raml:124-136)
            cbs = _cb.build(stage.get("callbacks", cfg.callbacks), stage.get("extra_callbacks"))
            for cb in cbs:
                cb.on_train_begin(tr_)
            iteration = 0
            mpath = os.path.join(base, "metrics", "metrics-%d.%d.csv" % (fi, si))
            fields = ["epoch", "loss"] + metric_names + ["val_loss"] + ["val_" + m for m in metric_names] + ["lr"]
            rows: List[Dict] = []
            best = None
            pm = cfg.primary_metric
            for epoch in range(int(stage.get("epochs", 1))):
                order = rng.permutation(tr_idx)
                steps = max(1, len(order) // B)
                agg: Dict[str, float] = {}
                def _acc(m):
                    for k, v in (m or {}).items():
                        agg[k] = agg.get(k, 0.0) + v / steps          # Keras progress-bar averaging: equal weight per batch

                for s in range(steps):
                    ids = [order[(s * B + j) % len(order)] for j in range(B)]
                    _stack(ds, ids, shape, himgs[s & 1], hmasks[s & 1])
                    for cb in cbs:
                        cb.on_batch_begin(tr_, iteration)
                    iteration += 1
                    _acc(tr_.step_from_host_pipelined(himgs[s & 1], hmasks[s & 1]))   # metrics of the previous step
                _acc(tr_.flush_host_pipeline())
                val = evaluate(net, tr_, ds, va_idx, shape, himg, hmask)
                row = {"epoch": epoch, "loss": agg.get("loss", float("nan")), "lr": tr_.get_lr()}
                for mname in metric_names:
                    row[mname] = agg.get(mname, float("nan"))
                row["val_loss"] = val.get("loss", float("nan"))
                for mname in metric_names:
                    row["val_" + mname] = val.get(mname, float("nan"))
                rows.append(row)
                with open(mpath, "w", newline="") as f:
                    w = csv.DictWriter(f, fieldnames=fields)
                    w.writeheader()
                    w.writerows(rows)
                key = pm if pm in row else "val_loss"
                if _better(cfg.primary_metric_mode, key, row[key], best):
                    best = row[key]
                    np.savez(os.path.join(base, "weights", "best-%d.%d.weights.npz" % (fi, si)), **net.get_weights())
                for cb in cbs:
                    cb.on_epoch_end(tr_, epoch, row)
                if any(cb.stop_training for cb in cbs):
                    break
            results.append({"fold": fi, "stage": si, "best_" + pm: None if best is None else float(best), "epochs": len(rows)})
    with open(summary, "w") as f:
        yaml.safe_dump({"completed": True, "folds": len(folds), "results": results,
                        "finished_at": time.strftime("%Y-%m-%d %H:%M:%S")}, f)
    return results


def evaluate(net, trainer, ds, idx, shape, himg, hmask) -> Dict[str, float]:
    """Validation pass: inference-mode BatchNorm (moving statistics), no augmentation, metrics averaged over batches.
    A final partial batch is padded by wrapping around (the engine's graph has a fixed batch size)."""
    import torch
    if len(idx) == 0:
        return {}
    B = net.batch
    steps = (len(idx) + B - 1) // B
    agg: Dict[str, float] = {}
    net.training = False
    try:
        for s in range(steps):
            ids = [idx[(s * B + j) % len(idx)] for j in range(B)]
            _stack(ds, ids, shape, himg, hmask)
            trainer.set_batch(himg.to(net.device, non_blocking=True), hmask.to(net.device, non_blocking=True))
            net.prep_weights()
            net.forward()
            for k, v in trainer.metrics().items():
                agg[k] = agg.get(k, 0.0) + v / steps
    finally:
        net.training = True
    return agg
