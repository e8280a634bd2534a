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


class _Concat:
    """`extra_train_data:` (reference FAQ.md:50-61, segmentation.py:29 `extra_train`): a registered dataset whose items are
    appended to the TRAINING side of every fold (never to validation)."""

    def __init__(self, a, b):
        self.a, self.b = a, b

    def __len__(self):
        return len(self.a) + len(self.b)

    def __getitem__(self, i):
        i = int(i)
        return self.a[i] if i < len(self.a) else self.b[i - len(self.a)]

    def isPositive(self, i):
        i = int(i)
        d, j = (self.a, i) if i < len(self.a) else (self.b, i - len(self.a))
        return d.isPositive(j) if hasattr(d, "isPositive") else bool(np.asarray(d[j].y).any())


def _select_negatives(ds, idx, mode, rng):
    """Stage keys `negatives` / `validation_negatives` (README.md:385-415): none = positives only, real = everything,
    integer N = N negative examples per positive one (drawn without replacement)."""
    if mode in (None, "real", "all"):
        return idx
    pos_fn = ds.isPositive if hasattr(ds, "isPositive") else (lambda i: bool(np.asarray(ds[int(i)].y).any()))
    flags = np.array([bool(pos_fn(int(i))) for i in idx])
    pos, neg = idx[flags], idx[~flags]
    if mode == "none":
        return pos
    n = int(mode) * len(pos)
    take = neg if n >= len(neg) else rng.choice(neg, size=n, replace=False)
    return np.sort(np.concatenate([pos, take]))


def _weights_file(base, spec):
    """`initial_weights: ./exp/weights/best-0.1.weights` (README.md:382): path relative to the config, .npz appended if needed."""
    p = spec if os.path.isabs(spec) else os.path.join(base, spec)
    for cand in (p, p + ".npz"):
        if os.path.exists(cand):
            return cand
    raise FileNotFoundError("initial_weights: " + p)


def run_fit(cfg, ds, subsample=1.0, foldsToExecute=None, start_from_stage=0):
    import torch
    from . import ddp
    from .trainer import Trainer

    # `cfg.gpus = N` (reference FAQ.md:108-112 -> keras multi_gpu_model [DEP]) is one process per GPU here: launch the same
    # script with `python -m torch.distributed.run --nproc-per-node N ...`; every rank trains on its shard of each global
    # batch, gradients are all-reduced (ddp.py), rank 0 writes weights / metrics / summary.
    rank, local_rank, world = ddp.env_world()
    if world > 1:
        torch.cuda.set_device(local_rank)
        cfg.device = "cuda:%d" % local_rank
        ddp.init(device=torch.device(cfg.device))
    elif int(getattr(cfg, "gpus", 1) or 1) > 1:
        raise NotImplementedError("cfg.gpus = %d: data-parallel training runs as one process per GPU -- launch this script with "
                                  "`python -m torch.distributed.run --nproc-per-node %d <script>`" % (cfg.gpus, cfg.gpus))
    is_main = rank == 0

    base = cfg._dir()
    summary = os.path.join(base, "summary.yaml")
    if os.path.exists(summary) and not cfg.allowResume:
        raise ValueError("Experiment is already finished!")
    for d in ("weights", "metrics"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    ddp.barrier()   # every rank has seen the same "finished / not finished" state before anything is written
    n_all = len(ds)
    extra_name = cfg.extra.get("extra_train_data")
    extra_idx = np.zeros(0, dtype=np.int64)
    if extra_name:
        from . import segmentation as _seg
        if extra_name not in _seg.extra_train:
            raise ValueError("extra_train_data '%s' is not registered in segmentation.extra_train" % extra_name)
        extra = _seg.extra_train[extra_name]
        extra_idx = np.arange(n_all, n_all + len(extra))
        ds = _Concat(ds, extra)
    cells = None
    if getattr(cfg, "crops", 0) and cfg.crops > 1:
        from .crops import CellDataSet
        n_img = len(ds)                      # folds are split over IMAGES; every image contributes its N x N cells
        cells = CellDataSet(ds, cfg.crops)
        ds = cells
        extra_idx = cells.expand(extra_idx)
    rng = np.random.default_rng(cfg.random_state)
    all_idx = np.arange(n_all)
    if cfg.testSplit > 0:
        perm = rng.permutation(n_all)
        n_test = int(round(n_all * cfg.testSplit))
        all_idx = np.sort(perm[n_test:])
    folds = cfg.kfold(len(all_idx))
    B, shape = cfg.batch, cfg.net_shape()   # `crops: N`: the network (and every host buffer) has the CELL shape
    metric_names = [m for m in (_METRIC_ALIASES.get(m) for m in cfg.metrics) if m]
    results = []
    for fi, (tr, va) in enumerate(folds):
        if foldsToExecute is not None and fi not in foldsToExecute:
            continue
        tr_idx, va_idx = all_idx[tr], all_idx[va]
        if cells is not None:
            tr_idx, va_idx = cells.expand(tr_idx), cells.expand(va_idx)
        if subsample < 1.0:
            tr_idx = tr_idx[: max(1, int(len(tr_idx) * subsample))]
        tr_idx = np.concatenate([tr_idx, extra_idx]).astype(np.int64)
        net = cfg.createNet()
        frozen = bool(cfg.freeze_encoder)
        # two pinned host batches: the loader fills one while the previous one is still being copied / trained on
        himgs = [torch.zeros((B, shape[0], shape[1], shape[2]), dtype=torch.uint8).pin_memory() for _ in range(2)]
        hmasks = [torch.zeros((B, shape[0], shape[1], cfg.classes), dtype=torch.uint8).pin_memory() for _ in range(2)]
        himg, hmask = himgs[0], hmasks[0]
        # decode + resize run in a thread pool a couple of batches ahead of the device (loader.py); `loader_workers: 0` in the
        # config keeps everything on the calling thread (datasets whose __getitem__ is not thread safe)
        from .loader import HostLoader
        # `device_resize: true`: workers only decode; the resize to `shape` (cv2 arithmetic) runs on the device (loader.RawBatch)
        loader = HostLoader(ds, shape, cfg.classes, B, workers=int(cfg.extra.get("loader_workers", 4)),
                            device_resize=bool(cfg.extra.get("device_resize", False)))
        for si, stage in enumerate(cfg.stages):
            # stage keys that change the encoder's trainability apply whether or not the stage is executed; an explicit
            # `unfreeze_encoder: false` re-freezes (the reference's Stage sets trainability from the key's value)
            if stage.get("freeze_encoder") is not None:
                frozen = bool(stage["freeze_encoder"])
            if stage.get("unfreeze_encoder") is not None:
                frozen = not bool(stage["unfreeze_encoder"])
            wpath = os.path.join(base, "weights", "best-%d.%d.weights.npz" % (fi, si))
            mpath = os.path.join(base, "metrics", "metrics-%d.%d.csv" % (fi, si))
            dpath = os.path.join(base, "metrics", "metrics-%d.%d.done" % (fi, si))
            if si < start_from_stage:
                # fit(start_from_stage=k) (musket_core generic_config skip_stage [DEP]): a skipped stage contributes its best
                # weights (when they exist) and its freeze / unfreeze keys; it is not trained
                if os.path.exists(wpath):
                    net.set_weights(dict(np.load(wpath)))
                continue
            if cfg.allowResume and os.path.exists(wpath) and os.path.exists(mpath):
                # setAllowResume(True) (FAQ.md:3-12): a (fold, stage) that ran to its end -- all epochs, or stopped early by a
                # callback: the `.done` marker written when the epoch loop exits -- is not re-run; its best weights seed the
                # next stage.  (Logs written before the marker existed count as finished when they hold every epoch.)
                done = list(csv.DictReader(open(mpath)))
                if os.path.exists(dpath) or len(done) >= int(stage.get("epochs", 1)):
                    net.set_weights(dict(np.load(wpath)))
                    vals = [float(r[cfg.primary_metric]) for r in done if cfg.primary_metric in r]
                    mode = cfg.primary_metric_mode if cfg.primary_metric_mode != "auto" else ("min" if "loss" in cfg.primary_metric else "max")
                    best_done = (min(vals) if mode == "min" else max(vals)) if vals else None
                    results.append({"fold": fi, "stage": si, "best_" + cfg.primary_metric: best_done, "epochs": len(done),
                                    "resumed": True})
                    continue
            if is_main and os.path.exists(dpath):
                os.remove(dpath)
            if stage.get("initial_weights"):
                net.set_weights(dict(np.load(_weights_file(base, stage["initial_weights"]))), strict=False)
            net.loss.set_weights(*parse_loss(stage.get("loss", cfg.loss)))
            st_tr = _select_negatives(ds, tr_idx, stage.get("negatives"), rng)
            st_va = _select_negatives(ds, va_idx, stage.get("validation_negatives"), rng)
            if len(st_tr) == 0:
                raise ValueError("stage %d: no training samples left after `negatives: %s`" % (si, stage.get("negatives")))
            tr_ = Trainer(net, optimizer=cfg.optimizer, lr=stage.get("lr", cfg.lr), clipnorm=cfg.clipnorm,
                          clipvalue=cfg.clipvalue, freeze_encoder=frozen, world_size=world,
                          augment=parse_augmentation(cfg.augmentation, seed=cfg.random_state + 1000 * fi + si + 7919 * rank))
            if world > 1:
                ddp.broadcast_(net.flat_p, 0)
            tr_.enable_host_feed()
            # stage `callbacks:` replaces the config-level block, `extra_callbacks:` adds to it (StageConfig, segmentation.raml:124-136)
            cbs = _cb.build(stage.get("callbacks", cfg.callbacks), stage.get("extra_callbacks"))
            tr_.steps_per_epoch = max(1, len(st_tr) // (B * world))
            for cb in cbs:
                cb.on_train_begin(tr_)
            iteration = 0
            fields = ["epoch", "loss"] + metric_names + ["val_loss"] + ["val_" + m for m in metric_names] + ["lr"]
            rows: List[Dict] = []
            best = None
            pm = cfg.primary_metric
            for epoch in range(int(stage.get("epochs", 1))):
                order = rng.permutation(st_tr)            # same seed on every rank -> same permutation
                if world > 1:
                    mine = ddp.shard_indices(order, rank, world, B)
                    order = mine if len(mine) >= B else order   # too few samples for one global batch: no sharding
                steps = max(1, len(order) // B)
                agg: Dict[str, float] = {}
                def _acc(m):
                    for k, v in (m or {}).items():
                        agg[k] = agg.get(k, 0.0) + v / steps          # Keras progress-bar averaging: equal weight per batch

                batches = [[order[(s * B + j) % len(order)] for j in range(B)] for s in range(steps)]
                for bimg, bmask in loader.iterate(batches):
                    for cb in cbs:
                        cb.on_batch_begin(tr_, iteration)
                    iteration += 1
                    _acc(tr_.step_from_host_pipelined(bimg, bmask))   # metrics of the previous step
                _acc(tr_.flush_host_pipeline())
                if world > 1:   # epoch metrics = mean over ranks; BatchNorm moving statistics averaged before validation
                    keys = sorted(agg)
                    t = torch.tensor([agg[k] for k in keys], dtype=torch.float64, device=net.device)
                    ddp.mean_over_ranks_(t)
                    agg = {k: float(v) for k, v in zip(keys, t.cpu())}
                    ddp.sync_buffers_mean_(net.buffers)
                val = evaluate(net, tr_, ds, st_va, shape, himg, hmask)
                row = {"epoch": epoch, "loss": agg.get("loss", float("nan")), "lr": tr_.get_lr()}
                for mname in metric_names:
                    row[mname] = agg.get(mname, float("nan"))
                row["val_loss"] = val.get("loss", float("nan"))
                for mname in metric_names:
                    row["val_" + mname] = val.get(mname, float("nan"))
                rows.append(row)
                if is_main:
                    with open(mpath, "w", newline="") as f:
                        w = csv.DictWriter(f, fieldnames=fields)
                        w.writeheader()
                        w.writerows(rows)
                key = pm if pm in row else "val_loss"
                if _better(cfg.primary_metric_mode, key, row[key], best):
                    best = row[key]
                    if is_main:
                        np.savez(wpath, **net.get_weights())
                ddp.barrier()   # files of this epoch are complete before any rank may read them (initial_weights, resume)
                for cb in cbs:
                    cb.on_epoch_end(tr_, epoch, row)
                if any(cb.stop_training for cb in cbs):
                    break
            if is_main:
                with open(dpath, "w") as f:   # the stage ran to its end (all epochs or an early stop): resume skips it
                    f.write("epochs: %d\n" % len(rows))
            results.append({"fold": fi, "stage": si, "best_" + pm: None if best is None else float(best), "epochs": len(rows)})
        loader.close()
    if is_main:
        with open(summary, "w") as f:
            yaml.safe_dump({"completed": True, "folds": len(folds), "results": results, "world_size": world,
                            "finished_at": time.strftime("%Y-%m-%d %H:%M:%S")}, f)
    ddp.barrier()
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


class LRFinder:
    """Result of cfg.lr_find (README.md:455-470; Pavel Surmenok's keras_lr_finder as musket wraps it): the learning rate is
    multiplied by a constant factor every batch from start_lr to end_lr; the sweep stops early once the loss explodes."""

    def __init__(self):
        self.lrs: List[float] = []
        self.losses: List[float] = []

    def best_lr(self, sma=1, n_skip_beginning=2, n_skip_end=1) -> float:
        """learning rate at the steepest loss decrease (what plot_loss_change lets a user read off)."""
        d = self.derivatives(sma)
        lo, hi = n_skip_beginning, len(d) - n_skip_end
        if hi <= lo:
            return self.lrs[len(self.lrs) // 2]
        return self.lrs[lo + int(np.argmin(d[lo:hi]))]

    def derivatives(self, sma=1):
        l = np.asarray(self.losses)
        d = np.zeros_like(l)
        d[sma:] = (l[sma:] - l[:-sma]) / sma
        return d

    def plot_loss(self, n_skip_beginning=10, n_skip_end=5):
        import matplotlib.pyplot as plt
        plt.ylabel("loss"); plt.xlabel("learning rate (log scale)")
        plt.plot(self.lrs[n_skip_beginning:-n_skip_end], self.losses[n_skip_beginning:-n_skip_end]); plt.xscale("log")

    def plot_loss_change(self, sma=1, n_skip_beginning=10, n_skip_end=5, y_lim=(-0.01, 0.01)):
        import matplotlib.pyplot as plt
        d = self.derivatives(sma)
        plt.ylabel("rate of loss change"); plt.xlabel("learning rate (log scale)")
        plt.plot(self.lrs[n_skip_beginning:-n_skip_end], d[n_skip_beginning:-n_skip_end]); plt.xscale("log"); plt.ylim(y_lim)


def run_lr_find(cfg, ds, start_lr=1e-5, end_lr=1.0, epochs=1, stage=0) -> LRFinder:
    import torch
    from .trainer import Trainer

    B, shape = cfg.batch, cfg.net_shape()
    net = cfg.createNet()
    st = cfg.stages[stage] if cfg.stages else {}
    net.loss.set_weights(*parse_loss(st.get("loss", cfg.loss)))
    tr_ = Trainer(net, optimizer=cfg.optimizer, lr=start_lr, clipnorm=cfg.clipnorm, clipvalue=cfg.clipvalue,
                  freeze_encoder=bool(cfg.freeze_encoder), augment=parse_augmentation(cfg.augmentation, seed=cfg.random_state))
    tr_.enable_host_feed()
    himg = torch.zeros((B, shape[0], shape[1], shape[2]), dtype=torch.uint8).pin_memory()
    hmask = torch.zeros((B, shape[0], shape[1], cfg.classes), dtype=torch.uint8).pin_memory()
    rng = np.random.default_rng(cfg.random_state)
    steps = max(1, len(ds) // B)
    total = max(2, steps * int(epochs))
    factor = (float(end_lr) / float(start_lr)) ** (1.0 / (total - 1))
    out, lr, best = LRFinder(), float(start_lr), None
    for ep in range(int(epochs)):
        order = rng.permutation(len(ds))
        for s in range(steps):
            ids = [order[(s * B + j) % len(order)] for j in range(B)]
            _stack(ds, ids, shape, himg, hmask)
            tr_.set_lr(lr)
            m = tr_.step_from_host(himg, hmask)
            loss = m["loss"]
            out.lrs.append(lr)
            out.losses.append(loss)
            if not np.isfinite(loss) or (best is not None and loss > 4.0 * best):
                return out
            best = loss if best is None else min(best, loss)
            lr *= factor
    return out
