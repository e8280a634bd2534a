"""One training step = [augment] -> weight prep -> forward -> loss -> backward -> [all-reduce] -> optimizer,
all libstp kernels on one stream, captured once into a CUDA graph and replayed (no per-step host work, no
host<->device copy on the step path when the sample pool is device resident).

Replaces the reference's per-step path: imgaug BackgroundAugmenter queue -> np.array stack -> feed_dict H2D ->
TF session.run (SURVEY.md 3.2 "HOT LOOP per step").
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import lib as _lib
from .engine import _stream
from .models import SegNet


@dataclass
class AugmentConfig:
    """Device-fused subset of schemas/augmenters.raml: Fliplr, Flipud, Affine, Multiply, Add, Invert + musket's Rotate90."""
    fliplr: float = 0.0
    flipud: float = 0.0
    affine: bool = False
    scale: Tuple[float, float] = (1.0, 1.0)
    translate_x: Tuple[float, float] = (0.0, 0.0)
    translate_y: Tuple[float, float] = (0.0, 0.0)
    rotate: Tuple[float, float] = (0.0, 0.0)
    shear: Tuple[float, float] = (0.0, 0.0)
    multiply: Optional[Tuple[float, float]] = None
    add: Optional[Tuple[int, int]] = None
    mul_rint: bool = False
    seed: int = 0
    rot90: bool = False
    invert: float = 0.0
    color_order: Tuple[int, int, int] = (0, 1, 2)   # 0 Multiply, 1 Add, 2 Invert, in YAML order
    flip_before_rot90: int = 0   # bit 0: Fliplr precedes Rotate90 in the YAML block, bit 1: Flipud does (stp.h)
    # leading crop / pad augmenters, in YAML order: (kind, ranged, a, b, c, d) per include/stp.h stp_croppad_op
    crop_pad: Tuple[Tuple[int, int, float, float, float, float], ...] = ()
    # pixel-wise colour stage in YAML order when an extended augmenter (AddElementwise, MultiplyElementwise, Dropout,
    # AdditiveGaussianNoise, Grayscale) or a OneOf group is present: (kind, per_channel, a, b, group_id, group_size, group_member)
    # per include/stp.h stp_aug_pix_op; Multiply / Add / Invert entries use the per-sample draws configured above
    pix_ops: Tuple[Tuple[int, float, float, float, int, int, int], ...] = ()
    # neighbourhood augmenters (GaussianBlur, AverageBlur, MedianBlur, Sharpen, Emboss, EdgeDetect) interleaved with the pixel-wise
    # ones: the colour block in YAML order as ("pix", op) / ("nb", (kind, a, b, c, d, group_id, group_size, group_member)) entries;
    # empty = the block is `pix_ops` alone.  Position in this sequence = the op's Philox call index.
    colour_seq: Tuple[Tuple[str, tuple], ...] = ()

    def enabled(self) -> bool:
        return bool(self.fliplr or self.flipud or self.affine or self.multiply or self.add or self.rot90 or self.invert or
                    self.crop_pad or self.pix_ops or self.colour_seq)

    def pix_c(self, ops=None, k_base=0) -> _lib.AugPixSpec:
        ops = self.pix_ops if ops is None else ops
        spec = _lib.AugPixSpec()
        spec.n_ops, spec.mul_rint, spec.k_base = len(ops), int(self.mul_rint), int(k_base)
        for i, (kind, pc, a, b, gid, gsz, gm) in enumerate(ops):
            spec.ops[i] = _lib.AugPixOp(int(kind), float(pc), float(a), float(b), int(gid), int(gsz), int(gm))
        return spec

    def colour_runs(self):
        """the colour block as launches: ("pix", k_base, [ops]) runs of consecutive pixel-wise ops and ("nb", k_index, op) entries"""
        seq = self.colour_seq or tuple(("pix", op) for op in self.pix_ops)
        runs = []
        for k, (typ, op) in enumerate(seq):
            if typ == "pix" and runs and runs[-1][0] == "pix":
                runs[-1][2].append(op)
            elif typ == "pix":
                runs.append(("pix", k, [op]))
            else:
                runs.append(("nb", k, op))
        return runs

    def croppad_c(self) -> _lib.CropPadSpec:
        spec = _lib.CropPadSpec()
        spec.n_ops = len(self.crop_pad)
        for i, (kind, ranged, a, b, c, d) in enumerate(self.crop_pad):
            spec.ops[i] = _lib.CropPadOp(int(kind), int(ranged), float(a), float(b), float(c), float(d))
        return spec

    def to_c(self) -> _lib.AugSpec:
        m = self.multiply or (1.0, 1.0)
        a = self.add or (0, 0)
        spec = _lib.AugSpec(self.fliplr, self.flipud, int(self.affine), self.scale[0], self.scale[1],
                            self.translate_x[0], self.translate_x[1], self.translate_y[0], self.translate_y[1],
                            self.rotate[0], self.rotate[1], self.shear[0], self.shear[1],
                            int(self.multiply is not None), m[0], m[1], int(self.add is not None), int(a[0]), int(a[1]),
                            int(self.mul_rint), int(self.rot90), float(self.invert))
        for i, o in enumerate(self.color_order):
            spec.color_order[i] = int(o)
        spec.flip_before_rot90 = int(self.flip_before_rot90)
        return spec


class Trainer:
    def __init__(self, net: SegNet, optimizer="Adam", lr=None, beta_1=0.9, beta_2=0.999, epsilon=1e-7, momentum=0.0,
                 nesterov=False, rho=0.9, clipnorm=None, clipvalue=None, augment: Optional[AugmentConfig] = None,
                 world_size=1, process_group=None, freeze_encoder=False):
        self.net, self.L = net, net.L
        self.opt = (optimizer or "Adam").lower()
        if self.opt not in ("adam", "sgd", "rmsprop", "nadam"):
            raise ValueError("unknown optimizer " + str(optimizer))
        # freeze_encoder (README.md:281-304): the encoder parameters come first in the flat buffers, so freezing them is an
        # offset into the one optimizer launch (their gradients are still computed; nothing reads them)
        self.opt_off = net.encoder_floats() if freeze_encoder else 0
        self.lr = float(lr) if lr is not None else (0.01 if self.opt == "sgd" else 0.002 if self.opt == "nadam" else 1e-3)
        self.b1, self.b2, self.eps, self.mu, self.nesterov, self.rho = beta_1, beta_2, epsilon, momentum, nesterov, rho
        self.clipnorm, self.clipvalue = clipnorm or 0.0, clipvalue or 0.0
        dev = net.device
        self.m = torch.zeros(net.n_flat, dtype=torch.float32, device=dev)
        self.v = torch.zeros(net.n_flat, dtype=torch.float32, device=dev) if self.opt in ("adam", "nadam") else None
        self.nadam_sched = torch.tensor([1.0, 0.0, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        # optimizer iteration counter (Keras `iterations`): belongs to THIS optimizer, starts at 0 with its zeroed moments -- the
        # reference recompiles the model per stage, so bias correction restarts with every stage.  net.d_step keeps counting
        # across stages: it only keys the augmentation RNG stream / the pool cursor.
        self.d_opt_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.sumsq_partial = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.world_size, self.pg = world_size, process_group
        self.overlap_allreduce = True   # bucketed all-reduce overlapping the early layers' backward (see _capture_ddp)
        self.graph_bwd2: Optional[torch.cuda.CUDAGraph] = None
        self.augment = augment if (augment is not None and augment.enabled()) else None
        self.pool_img: Optional[torch.Tensor] = None
        self.pool_mask: Optional[torch.Tensor] = None
        self.aug_params = torch.zeros(net.batch * C.sizeof(_lib.AugSample), dtype=torch.uint8, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_opt: Optional[torch.cuda.CUDAGraph] = None
        self.lr_scale = torch.ones(1, dtype=torch.float32, device=dev)   # device-side lr multiplier (schedules)
        self._lr_scale_host = torch.ones(1, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.ones(1)
        self._gx = _lib.GradXform(1.0 / world_size, self.clipnorm, self.clipvalue,
                                  self.sumsq.data_ptr() if self.clipnorm > 0 else None, self.lr_scale.data_ptr())
        self._ident = AugmentConfig()
        self._cp_img = self._cp_mask = self._cp_items = None   # staging of the crop / pad augmenter stage (run_augment)
        self._nb_img = self._nb_work = self._nb_table = None   # scratch batch / tap tables of the neighbourhood augmenters

    # ---- learning-rate schedules ----------------------------------------------------------------
    def set_lr(self, lr: float):
        """Change the learning rate used by subsequent steps (also replays of the captured graph): the optimizer kernel
        multiplies its captured base rate by a device scalar."""
        self._lr_scale_host[0] = float(lr) / self.lr
        self.lr_scale.copy_(self._lr_scale_host, non_blocking=True)
        self.current_lr = float(lr)

    def get_lr(self) -> float:
        return getattr(self, "current_lr", self.lr)

    # ---- data ---------------------------------------------------------------------------------
    def set_pool(self, images: torch.Tensor, masks: torch.Tensor):
        """Device-resident sample pool uint8 [P,H,W,C] / [P,H,W,classes]; step i draws samples (i*B+j) % P."""
        H, W, CI = self.net.input_shape
        assert images.dtype == torch.uint8 and masks.dtype == torch.uint8
        assert tuple(images.shape[1:]) == (H, W, CI) and tuple(masks.shape[1:]) == (H, W, self.net.classes)
        self.pool_img = images.to(self.net.device).contiguous()
        self.pool_mask = masks.to(self.net.device).contiguous()

    def set_batch(self, images: torch.Tensor, masks: torch.Tensor, non_blocking=True):
        """Copy one (already augmented or raw) batch into the network's input buffers."""
        self.net.img.storage.copy_(images.reshape(-1), non_blocking=non_blocking)
        self.net.mask.storage.copy_(masks.reshape(-1), non_blocking=non_blocking)

    # ---- pieces -------------------------------------------------------------------------------
    def run_augment(self):
        """pool -> (img, mask) input buffers through the fused K1 kernel (identity spec = plain gather)."""
        net, st = self.net, _stream()
        H, W, CI = net.input_shape
        cfg = self.augment or self._ident
        spec = cfg.to_c()
        src_img, src_mask, pool_n = self.pool_img, self.pool_mask, self.pool_img.shape[0]
        if cfg.crop_pad:
            # leading Pad / PadToFixedSize / CropToFixedSize / CropAndPad: one window per sample (drawn on the device), brought
            # back to `shape` by the cv2-arithmetic resize (cubic for the image, nearest for the mask) into a staging batch that
            # the fused flip / affine / colour kernel then reads in batch order
            if self._cp_img is None:
                dev = net.device
                self._cp_img = torch.zeros((net.batch, H, W, CI), dtype=torch.uint8, device=dev)
                self._cp_mask = torch.zeros((net.batch, H, W, net.classes), dtype=torch.uint8, device=dev)
                self._cp_items = torch.zeros(2 * net.batch * C.sizeof(_lib.ResizeItem), dtype=torch.uint8, device=dev)
            cps = cfg.croppad_c()
            it_img = self._cp_items.data_ptr()
            it_mask = it_img + net.batch * C.sizeof(_lib.ResizeItem)
            self.L.croppad_draw(C.byref(cps), cfg.seed, net.d_step.data_ptr(), net.batch, pool_n, H, W, CI, net.classes, it_img, it_mask, st)
            self.L.resize_u8(self.pool_img.data_ptr(), it_img, net.batch, CI, self._cp_img.data_ptr(), H, W, _lib.RESIZE_CUBIC, st)
            self.L.resize_u8(self.pool_mask.data_ptr(), it_mask, net.batch, net.classes, self._cp_mask.data_ptr(), H, W,
                             _lib.RESIZE_NEAREST, st)
            src_img, src_mask, pool_n = self._cp_img, self._cp_mask, net.batch   # sample i of the staging batch is batch item i
        self.L.augment_draw(C.byref(spec), cfg.seed, net.d_step.data_ptr(), net.batch, pool_n, H, W,
                            self.aug_params.data_ptr(), st)
        runs = cfg.colour_runs() if (cfg.pix_ops or cfg.colour_seq) else []
        n_nb = sum(1 for r in runs if r[0] == "nb")
        if n_nb and self._nb_img is None:
            self._nb_img = torch.zeros(net.batch * H * W * CI, dtype=torch.uint8, device=net.device)
            self._nb_work = torch.zeros(max(int(self.L.augment_neighbourhood_workspace(net.batch)), 16), dtype=torch.uint8,
                                        device=net.device)
        # neighbourhood ops (blur / sharpen / emboss / edge-detect) ping-pong between the network's image buffer and a scratch
        # batch; the gather kernel starts in whichever of the two makes the LAST op land in the network's buffer
        cur = self._nb_img if n_nb % 2 else net.img.storage
        self.L.augment_apply(src_img.data_ptr(), src_mask.data_ptr(), self.aug_params.data_ptr(),
                             cur.data_ptr(), net.mask.storage.data_ptr(), net.batch, H, W, CI, net.classes,
                             int(cfg.mul_rint) | (2 if runs else 0), st)
        for typ, k, payload in runs:   # the colour stage in YAML order (incl. Multiply / Add / Invert), pixel-wise runs in place
            if typ == "pix":
                ps = cfg.pix_c(payload, k)
                self.L.augment_pixel_ops(cur.data_ptr(), self.aug_params.data_ptr(), C.byref(ps), cfg.seed,
                                         net.d_step.data_ptr(), net.batch, H, W, CI, st)
            else:
                other = net.img.storage if cur is self._nb_img else self._nb_img
                kind, a, b, c, d, gid, gsz, gm = payload
                table = None
                if int(kind) == _lib.NB_KINDS["DirectedEdgeDetect"]:
                    if self._nb_table is None:
                        self._nb_table = torch.from_numpy(_lib.directed_edge_table()).to(net.device).contiguous()
                    table = self._nb_table.data_ptr()
                op = _lib.AugNbOp(int(kind), float(a), float(b), float(c), float(d), int(k), int(gid), int(gsz), int(gm), table)
                self.L.augment_neighbourhood(cur.data_ptr(), other.data_ptr(), self.aug_params.data_ptr(), C.byref(op), cfg.seed,
                                             net.d_step.data_ptr(), net.batch, H, W, CI, self._nb_work.data_ptr(),
                                             self._nb_work.numel(), st)
                cur = other

    def allreduce(self):
        if self.world_size > 1:
            from . import ddp
            ddp.allreduce_sum_(self.net.flat_g, self.pg)   # the 1/world mean is folded into the optimizer kernel

    def run_optimizer(self):
        net, st = self.net, _stream()
        off, cnt = self.opt_off, net.n_flat - self.opt_off
        if cnt <= 0:
            self.L.step_advance(net.d_step.data_ptr(), st)
            self.L.step_advance(self.d_opt_step.data_ptr(), st)
            return
        p, g, m = net.flat_p.data_ptr() + 4 * off, net.flat_g.data_ptr() + 4 * off, self.m.data_ptr() + 4 * off
        v = self.v.data_ptr() + 4 * off if self.v is not None else None
        if self.clipnorm > 0:  # keras clipnorm: global norm over the TRAINABLE weights' gradients
            self.L.sumsq(g, cnt, self.sumsq_partial.data_ptr(), self.sumsq.data_ptr(), st)
        gx = C.byref(self._gx)
        if self.opt == "adam":
            self.L.adam(p, g, m, v, cnt, self.lr, self.b1, self.b2, self.eps, gx, self.d_opt_step.data_ptr(), st)
        elif self.opt == "nadam":
            self.L.nadam(p, g, m, v, self.nadam_sched.data_ptr(), cnt, self.lr, self.b1, self.b2, self.eps, 0.004, gx,
                         self.d_opt_step.data_ptr(), st)
        elif self.opt == "sgd":
            self.L.sgd(p, g, m, cnt, self.lr, self.mu, int(self.nesterov), gx, st)
        else:
            self.L.rmsprop(p, g, m, cnt, self.lr, self.rho, self.eps, gx, st)
        self.L.step_advance(net.d_step.data_ptr(), st)
        self.L.step_advance(self.d_opt_step.data_ptr(), st)

    # ---- whole step ---------------------------------------------------------------------------
    def step_eager(self, from_pool=True):
        net = self.net
        net.training = True
        if from_pool and self.pool_img is not None:
            self.run_augment()
        net.prep_weights()
        net.forward()
        net.backward()
        self.allreduce()
        self.run_optimizer()

    def step_compute(self, from_pool=True):
        net = self.net
        net.training = True
        if from_pool and self.pool_img is not None:
            self.run_augment()
        net.prep_weights()
        net.forward()
        net.backward()

    def capture(self, from_pool=True):
        """Warm up eagerly once on a side stream, then capture the step into a CUDA graph.  With world_size > 1 the
        step is TWO graphs (augment+forward+backward | optimizer) with the NCCL all-reduce enqueued between them on
        the same stream, so no collective is captured."""
        if self.world_size > 1:
            return self._capture_ddp(from_pool)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        saved = self._snapshot()
        with torch.cuda.stream(s):
            self.step_eager(from_pool)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(saved)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step_eager(from_pool)
        self._restore(saved)  # capture does not execute, but be explicit
        self.graph = g
        return g

    def _capture_ddp(self, from_pool=True):
        """Data-parallel step = three graphs with two collectives between them:
            G1  augment + forward + backward of the LATE layers (>= 90 % of the parameters: decoder + deep encoder stages)
            --  all-reduce of their gradient range, asynchronous (NCCL stream), overlapping ...
            G2  ... the backward of the early layers
            --  all-reduce of the remaining range
            G3  optimizer (waits for both collectives)."""
        net = self.net
        self._split_op, self._split_off = net.split_for_overlap(0.9) if self.overlap_allreduce else (0, 0)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        saved = self._snapshot()
        with torch.cuda.stream(s):
            self.step_eager(from_pool)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(saved)
        g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            net.training = True
            if from_pool and self.pool_img is not None:
                self.run_augment()
            net.prep_weights()
            net.forward()
            net.backward(self._split_op, None)
        if self._split_op > 0:
            with torch.cuda.graph(g2, pool=g1.pool()):
                net.backward(0, self._split_op)
        else:
            g2 = None
        with torch.cuda.graph(g3, pool=g1.pool()):
            self.run_optimizer()
        self._restore(saved)
        self.graph, self.graph_bwd2, self.graph_opt = g1, g2, g3
        return g1

    def _ddp_step(self):
        from . import ddp
        g = self.net.flat_g
        self.graph.replay()
        if self.graph_bwd2 is None:
            ddp.allreduce_sum_(g, self.pg)
        else:
            w1 = ddp.allreduce_sum_async(g[self._split_off:], self.pg)   # late layers: overlaps the early layers' backward
            self.graph_bwd2.replay()
            w2 = ddp.allreduce_sum_async(g[:self._split_off], self.pg)
            for w in (w1, w2):
                if w is not None:
                    w.wait()
        self.graph_opt.replay()

    def step(self):
        if self.graph is not None:
            if self.world_size > 1:
                self._ddp_step()
            else:
                self.graph.replay()
        else:
            self.step_eager()

    # ---- host-fed steps (the public fit() path and bench.py's e2e number) -----------------------
    def enable_host_feed(self):
        """Device staging buffers + graph capture; afterwards a host-fed step is: async H2D of the raw uint8 batch from
        pinned memory -> graph replay (augment .. optimizer) -> D2H of the 16-float result.

        step_from_host() is synchronous (returns this step's metrics).  step_from_host_pipelined() double-buffers the
        staging area on a copy stream so the H2D of step k overlaps the compute of step k-1 and returns the metrics of
        the PREVIOUS step (flush_host_pipeline() returns the last one) -- what fit() and bench.py's e2e leg drive."""
        net = self.net
        H, W, CI = net.input_shape
        dev = net.device
        self.pool_img = torch.zeros((net.batch, H, W, CI), dtype=torch.uint8, device=dev)
        self.pool_mask = torch.zeros((net.batch, H, W, net.classes), dtype=torch.uint8, device=dev)
        self._res_host = torch.zeros(16, dtype=torch.float32).pin_memory()
        self._stage_img = [torch.zeros_like(self.pool_img) for _ in range(2)]
        self._stage_mask = [torch.zeros_like(self.pool_mask) for _ in range(2)]
        self._res_ring = [torch.zeros(16, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._ev_copied = [torch.cuda.Event() for _ in range(2)]
        self._ev_free = [torch.cuda.Event() for _ in range(2)]
        self._ev_done = [torch.cuda.Event() for _ in range(2)]
        self._pipe_k = 0
        self._pipe_pending = None
        self.capture(from_pool=True)

    @staticmethod
    def _metrics_from(r) -> Dict[str, float]:
        return {"loss": float(r[_lib.L_LOSS]), "binary_crossentropy": float(r[_lib.L_BCE]), "dice": float(r[_lib.L_DICE]),
                "iou": float(r[_lib.L_IOU]), "binary_accuracy": float(r[_lib.L_ACC]), "iot": float(r[_lib.L_IOT]),
                "categorical_crossentropy": float(r[_lib.L_CCE]), "categorical_accuracy": float(r[_lib.L_CACC])}

    def step_from_host(self, images: torch.Tensor, masks: torch.Tensor, read_metrics=True):
        self.pool_img.copy_(images, non_blocking=True)
        self.pool_mask.copy_(masks, non_blocking=True)
        self.step()
        if not read_metrics:
            return None
        self._res_host.copy_(self.net.loss.result, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._metrics_from(self._res_host)

    def step_from_host_pipelined(self, images: torch.Tensor, masks: torch.Tensor):
        """Enqueue one host-fed step; returns the metrics of the previously enqueued step (None for the first call).
        `images` / `masks` (pinned host tensors) may be overwritten once the NEXT call returns."""
        b = self._pipe_k & 1
        main = torch.cuda.current_stream()
        raw = images if masks is None and hasattr(images, "arena_img") else None   # loader.RawBatch: on-device ingest
        if raw is not None:
            self._ingest_raw(raw, b, main)
        else:
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._ev_free[b])          # staging slot b was drained by step k-2
                self._stage_img[b].copy_(images, non_blocking=True)
                self._stage_mask[b].copy_(masks, non_blocking=True)
                self._ev_copied[b].record(self._copy_stream)
            main.wait_event(self._ev_copied[b])
            self.pool_img.copy_(self._stage_img[b], non_blocking=True)   # 17 MB device copy, ~5 us
            self.pool_mask.copy_(self._stage_mask[b], non_blocking=True)
            self._ev_free[b].record(main)
        self.step()
        self._res_ring[b].copy_(self.net.loss.result, non_blocking=True)
        self._ev_done[b].record(main)
        prev, self._pipe_pending = self._pipe_pending, b
        self._pipe_k += 1
        if prev is None:
            return None
        self._ev_done[prev].synchronize()
        return self._metrics_from(self._res_ring[prev])

    def _ingest_raw(self, raw, b: int, main):
        """on-device ingest: H2D of the UNRESIZED samples (loader.RawBatch: byte arenas + one
        stp_resize_item per sample) on the copy stream, then stp_resize_u8 -- cv2.resize arithmetic, cubic for images / nearest
        for masks -- writes the network-shape batch straight into the step's input pool."""
        net = self.net
        H, W, CI = net.input_shape
        if not hasattr(self, "_raw_dev"):
            self._raw_dev = [None, None]
        need = (raw.arena_img.numel(), raw.arena_mask.numel())
        cur = self._raw_dev[b]
        if cur is None or cur[0].numel() < need[0] or cur[1].numel() < need[1]:
            dev = net.device
            torch.cuda.current_stream().synchronize()   # (re)allocation of a staging arena: rare, not on the steady-state path
            cur = (torch.zeros(need[0], dtype=torch.uint8, device=dev), torch.zeros(need[1], dtype=torch.uint8, device=dev),
                   torch.zeros(raw.items_img.numel(), dtype=torch.uint8, device=dev),
                   torch.zeros(raw.items_mask.numel(), dtype=torch.uint8, device=dev))
            self._raw_dev[b] = cur
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ev_free[b])
            cur[0][:raw.used_img].copy_(raw.arena_img[:raw.used_img], non_blocking=True)
            cur[1][:raw.used_mask].copy_(raw.arena_mask[:raw.used_mask], non_blocking=True)
            cur[2].copy_(raw.items_img, non_blocking=True)
            cur[3].copy_(raw.items_mask, non_blocking=True)
            self._ev_copied[b].record(self._copy_stream)
        main.wait_event(self._ev_copied[b])
        st = main.cuda_stream
        self.L.resize_u8(cur[0].data_ptr(), cur[2].data_ptr(), raw.n, CI, self.pool_img.data_ptr(), H, W, _lib.RESIZE_CUBIC, st)
        self.L.resize_u8(cur[1].data_ptr(), cur[3].data_ptr(), raw.n, net.classes, self.pool_mask.data_ptr(), H, W, _lib.RESIZE_NEAREST, st)
        self._ev_free[b].record(main)
        self.last_h2d_bytes = raw.used_img + raw.used_mask + 2 * raw.items_img.numel()

    def flush_host_pipeline(self):
        """Metrics of the last enqueued pipelined step (waits for it)."""
        prev, self._pipe_pending = self._pipe_pending, None
        if prev is None:
            return None
        self._ev_done[prev].synchronize()
        return self._metrics_from(self._res_ring[prev])

    def loss_value(self) -> float:
        return float(self.net.loss.result[_lib.L_LOSS].item())

    def metrics(self) -> Dict[str, float]:
        return self._metrics_from(self.net.loss.result.detach().cpu().numpy())

    # ---- state --------------------------------------------------------------------------------
    def _snapshot(self):
        n = self.net
        return dict(p=n.flat_p.clone(), m=self.m.clone(), v=None if self.v is None else self.v.clone(),
                    step=n.d_step.clone(), opt_step=self.d_opt_step.clone(), sched=self.nadam_sched.clone(),
                    bufs={k: b.clone() for k, b in n.buffers.items()})

    def _restore(self, s):
        n = self.net
        n.flat_p.copy_(s["p"])
        self.m.copy_(s["m"])
        if self.v is not None:
            self.v.copy_(s["v"])
        n.d_step.copy_(s["step"])
        self.d_opt_step.copy_(s["opt_step"])
        self.nadam_sched.copy_(s["sched"])
        for k, b in n.buffers.items():
            b.copy_(s["bufs"][k])

    def state_dict(self):
        n = self.net
        return {"weights": n.get_weights(), "m": self.m.cpu().numpy(), "v": None if self.v is None else self.v.cpu().numpy(),
                "step": int(n.d_step.item()), "opt_step": int(self.d_opt_step.item())}

    def load_state_dict(self, d):
        n = self.net
        n.set_weights(d["weights"])
        self.m.copy_(torch.from_numpy(d["m"]))
        if self.v is not None and d.get("v") is not None:
            self.v.copy_(torch.from_numpy(d["v"]))
        n.d_step.fill_(int(d["step"]))
        self.d_opt_step.fill_(int(d.get("opt_step", d["step"])))
