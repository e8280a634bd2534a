"""Static-graph training engine over libstp: explicit forward / backward op lists on preallocated NHWC bf16
buffers, flat fp32 parameter / gradient / optimizer-state buffers, zero-copy concat (skip tensors are channel
slices of the decoder's concat buffers), whole-step CUDA-graph capture.

PyTorch is used for device memory, streams, CUDA graphs and torch.distributed only -- all arithmetic on the
step path is libstp kernels (no autograd, no torch ops).  Mirrors what the reference reaches through
keras Model.train_on_batch (reference segmentation.py:249-260 -> generic.Stage.execute -> fit_generator).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import lib as _lib

BF16, F32, U8 = _lib.BF16, _lib.F32, _lib.U8
_TORCH_DT = {BF16: torch.bfloat16, F32: torch.float32, U8: torch.uint8}
_ESIZE = {BF16: 2, F32: 4, U8: 1}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# Critical-path experiments ONLY (results are invalid): STP_SKIP="wgrad,bn_apply,bn_bwd_reduce,bn_bwd_apply,dgrad,conv_fwd"
# drops whole kernel categories from the step so that scripts/knockout.py can measure how much of the captured step each
# category actually costs once overlap (side-stream weight gradients, PDL) is accounted for.
_SKIP = set(filter(None, os.environ.get("STP_SKIP", "").split(",")))


class Buf:
    """NHWC device tensor, possibly a channel slice [c_off, c_off+c) of a wider root buffer (ld = root.c)."""

    def __init__(self, net: "Net", n, h, w, c, dtype=None, root: Optional["Buf"] = None, c_off=0, name=""):
        if dtype is None:   # activations / gradients: bf16, or fp32 in parity mode (Net.precision)
            dtype = root.dtype if root is not None else net.act_dtype
        self.net, self.n, self.h, self.w, self.c, self.dtype, self.name = net, n, h, w, c, dtype, name
        self.root = root.root if root is not None else self
        self.c_off = c_off if root is None else root.c_off + c_off
        if root is None:
            self.storage = torch.zeros(n * h * w * c, dtype=_TORCH_DT[dtype], device=net.device)
            self.ld = c
        else:
            self.storage = root.root.storage
            self.ld = root.root.ld
        ptr = self.storage.data_ptr() + self.c_off * _ESIZE[dtype]
        self.st = _lib.Tensor(ptr, n, h, w, c, self.ld, dtype)
        self.ref = C.byref(self.st)
        self._grad: Optional[Buf] = None
        if name:
            net.bufs[name] = self

    @property
    def rows(self):
        return self.n * self.h * self.w

    def slice(self, c_off, c, name="") -> "Buf":
        return Buf(self.net, self.n, self.h, self.w, c, self.dtype, root=self, c_off=c_off, name=name)

    def grad(self) -> "Buf":
        if self._grad is None:
            if self.root is self:
                self._grad = Buf(self.net, self.n, self.h, self.w, self.c, self.dtype, name="d_" + self.name)
            else:
                self._grad = self.root.grad().slice(self.c_off, self.c, name="d_" + self.name)
        return self._grad

    def set_grad(self, g: "Buf"):
        self._grad = g

    def torch(self) -> torch.Tensor:
        """[n,h,w,c] (strided) torch view, for tests / host access."""
        full = self.storage.view(self.n, self.h, self.w, self.ld)
        return full[..., self.c_off:self.c_off + self.c]

    def key(self):
        return (id(self.root), self.c_off, self.c)


class _NullCtx:
    def __init__(self, ws):
        self.ws = ws

    def __enter__(self):
        return self.ws

    def __exit__(self, *a):
        return False


class _SideCtx:
    def __init__(self, stream, ws):
        self.ctx, self.ws = torch.cuda.stream(stream), ws

    def __enter__(self):
        self.ctx.__enter__()
        return self.ws

    def __exit__(self, *a):
        return self.ctx.__exit__(*a)


class Param:
    def __init__(self, name, shape, kind, init):
        self.name, self.shape, self.kind, self.init = name, tuple(shape), kind, init
        self.size = int(np.prod(shape))
        self.offset = -1  # in floats, into the flat buffers
        self.trainable = True


class Op:
    acc: List[bool] = []

    def prepare(self):
        pass

    def fwd(self):
        pass

    def bwd(self):
        pass

    def grad_writes(self) -> List[Buf]:
        """grad buffers this op's bwd writes (in order), used to resolve first-writer vs accumulate."""
        return []


class Net:
    def __init__(self, batch: int, device="cuda:0", seed: int = 0, precision: str = "bf16"):
        """precision = "bf16" (product path: bf16 activations / gradients / weight copies on the tcgen05 kernels, fp32
        accumulation, fp32 master weights) or "fp32" (PARITY MODE, csrc/f32_path.cu: fp32 activations, gradients and weights,
        CUDA-core FFMA convolutions, double-precision reductions -- the mode in which the 100-step loss curve is held to the
        1e-3 of north_star; same graph, same ops, same C ABI entry points)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        self.act_dtype = F32 if precision == "fp32" else BF16
        self.L = _lib.Lib()
        self.device = torch.device(device)
        self.batch = batch
        self.ops: List[Op] = []
        self.bufs: Dict[str, "Buf"] = {}  # named activation buffers (tests / debugging)
        self.params: Dict[str, Param] = {}
        self.buffers: Dict[str, torch.Tensor] = {}  # BN moving stats (fp32)
        self.gen = np.random.default_rng(seed)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.training = True
        self.finalized = False
        self._ws_bytes = 0
        self._partial_floats = 2 * _lib.BN_MAX_PARTIALS * 8
        self.encoder_param_names: List[str] = []
        self.fuse_bn_stats = precision == "bf16"
        # BatchNorm-backward reduction inside the producing dgrad's epilogue (stp_conv_dgrad_bn): 33 launches fewer per
        # U-Net/ResNet-34 step.  Round 1 (4 epilogue warps) measured it 1.3 % slower; with the 8-warp epilogue it is 1 % FASTER
        # (round 2, same box: 8.826 vs 8.911 ms, profiles/r2_s4_bench*.json.log) -> on by default; STP_FUSE_BN_BWD=0 turns it off.
        self.fuse_bn_bwd = os.environ.get("STP_FUSE_BN_BWD", "1") == "1" and precision == "bf16"
        # 1x1 stride-1 dgrads (ResNet-50 bottlenecks, MobileNetV2 expansions) run on the same halo kernel since it serves them as
        # plain GEMMs, so their BatchNorm-backward reductions can ride its epilogue too; STP_FUSE_BN_BWD_1X1=0 turns that off.
        self.fuse_bn_bwd_1x1 = os.environ.get("STP_FUSE_BN_BWD_1X1", "1") == "1"

    # ---- parameters ---------------------------------------------------------------------------
    def add_param(self, name, shape, kind, init) -> Param:
        assert name not in self.params, name
        p = Param(name, shape, kind, init)
        self.params[name] = p
        return p

    def need_ws(self, nbytes):
        self._ws_bytes = max(self._ws_bytes, int(nbytes))

    def need_partial(self, nfloats):
        self._partial_floats = max(self._partial_floats, int(nfloats))

    def finalize(self):
        off = 0
        for p in self.params.values():
            p.offset = off
            off += (p.size + 7) // 8 * 8  # 8 elements: the bf16 copies stay 16-byte aligned
        self.n_flat = max(off, 8)
        dev = self.device
        self.flat_p = torch.zeros(self.n_flat, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.n_flat, dtype=torch.float32, device=dev)
        self.flat_wf = torch.zeros(self.n_flat, dtype=torch.bfloat16, device=dev)  # bf16 KRSC copies
        self.flat_wd = torch.zeros(self.n_flat, dtype=torch.bfloat16, device=dev)  # bf16 dgrad-layout copies
        self.ws = torch.zeros(max(self._ws_bytes, 16), dtype=torch.uint8, device=dev)
        # weight gradients run on a side stream (own workspace), overlapping the BatchNorm / dgrad chain of the layers
        # below: nothing downstream of a wgrad reads its result before the optimizer
        self.ws_side = torch.zeros(max(self._ws_bytes, 16), dtype=torch.uint8, device=dev)
        self.side_stream = torch.cuda.Stream(device=dev) if self.device.type == "cuda" else None
        self.overlap_wgrad = True
        self.partial = torch.zeros(self._partial_floats, dtype=torch.float32, device=dev)
        self.d_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sync = torch.zeros(4, dtype=torch.int32, device=dev)  # last-block tickets of the fused reduce+finalize kernels
        cmax = max([8] + [op.x.c for op in self.ops if isinstance(op, BNRelu)] +
                   [op.y.c for op in self.ops if isinstance(op, Conv) and op.b is not None])
        self.bn_acc = torch.zeros(2 * cmax, dtype=torch.float64, device=dev)  # conv-epilogue BatchNorm sums (returned to zero)
        # a BatchNorm whose input is written by exactly one conv gets its statistics from that conv's epilogue
        writers: Dict[int, List[Op]] = {}
        for op in self.ops:
            if isinstance(op, (Conv, StemConv, BNRelu, MaxPool)):
                writers.setdefault(id(op.y), []).append(op)
        for op in self.ops:
            if isinstance(op, BNRelu):
                w = writers.get(id(op.x), [])
                if len(w) == 1 and isinstance(w[0], (Conv, StemConv)) and w[0].y.dtype == BF16 and self.fuse_bn_stats:
                    w[0].bn_next = op
                    op.stats_from_conv = True
        # a BatchNorm(+ReLU) whose output gradient is written by exactly ONE op, a convolution's dgrad that covers exactly
        # that tensor, gets its backward reduction (dgamma, dbeta, bcoef) from that dgrad's epilogue (stp_conv_dgrad_bn)
        if self.fuse_bn_bwd:
            gwriters: Dict[Tuple[int, int, int], List[Op]] = {}
            roots: Dict[int, int] = {}
            for op in self.ops:
                for g in op.grad_writes():
                    gwriters.setdefault(g.key(), []).append(op)
                    roots[id(g.root)] = roots.get(id(g.root), 0) + 1
            for op in self.ops:
                if isinstance(op, BNRelu) and op.up == 1:   # (ReLU6 masks: only in the streaming 1x1 GEMM's epilogue, see below)
                    gk = op.y.grad().key()
                    w = gwriters.get(gk, [])
                    # no other op may write an overlapping slice of the same gradient buffer
                    overlap = [k for k in gwriters if k[0] == gk[0] and k != gk and k[1] < gk[1] + gk[2] and gk[1] < k[1] + k[2]]
                    if (len(w) == 1 and not overlap and type(w[0]) is Conv and w[0].needs_dgrad and w[0].x.key() == op.y.key()
                            and w[0].desc.stride == 1 and w[0].desc.up == 1 and (w[0].k > 1 or self.fuse_bn_bwd_1x1)
                            and (int(op.relu) != 2 or (w[0].k == 1 and self.fuse_bn_bwd_1x1))):
                        w[0].bnb_prev = op
                        op.reduce_from_dgrad = True
        host = np.zeros(self.n_flat, dtype=np.float32)
        for p in self.params.values():
            host[p.offset:p.offset + p.size] = p.init().reshape(-1)
        self.flat_p.copy_(torch.from_numpy(host))
        # resolve first-writer / accumulate for every grad buffer, in backward execution order
        seen: List[Tuple[int, int, int]] = []

        def covered(b: Buf):
            rid, lo, c = b.key()
            return any(r == rid and o <= lo and lo + c <= o + cc for (r, o, cc) in seen)

        for op in reversed(self.ops):
            acc = []
            for g in op.grad_writes():
                acc.append(covered(g))
                seen.append(g.key())
            op.acc = acc
        for op in self.ops:
            op.prepare()
        # table for the one-launch weight prep (bf16 KRSC + tap-flipped dgrad copies of every conv layer)
        items, tile = [], 0
        for op in self.ops:
            if isinstance(op, Conv):
                co, r, s_, ci = op.w.shape
                items.append([op.w.offset, co, r, s_, ci, int(op.needs_dgrad), tile, 0])
                tile += ((co + 31) // 32) * ((ci + 31) // 32)
            elif isinstance(op, DWConv):   # [k][k][C] == a Cout = 1 conv weight: bf16 copy only
                k, _, ci = op.w.shape
                items.append([op.w.offset, 1, k, k, ci, 0, tile, 0])
                tile += (ci + 31) // 32
        self._wprep_tiles = tile
        self._wprep_items = torch.tensor(items, dtype=torch.int64, device=dev) if items else None
        self.finalized = True

    def encoder_floats(self) -> int:
        """Number of leading floats of the flat buffers that belong to encoder parameters (they are created first)."""
        end = 0
        for name in self.encoder_param_names:
            p = self.params[name]
            end = max(end, p.offset + (p.size + 7) // 8 * 8)
        return end

    # pointers into flat buffers
    def pp(self, p: Param):
        return self.flat_p.data_ptr() + 4 * p.offset

    def pg(self, p: Param):
        return self.flat_g.data_ptr() + 4 * p.offset

    def pwf(self, p: Param):
        """forward conv operand: the bf16 KRSC copy (parity mode: the fp32 master itself)"""
        if self.precision == "fp32":
            return self.pp(p)
        return self.flat_wf.data_ptr() + 2 * p.offset

    def pwd(self, p: Param):
        """dgrad conv operand: the tap-flipped bf16 copy (parity mode: the fp32 master, flipped inside the kernel)"""
        if self.precision == "fp32":
            return self.pp(p)
        return self.flat_wd.data_ptr() + 2 * p.offset

    # ---- execution ----------------------------------------------------------------------------
    def prep_weights(self):
        if self.precision == "fp32":   # parity mode: the kernels read the fp32 master weights directly
            return
        if self._wprep_items is not None:
            self.L.weight_prep_batched(self.flat_p.data_ptr(), self.flat_wf.data_ptr(), self.flat_wd.data_ptr(),
                                       self._wprep_items.data_ptr(), self._wprep_items.shape[0], self._wprep_tiles,
                                       _stream())
        for op in self.ops:
            if isinstance(op, StemConv):
                op.prep_weights(_stream())

    def forward(self):
        for op in self.ops:
            op.fwd()

    def backward(self, lo: int = 0, hi: Optional[int] = None):
        """back-propagate through ops[lo:hi] in reverse order (whole graph by default); joins the weight-gradient side
        stream, so on return every gradient produced by these ops is complete on the current stream."""
        ops = self.ops[lo:hi]
        for op in reversed(ops):
            op.bwd()
        if self.overlap_wgrad and self.side_stream is not None:
            torch.cuda.current_stream().wait_stream(self.side_stream)

    def split_for_overlap(self, min_tail_fraction: float = 0.5) -> Tuple[int, int]:
        """(op index, float offset) cutting the network into an early part and a late part holding at least
        `min_tail_fraction` of the parameters: ops[i:] own exactly the parameters at flat offsets >= off (parameters are
        created in op order).  Used by the data-parallel step to all-reduce the late layers' gradients while the early
        layers are still back-propagating."""
        from . import ddp
        off = ddp.bucket_split([(p.offset, (p.size + 7) // 8 * 8) for p in self.params.values()], min_tail_fraction)
        first = None
        for i, op in enumerate(self.ops):
            ps = [getattr(op, a) for a in ("w", "b", "gamma", "beta") if isinstance(getattr(op, a, None), Param)]
            if ps and min(p.offset for p in ps) >= off:
                first = i
                break
        if first is None or off == 0:
            return 0, 0
        # every op from `first` on must own only tail parameters, every earlier op only head parameters
        for i, op in enumerate(self.ops):
            for a in ("w", "b", "gamma", "beta"):
                p = getattr(op, a, None)
                if isinstance(p, Param) and ((i >= first) != (p.offset >= off)):
                    return 0, 0
        return first, off

    def wgrad_stream(self):
        """context in which a weight-gradient kernel is enqueued: the side stream, ordered after everything already on
        the current stream (the layer's dY is complete), or the current stream when overlap is off."""
        if not (self.overlap_wgrad and self.side_stream is not None):
            return _NullCtx(self.ws)
        self.side_stream.wait_stream(torch.cuda.current_stream())
        return _SideCtx(self.side_stream, self.ws_side)

    # ---- weights in Keras layout ----------------------------------------------------------------
    def get_grads(self) -> Dict[str, np.ndarray]:
        """last backward's parameter gradients, Keras layouts (tests / debugging)."""
        return self._export(self.flat_g, False)

    def get_weights(self) -> Dict[str, np.ndarray]:
        return self._export(self.flat_p, True)

    def _export(self, flat, with_buffers) -> Dict[str, np.ndarray]:
        host = flat.detach().cpu().numpy()
        out = {}
        for p in self.params.values():
            a = host[p.offset:p.offset + p.size].reshape(p.shape)
            if p.kind == "conv":  # KRSC(padded) -> Keras HWIO
                cin = getattr(p, "cin_real", p.shape[3])
                a = np.transpose(a[:getattr(p, "cout_real", p.shape[0]), ..., :cin], (1, 2, 3, 0))
            elif p.kind == "dwconv":  # [k][k][C] -> Keras DepthwiseConv2D (k, k, C, 1)
                a = a[..., None]
            elif p.kind == "bias":
                a = a[:getattr(p, "cout_real", p.shape[0])]
            elif p.kind == "convT":  # internal [Cout][R][S][Cin], taps flipped -> Keras Conv2DTranspose (kh, kw, Cout, Cin)
                a = np.transpose(a[:, ::-1, ::-1, :], (1, 2, 0, 3))
            out[p.name] = np.ascontiguousarray(a)
        if with_buffers:
            for k, v in self.buffers.items():
                out[k] = v.detach().cpu().numpy().copy()
        return out

    def set_weights(self, d: Dict[str, np.ndarray], strict=True):
        host = self.flat_p.detach().cpu().numpy().copy()
        for p in self.params.values():
            if p.name not in d:
                if strict:
                    raise KeyError(p.name)
                continue
            a = np.asarray(d[p.name], dtype=np.float32)
            if p.kind == "convT":
                a = np.ascontiguousarray(np.transpose(a, (2, 0, 1, 3))[:, ::-1, ::-1, :])
            if p.kind == "dwconv":
                a = a.reshape(p.shape)
            if p.kind == "conv":
                a = np.transpose(a, (3, 0, 1, 2))  # HWIO -> KRSC
                full = np.zeros(p.shape, dtype=np.float32)
                full[:a.shape[0], ..., :a.shape[3]] = a
                a = full
            elif p.kind == "bias" and a.shape[0] < p.shape[0]:
                a = np.concatenate([a, np.zeros(p.shape[0] - a.shape[0], np.float32)])
            host[p.offset:p.offset + p.size] = a.reshape(-1)
        self.flat_p.copy_(torch.from_numpy(host))
        for k, v in self.buffers.items():
            if k in d:
                v.copy_(torch.from_numpy(np.asarray(d[k], dtype=np.float32)))


# -------------------------------------------------------------------------------------------------
# ops
# -------------------------------------------------------------------------------------------------
class InputNorm(Op):
    """bn_data (BatchNormalization(scale=False)) on the raw uint8 image fused with the bf16 cast and the
    pad-to-8-channels; channel c_img is the constant-ones channel (see stp_stem_wgrad_post)."""

    def __init__(self, net: Net, img: Buf, y: Buf, name: str, eps: float, momentum=0.99):
        self.net, self.img, self.y, self.eps, self.momentum = net, img, y, eps, momentum
        c = img.c
        self.beta = net.add_param(name + "/beta", (c,), "beta", lambda: np.zeros(c, np.float32))
        net.buffers[name + "/moving_mean"] = torch.zeros(c, device=net.device)
        net.buffers[name + "/moving_variance"] = torch.ones(c, device=net.device)
        self.mm, self.mv = net.buffers[name + "/moving_mean"], net.buffers[name + "/moving_variance"]
        self.coef = torch.zeros(4 * c, device=net.device)
        self.nblk = net.L.bn_nblk(img.rows, c)
        net.need_partial(2 * self.nblk * 8)
        net.ops.append(self)

    def prepare(self):
        pass

    def fwd(self):
        n, L, st = self.net, self.net.L, _stream()
        c = self.img.c
        if n.training:
            L.bn_stats(self.img.ref, n.partial.data_ptr(), st)
            L.bn_finalize(n.partial.data_ptr(), self.nblk, c, self.img.rows, None, n.pp(self.beta), self.eps,
                          self.momentum, self.mm.data_ptr(), self.mv.data_ptr(), self.coef.data_ptr(), st)
        else:
            L.bn_coef_infer(None, n.pp(self.beta), self.mm.data_ptr(), self.mv.data_ptr(), self.eps, c,
                            self.coef.data_ptr(), st)
        L.stem_prep(self.img.storage.data_ptr(), self.img.n, self.img.h, self.img.w, c, self.coef.data_ptr(),
                    self.y.ref, st)


class InputCast(Op):
    """raw uint8 image -> bf16 with the channel count padded to 8 (keras.applications VGG16 has no input BatchNorm; the
    reference feeds raw 0..255 pixels).  Padded channels meet zero weights (Conv zeroes their gradient columns)."""

    def __init__(self, net: Net, img: Buf, y: Buf):
        self.net, self.img, self.y = net, img, y
        c = img.c
        ident = np.zeros(4 * c, np.float32)
        ident[c:3 * c] = 1.0  # mean 0, invstd 1, scale 1, shift 0
        self.coef = torch.from_numpy(ident).to(net.device)
        net.ops.append(self)

    def fwd(self):
        n = self.net
        n.L.stem_prep(self.img.storage.data_ptr(), self.img.n, self.img.h, self.img.w, self.img.c, self.coef.data_ptr(),
                      self.y.ref, _stream())


class Conv(Op):
    def __init__(self, net: Net, x: Buf, y: Buf, name: str, k: int, stride=1, pad=0, residual: Optional[Buf] = None,
                 bias=False, init="he_uniform", needs_dgrad=True, cin_real: Optional[int] = None,
                 stem_beta: Optional[Param] = None, up=1, relu=False, transposed=False, cout_real: Optional[int] = None):
        self.net, self.x, self.y, self.name, self.k = net, x, y, name, k
        self.residual, self.needs_dgrad = residual, needs_dgrad
        self.relu = relu  # conv + bias + ReLU in one kernel (VGG encoder, decoder without BatchNorm); y is post-ReLU
        self.desc = _lib.ConvDesc(k, k, stride, pad, pad, up, _lib.CONV_RELU if relu else 0)
        cin, cout = x.c, y.c
        cr = cin_real or cin
        co_r = cout_real or cout  # output channels beyond cout_real are padding: zero weights, zero gradients
        fan_in, fan_out = k * k * cr, k * k * co_r
        lim = math.sqrt(6.0 / fan_in) if init == "he_uniform" else math.sqrt(6.0 / (fan_in + fan_out))

        def mk():
            # draw in Keras (kh,kw,Cin,Cout) order so the stream matches a Keras-layout initialiser, then -> KRSC
            w = net.gen.uniform(-lim, lim, size=(k, k, cr, co_r)).astype(np.float32)
            full = np.zeros((cout, k, k, cin), np.float32)
            full[:co_r, ..., :cr] = np.transpose(w, (3, 0, 1, 2))
            return full

        # transposed=True: keras Conv2DTranspose(k, strides=up, 'same') run as a convolution of the zero-inserted input
        # with the spatially flipped kernel (pad before = k-1-p); the parameter is exported in Keras (kh,kw,Cout,Cin) order
        self.w = net.add_param(name + "/kernel", (cout, k, k, cin), "convT" if transposed else "conv", mk)
        self.w.cin_real = cr
        self.w.cout_real = co_r
        self.b = net.add_param(name + "/bias", (cout,), "bias", lambda: np.zeros(cout, np.float32)) if bias else None
        if self.b is not None:
            self.b.cout_real = co_r
        self.stem_beta = stem_beta
        self.cin_real = cr
        self.bn_next: Optional["BNRelu"] = None
        self.bnb_prev: Optional["BNRelu"] = None   # BatchNorm whose backward reduction this conv's dgrad performs
        net.need_ws(net.L.conv_wgrad_workspace(C.byref(self.desc), x.ref, y.ref))
        if bias:
            net.need_partial(2 * net.L.bn_nblk(y.rows, y.c) * y.c)
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()] if self.needs_dgrad else []

    def prepare(self):
        n = self.net
        self.dref = C.byref(self.desc)
        self.res_ref = self.residual.ref if self.residual is not None else None
        self.dy = self.y.grad()
        if self.needs_dgrad:
            self.dx = self.x.grad()
            self.dx_res = self.dx.ref if self.acc[0] else None

    def prep_weights(self, st):
        n = self.net
        c = self.w.shape
        n.L.weight_prep(n.pp(self.w), n.pwf(self.w), n.pwd(self.w) if self.needs_dgrad else None, c[0], c[1], c[2],
                        c[3], st)

    def fwd(self):
        n = self.net
        if "conv_fwd" in _SKIP:
            return
        if self.bn_next is not None and n.training:
            n.L.conv_fwd_bn(self.dref, self.x.ref, n.pwf(self.w), n.pp(self.b) if self.b else None, self.res_ref,
                            self.y.ref, self.bn_next.bn_fwd_struct(), n.ws.data_ptr(), n.ws.numel(), _stream())
        else:
            n.L.conv_fwd(self.dref, self.x.ref, n.pwf(self.w), n.pp(self.b) if self.b else None, self.res_ref,
                         self.y.ref, n.ws.data_ptr(), n.ws.numel(), _stream())

    def bwd(self):
        n = self.net
        if self.relu:
            # dz = dy * (y > 0), in place in the gradient buffer of y (nothing else reads dy afterwards)
            n.L.relu_bwd(self.dy.ref, self.y.ref, 1, None, self.dy.ref, _stream())
        if self.b is not None:
            n.L.bias_grad(self.dy.ref, n.partial.data_ptr(), n.sync.data_ptr(), n.bn_acc.data_ptr(), n.pg(self.b), _stream())
        with n.wgrad_stream() as ws:  # forked first: the wgrad overlaps this layer's dgrad and the BatchNorm backward below
            st = _stream()
            if "wgrad" not in _SKIP:
                n.L.conv_wgrad(self.dref, self.x.ref, self.dy.ref, n.pg(self.w), ws.data_ptr(), ws.numel(), st)
            if self.stem_beta is not None or self.cin_real < self.w.shape[3]:
                c = self.w.shape
                # (parity mode passes -R: d(beta) is then formed with the fp32 weights instead of their bf16 rounding)
                n.L.stem_wgrad_post(n.pg(self.w), n.pp(self.w), c[0], -c[1] if n.precision == "fp32" else c[1], c[2], c[3],
                                    self.cin_real, n.pg(self.stem_beta) if self.stem_beta is not None else None, st)
        if self.needs_dgrad and "dgrad" not in _SKIP:
            if self.bnb_prev is not None and self.dx_res is None:
                n.L.conv_dgrad_bn(self.dref, self.dy.ref, n.pwd(self.w), self.dx.ref, self.bnb_prev.bn_bwd_struct(),
                                  n.ws.data_ptr(), n.ws.numel(), _stream())
            else:
                n.L.conv_dgrad(self.dref, self.dy.ref, n.pwd(self.w), self.dx_res, self.dx.ref, n.ws.data_ptr(),
                               n.ws.numel(), _stream())


class StemConv(Op):
    """conv0: 7x7 stride-2 pad-3 convolution of the (8-channel padded) bn_data output, executed as a 4x4 stride-1
    convolution over the space-to-depth tensor [N, H/2, W/2, 32] that InputNorm writes (DESIGN.md "stem") so that it runs
    on the tcgen05 halo kernel.  The parameter stays in Keras geometry ([64][7][7][8] master, exported as (7,7,3,64)); the
    4x4x32 bf16 operand and the gradient mapping back are two tiny kernels."""

    def __init__(self, net: Net, x_s2d: Buf, y: Buf, name: str, cin_real: int, stem_beta: Param, init="he_uniform"):
        self.net, self.x, self.y, self.name = net, x_s2d, y, name
        assert x_s2d.c == 32 and y.h == x_s2d.h and y.w == x_s2d.w
        cout, k, cin = y.c, 7, 8
        lim = math.sqrt(6.0 / (k * k * cin_real)) if init == "he_uniform" else math.sqrt(6.0 / (k * k * (cin_real + cout)))

        def mk():
            w = net.gen.uniform(-lim, lim, size=(k, k, cin_real, cout)).astype(np.float32)
            full = np.zeros((cout, k, k, cin), np.float32)
            full[..., :cin_real] = np.transpose(w, (3, 0, 1, 2))
            return full

        self.w = net.add_param(name + "/kernel", (cout, k, k, cin), "conv", mk)
        self.w.cin_real = cin_real
        self.cin_real, self.stem_beta = cin_real, stem_beta
        self.desc = _lib.ConvDesc(4, 4, 1, 2, 2, 1, 0)
        self.bn_next: Optional["BNRelu"] = None
        self.w2 = torch.zeros(cout * 16 * 32, dtype=torch.bfloat16, device=net.device)
        self.dw2 = torch.zeros(cout * 16 * 32, dtype=torch.float32, device=net.device)
        net.need_ws(net.L.conv_wgrad_workspace(C.byref(self.desc), x_s2d.ref, y.ref))
        net.ops.append(self)

    def prepare(self):
        self.dref = C.byref(self.desc)
        self.dy = self.y.grad()

    def prep_weights(self, st):
        n = self.net
        n.L.stem_weight_s2d(n.pp(self.w), self.w2.data_ptr(), self.w.shape[0], st)

    def fwd(self):
        n = self.net
        if self.bn_next is not None and n.training:
            n.L.conv_fwd_bn(self.dref, self.x.ref, self.w2.data_ptr(), None, None, self.y.ref,
                            self.bn_next.bn_fwd_struct(), n.ws.data_ptr(), n.ws.numel(), _stream())
        else:
            n.L.conv_fwd(self.dref, self.x.ref, self.w2.data_ptr(), None, None, self.y.ref, n.ws.data_ptr(),
                         n.ws.numel(), _stream())

    def bwd(self):
        n = self.net
        c = self.w.shape
        with n.wgrad_stream() as ws:
            st = _stream()
            n.L.conv_wgrad(self.dref, self.x.ref, self.dy.ref, self.dw2.data_ptr(), ws.data_ptr(), ws.numel(), st)
            n.L.stem_wgrad_s2d_gather(self.dw2.data_ptr(), n.pg(self.w), c[0], st)
            n.L.stem_wgrad_post(n.pg(self.w), n.pp(self.w), c[0], c[1], c[2], c[3], self.cin_real, n.pg(self.stem_beta),
                                st)


class BNRelu(Op):
    """BatchNormalization (+ReLU) (+ fused UpSampling2D(2) on the output write)."""

    def __init__(self, net: Net, x: Buf, y: Buf, name: str, eps: float, relu=True, up=1, momentum=0.99,
                 extra_grad: Optional[Callable[[], Buf]] = None):
        self.net, self.x, self.y, self.eps, self.relu, self.up, self.momentum = net, x, y, eps, relu, up, momentum
        self.extra_grad = extra_grad
        c = x.c
        self.gamma = net.add_param(name + "/gamma", (c,), "gamma", lambda: np.ones(c, np.float32))
        self.beta = net.add_param(name + "/beta", (c,), "beta", lambda: np.zeros(c, np.float32))
        net.buffers[name + "/moving_mean"] = torch.zeros(c, device=net.device)
        net.buffers[name + "/moving_variance"] = torch.ones(c, device=net.device)
        self.mm, self.mv = net.buffers[name + "/moving_mean"], net.buffers[name + "/moving_variance"]
        self.coef = torch.zeros(4 * c, device=net.device)
        self.bcoef = torch.zeros(3 * c, device=net.device)
        self.nblk = net.L.bn_nblk(x.rows, c)
        net.need_partial(2 * self.nblk * c)
        self.stats_from_conv = False
        self.reduce_from_dgrad = False
        self._bn_fwd = None
        self._bn_bwd = None
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def bn_fwd_struct(self):
        """stp_bn_fwd for the conv that produces this layer's input (statistics come out of its epilogue)."""
        if self._bn_fwd is None:
            n = self.net
            self._bn_fwd = _lib.BnFwd(n.partial.data_ptr(), n.sync.data_ptr(), n.bn_acc.data_ptr(), n.pp(self.gamma),
                                      n.pp(self.beta), self.eps, self.momentum, self.mm.data_ptr(), self.mv.data_ptr(),
                                      self.coef.data_ptr())
        return C.byref(self._bn_fwd)

    def bn_bwd_struct(self):
        """stp_bn_bwd for the conv whose dgrad writes this layer's output gradient (reduction in its epilogue)."""
        if self._bn_bwd is None:
            n = self.net
            self._bn_bwd = _lib.BnBwd(C.pointer(self.x.st), self.coef.data_ptr(), int(self.relu), n.partial.data_ptr(),
                                      n.sync.data_ptr(), n.bn_acc.data_ptr(), n.pg(self.gamma), n.pg(self.beta),
                                      self.bcoef.data_ptr())
        return C.byref(self._bn_bwd)

    def prepare(self):
        self.dy = self.y.grad()
        self.dx = self.x.grad()
        extra = self.extra_grad() if self.extra_grad else None
        if extra is not None and self.acc[0]:
            raise RuntimeError("BNRelu: both identity-shortcut gradient and accumulate requested")
        self.res_ref = extra.ref if extra is not None else (self.dx.ref if self.acc[0] else None)

    def fwd(self):
        n, L, st = self.net, self.net.L, _stream()
        c = self.x.c
        if n.training:
            if not self.stats_from_conv:
                L.bn_stats_fused(self.x.ref, n.partial.data_ptr(), n.sync.data_ptr(), n.bn_acc.data_ptr(), n.pp(self.gamma), n.pp(self.beta),
                                 self.eps, self.momentum, self.mm.data_ptr(), self.mv.data_ptr(), self.coef.data_ptr(), st)
        else:
            L.bn_coef_infer(n.pp(self.gamma), n.pp(self.beta), self.mm.data_ptr(), self.mv.data_ptr(), self.eps, c,
                            self.coef.data_ptr(), st)
        if "bn_apply" not in _SKIP:
            L.bn_apply(self.x.ref, self.coef.data_ptr(), int(self.relu), self.up, self.y.ref, st)

    def bwd(self):
        n, L, st = self.net, self.net.L, _stream()
        c = self.x.c
        if "bn_bwd_apply" in _SKIP and "bn_bwd_reduce" in _SKIP:
            return
        if not self.reduce_from_dgrad and "bn_bwd_reduce" not in _SKIP:
            L.bn_bwd_reduce_fused(self.dy.ref, self.x.ref, self.coef.data_ptr(), int(self.relu), self.up,
                                  n.partial.data_ptr(), n.sync.data_ptr(), n.bn_acc.data_ptr(), n.pg(self.gamma), n.pg(self.beta),
                                  self.bcoef.data_ptr(), st)
        if "bn_bwd_apply" not in _SKIP:
            L.bn_bwd_apply(self.dy.ref, self.x.ref, self.coef.data_ptr(), self.bcoef.data_ptr(), int(self.relu), self.up,
                           self.res_ref, self.dx.ref, st)


class Add(Op):
    """keras.layers.Add of a decoder branch and an encoder skip (Linknet).  Backward: the branch's gradient buffer IS
    d(out) (aliased, no kernel); the skip receives d(out) as its first gradient contribution (copy) or accumulates it."""

    def __init__(self, net: Net, a: Buf, skip: Buf, y: Buf):
        self.net, self.a, self.skip, self.y = net, a, skip, y
        a.set_grad(y.grad())
        net.ops.append(self)

    def grad_writes(self):
        return [self.skip.grad()]

    def prepare(self):
        self.dy, self.ds = self.y.grad(), self.skip.grad()

    def fwd(self):
        self.net.L.add(self.a.ref, self.skip.ref, self.y.ref, _stream())

    def bwd(self):
        L = self.net.L
        if self.acc[0]:
            L.add(self.ds.ref, self.dy.ref, self.ds.ref, _stream())
        else:
            L.copy_up(self.dy.ref, 1, self.ds.ref, _stream())


class UpCopy(Op):
    """UpSampling2D(2) of a post-ReLU tensor straight into (a channel slice of) a concat buffer.  Backward: 2x2 sum of
    the upsampled gradient; positions where the (non-negative) source is exactly 0 receive 0, which the ReLU mask of
    the producing layer would apply anyway."""

    def __init__(self, net: Net, x: Buf, y: Buf):
        self.net, self.x, self.y = net, x, y
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        if self.acc[0]:
            raise RuntimeError("UpCopy: accumulate into dx unsupported")

    def fwd(self):
        self.net.L.copy_up(self.x.ref, 2, self.y.ref, _stream())

    def bwd(self):
        self.net.L.relu_bwd(self.dy.ref, self.x.ref, 2, None, self.dx.ref, _stream())


class Upsample2x(Op):
    """UpSampling2D(2, nearest) of a LINEAR tensor (the FPN top-down pathway upsamples 1x1-conv outputs, so the ReLU mask
    UpCopy applies in its backward would be wrong).  Backward: 2x2 sum of the gradient, accumulated into dx when another
    consumer of x already wrote it."""

    def __init__(self, net: Net, x: Buf, y: Buf):
        self.net, self.x, self.y = net, x, y
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        self.res_ref = self.dx.ref if self.acc[0] else None

    def fwd(self):
        self.net.L.copy_up(self.x.ref, 2, self.y.ref, _stream())

    def bwd(self):
        self.net.L.upsample2x_bwd(self.dy.ref, self.res_ref, self.dx.ref, _stream())


class Resize(Op):
    """keras UpSampling2D(rate, interpolation='bilinear') == TF1 legacy tf.image.resize_bilinear (FPN segmentation branches);
    y is typically a channel slice of the concat buffer.  Backward: deterministic gather."""

    def __init__(self, net: Net, x: Buf, y: Buf, align_corners=False):
        # align_corners=True: the BilinearUpsampling layer of the reference's DeepLabV3+ (impl/deeplab/model.py:81-100)
        self.net, self.x, self.y, self.align = net, x, y, bool(align_corners)
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        self.res_ref = self.dx.ref if self.acc[0] else None

    def fwd(self):
        L = self.net.L
        (L.resize_bilinear_ac_fwd if self.align else L.resize_bilinear_fwd)(self.x.ref, self.y.ref, _stream())

    def bwd(self):
        L = self.net.L
        (L.resize_bilinear_ac_bwd if self.align else L.resize_bilinear_bwd)(self.dy.ref, self.res_ref, self.dx.ref, _stream())


class MaxPool(Op):
    def __init__(self, net: Net, x: Buf, y: Buf, k, stride, pad):
        self.net, self.x, self.y, self.k, self.stride, self.pad = net, x, y, k, stride, pad
        self.argmax = torch.zeros(y.rows * y.c, dtype=torch.uint8, device=net.device)
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        self.res_ref = self.dx.ref if self.acc[0] else None

    def fwd(self):
        self.net.L.maxpool_fwd(self.x.ref, self.k, self.stride, self.pad, self.y.ref, self.argmax.data_ptr(), _stream())

    def bwd(self):
        self.net.L.maxpool_bwd(self.dy.ref, self.argmax.data_ptr(), self.k, self.stride, self.pad, self.res_ref,
                               self.dx.ref, _stream())


class AvgPool(Op):
    """keras AveragePooling2D(k, strides k) over windows that tile the input exactly (PSPNet pyramid pooling)."""

    def __init__(self, net: Net, x: Buf, y: Buf, k: int):
        assert x.h == y.h * k and x.w == y.w * k and x.c == y.c
        self.net, self.x, self.y, self.k = net, x, y, int(k)
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        self.res_ref = self.dx.ref if self.acc[0] else None

    def fwd(self):
        self.net.L.avgpool_fwd(self.x.ref, self.k, self.y.ref, _stream())

    def bwd(self):
        self.net.L.avgpool_bwd(self.dy.ref, self.k, self.res_ref, self.dx.ref, _stream())


class Head(Op):
    """final_conv (3x3 same, bias) -> logits f32 [M, classes]; sigmoid lives in the loss / predict kernels."""

    def __init__(self, net: Net, x: Buf, classes: int, name="final_conv", init="glorot_uniform"):
        self.net, self.x, self.classes = net, x, classes
        cin = x.c
        fan_in, fan_out = 9 * cin, 9 * classes
        lim = math.sqrt(6.0 / (fan_in + fan_out)) if init == "glorot_uniform" else math.sqrt(6.0 / fan_in)

        def mk():
            w = net.gen.uniform(-lim, lim, size=(3, 3, cin, classes)).astype(np.float32)
            return np.ascontiguousarray(np.transpose(w, (3, 0, 1, 2)))

        self.w = net.add_param(name + "/kernel", (classes, 3, 3, cin), "conv", mk)
        self.b = net.add_param(name + "/bias", (classes,), "bias", lambda: np.zeros(classes, np.float32))
        self.logits = torch.zeros(x.rows * classes, dtype=torch.float32, device=net.device)
        self.dlogits = torch.zeros(x.rows * classes, dtype=torch.float32, device=net.device)
        self.out_n, self.out_hw = x.n, x.h * x.w  # geometry of the logits the loss sees
        net.need_ws(net.L.head_bwd_workspace(x.ref, classes))
        net.need_ws(net.L.head_fwd_workspace(x.ref, classes))
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dx = self.x.grad()
        if self.acc[0]:
            raise RuntimeError("Head: accumulate into dx unsupported")

    def fwd(self):
        n = self.net
        n.L.head_fwd(self.x.ref, n.pp(self.w), n.pp(self.b), self.classes, self.logits.data_ptr(), n.ws.data_ptr(),
                     n.ws.numel(), _stream())

    def bwd(self):
        n = self.net
        n.L.head_bwd(self.x.ref, n.pp(self.w), self.dlogits.data_ptr(), self.classes, self.dx.ref, n.pg(self.w),
                     n.pg(self.b), n.ws.data_ptr(), n.ws.numel(), _stream())


class UpHead(Conv):
    """FPN head: Conv2D(classes, 3x3, same, bias) on the merged pyramid at H/4, then the x4 bilinear `last_upsample`
    (schema segmentation.raml:179-204) -> logits f32 [N*H*W, classes] for the loss / predict kernels.  The conv runs on
    the tensor-core conv path with the class dimension padded to 16 output channels (zero weights, zero gradients) and
    fp32 output; the gradient of the small logits is materialised as bf16 like every other activation gradient."""

    CPAD = 16

    def __init__(self, net: Net, x: Buf, classes: int, name="head_conv", up=4, init="glorot_uniform"):
        if classes > self.CPAD:
            raise NotImplementedError("UpHead: classes <= %d" % self.CPAD)
        small = Buf(net, x.n, x.h, x.w, self.CPAD, F32, name=name + "_out")
        small.set_grad(Buf(net, x.n, x.h, x.w, self.CPAD, net.act_dtype, name="d_" + name + "_out"))
        super().__init__(net, x, small, name, 3, pad=1, bias=True, init=init, cout_real=classes)
        self.classes, self.small = classes, small
        H, W = x.h * up, x.w * up
        self.out_n, self.out_hw = x.n, H * W
        self.logits = torch.zeros(x.n * H * W * classes, dtype=torch.float32, device=net.device)
        self.dlogits = torch.zeros(x.n * H * W * classes, dtype=torch.float32, device=net.device)
        self._lg = _lib.Tensor(self.logits.data_ptr(), x.n, H, W, classes, classes, F32)
        self._dlg = _lib.Tensor(self.dlogits.data_ptr(), x.n, H, W, classes, classes, F32)

    def fwd(self):
        super().fwd()
        self.net.L.resize_bilinear_fwd(self.small.ref, C.byref(self._lg), _stream())

    def bwd(self):
        self.net.L.resize_bilinear_bwd(C.byref(self._dlg), None, self.dy.ref, _stream())
        super().bwd()


class DWConv(Op):
    """keras DepthwiseConv2D(k, strides, 'same', dilation_rate, use_bias=False) (MobileNetV2 blocks of the reference's
    DeepLabV3+, impl/deeplab/model.py:252-255).  Parameter `<name>/depthwise_kernel`, Keras shape (k, k, C, 1); internal
    [k][k][C] fp32 master with a bf16 copy in the flat forward-weight buffer (made by the batched weight prep as a Cout = 1
    item)."""

    def __init__(self, net: Net, x: Buf, y: Buf, name: str, k=3, stride=1, dilation=1, init="glorot_uniform", pad=None):
        self.net, self.x, self.y, self.name = net, x, y, name
        c = x.c
        assert y.c == c

        def same(size):   # TF 'same': padding before = total // 2
            out = -(-size // stride)
            return max((out - 1) * stride + (k - 1) * dilation + 1 - size, 0) // 2, out

        def explicit(size):   # ZeroPadding2D((pad, pad_end)) + 'valid' (SepConv_BN with stride > 1, model.py:125-131)
            ke = (k - 1) * dilation + 1
            return pad, (size + (ke - 1) - ke) // stride + 1

        (ph, ho), (pw, wo) = (same(x.h), same(x.w)) if pad is None else (explicit(x.h), explicit(x.w))
        assert (ho, wo) == (y.h, y.w), ((ho, wo), (y.h, y.w))
        self.desc = _lib.DwConvDesc(k, stride, dilation, ph, pw)
        lim = math.sqrt(6.0 / (k * k * c + k * k)) if init == "glorot_uniform" else math.sqrt(6.0 / (k * k * c))
        self.w = net.add_param(name + "/depthwise_kernel", (k, k, c), "dwconv",
                               lambda: net.gen.uniform(-lim, lim, size=(k, k, c)).astype(np.float32))
        net.need_ws(net.L.dwconv_wgrad_workspace(C.byref(self.desc), x.ref, y.ref))
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dref = C.byref(self.desc)
        self.dy, self.dx = self.y.grad(), self.x.grad()
        self.dx_res = self.dx.ref if self.acc[0] else None

    def fwd(self):
        n = self.net
        n.L.dwconv_fwd(self.dref, self.x.ref, n.pwf(self.w), self.y.ref, _stream())

    def bwd(self):
        n = self.net
        with n.wgrad_stream() as ws:
            n.L.dwconv_wgrad(self.dref, self.x.ref, self.dy.ref, n.pg(self.w), ws.data_ptr(), ws.numel(), _stream())
        n.L.dwconv_dgrad(self.dref, self.dy.ref, n.pwf(self.w), self.dx_res, self.dx.ref, _stream())


class Relu(Op):
    """Stand-alone Activation('relu') (the pre-activation of SepConv_BN with depth_activation=False, impl/deeplab/model.py:133-134):
    the BatchNorm-apply kernel with identity coefficients; backward masks by y > 0."""

    def __init__(self, net: Net, x: Buf, y: Buf):
        self.net, self.x, self.y = net, x, y
        c = x.c
        ident = np.zeros(4 * c, np.float32)
        ident[c:3 * c] = 1.0   # mean 0, invstd 1, scale 1, shift 0
        self.coef = torch.from_numpy(ident).to(net.device)
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        self.res_ref = self.dx.ref if self.acc[0] else None

    def fwd(self):
        self.net.L.bn_apply(self.x.ref, self.coef.data_ptr(), 1, 1, self.y.ref, _stream())

    def bwd(self):
        self.net.L.relu_bwd(self.dy.ref, self.y.ref, 1, self.res_ref, self.dx.ref, _stream())


class GlobalAvgPool(Op):
    """AveragePooling2D over the whole map (DeepLabV3+ image-pooling branch, impl/deeplab/model.py:462): [N,h,w,C] -> [N,1,1,C]."""

    def __init__(self, net: Net, x: Buf, y: Buf):
        self.net, self.x, self.y = net, x, y
        assert (y.h, y.w, y.c) == (1, 1, x.c)
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        self.res_ref = self.dx.ref if self.acc[0] else None

    def fwd(self):
        self.net.L.global_avgpool_fwd(self.x.ref, self.y.ref, _stream())

    def bwd(self):
        self.net.L.global_avgpool_bwd(self.dy.ref, self.res_ref, self.dx.ref, _stream())


class Broadcast(Op):
    """BilinearUpsampling of a 1x1 map (impl/deeplab/model.py:469): every output pixel is the one input pixel; y is typically
    a channel slice of the ASPP concat buffer.  Backward: spatial sum."""

    def __init__(self, net: Net, x: Buf, y: Buf):
        self.net, self.x, self.y = net, x, y
        assert (x.h, x.w, x.c) == (1, 1, y.c)
        net.ops.append(self)

    def grad_writes(self):
        return [self.x.grad()]

    def prepare(self):
        self.dy, self.dx = self.y.grad(), self.x.grad()
        if self.acc[0]:
            raise RuntimeError("Broadcast: accumulate into dx unsupported")

    def fwd(self):
        self.net.L.broadcast_fwd(self.x.ref, self.y.ref, _stream())

    def bwd(self):
        self.net.L.broadcast_bwd(self.dy.ref, self.dx.ref, _stream())


class Dropout(Op):
    """keras Dropout(rate), in place on a post-activation buffer (impl/deeplab/model.py:486).  Training: x *= keep/(1-rate) with
    the Philox mask of (net.seed, salt, net.d_step, element); inference: identity.  Backward: the same mask on the gradient."""

    def __init__(self, net: Net, x: Buf, rate: float, salt: int):
        self.net, self.x, self.rate, self.salt = net, x, float(rate), int(salt)
        net.ops.append(self)

    def grad_writes(self):
        return []

    def prepare(self):
        self.dx = self.x.grad()

    def fwd(self):
        n = self.net
        if n.training and self.rate > 0.0:
            n.L.dropout(self.x.ref, self.rate, n.seed, self.salt, n.d_step.data_ptr(), self.x.ref, _stream())

    def bwd(self):
        n = self.net
        if self.rate > 0.0:
            n.L.dropout(self.dx.ref, self.rate, n.seed, self.salt, n.d_step.data_ptr(), self.dx.ref, _stream())


class ProbHead(Conv):
    """DeepLabV3+ head (impl/deeplab/model.py:494-500): Conv2D(classes, 1x1, bias, activation) at 1/8 resolution, then the
    align_corners BilinearUpsampling of the PROBABILITIES to the input size.  `logits` is what the loss / predict kernels
    read: activation(logits) == resize(activation(conv)) (stp_prob_head_fwd).  Class dimension padded to 16 like UpHead."""

    CPAD = 16
    ACT = {"sigmoid": 1, "softmax": 2}

    def __init__(self, net: Net, x: Buf, classes: int, out_hw, activation: str, name: str, init="glorot_uniform"):
        if classes > self.CPAD:
            raise NotImplementedError("ProbHead: classes <= %d" % self.CPAD)
        if activation not in self.ACT:
            raise NotImplementedError("DeepLabV3 head: activation sigmoid or softmax (got %r)" % (activation,))
        small = Buf(net, x.n, x.h, x.w, self.CPAD, F32, name=name + "_out")
        small.set_grad(Buf(net, x.n, x.h, x.w, self.CPAD, net.act_dtype, name="d_" + name + "_out"))
        super().__init__(net, x, small, name, 1, pad=0, bias=True, init=init, cout_real=classes)
        self.classes, self.small, self.act = classes, small, self.ACT[activation]
        H, W = out_hw
        self.out_n, self.out_hw = x.n, H * W
        self.logits = torch.zeros(x.n * H * W * classes, dtype=torch.float32, device=net.device)
        self.dlogits = torch.zeros(x.n * H * W * classes, dtype=torch.float32, device=net.device)
        self._lg = _lib.Tensor(self.logits.data_ptr(), x.n, H, W, classes, classes, F32)
        self._dlg = _lib.Tensor(self.dlogits.data_ptr(), x.n, H, W, classes, classes, F32)

    def fwd(self):
        super().fwd()
        self.net.L.prob_head_fwd(self.small.ref, self.classes, self.act, C.byref(self._lg), _stream())

    def bwd(self):
        self.net.L.prob_head_bwd(C.byref(self._dlg), C.byref(self._lg), self.small.ref, self.classes, self.act, self.dy.ref, _stream())
        super().bwd()


class Loss(Op):
    """w_bce*binary_crossentropy + w_dice*dice_loss + w_iou*iou_loss and the metrics, fused with the sigmoid; or
    w_lovasz*lovasz_loss on the logits (the reference's compile strips the final Activation for it)."""

    def __init__(self, net: Net, head: Head, mask: Buf, w_bce=1.0, w_dice=0.0, w_iou=0.0, w_lovasz=0.0, w_jaccard=0.0,
                 w_focal=0.0, w_cce=0.0, lovasz_act="elu"):
        self.net, self.head, self.mask = net, head, mask
        self.result = torch.zeros(16, dtype=torch.float32, device=net.device)
        self.lpartial = torch.zeros(net.L.loss_partial_floats(), dtype=torch.float32, device=net.device)
        self.count = head.out_n * head.out_hw * head.classes
        self.enabled = True
        self.lov_ws: Optional[torch.Tensor] = None
        self.lovasz_act = lovasz_act
        self.set_weights(w_bce, w_dice, w_iou, w_lovasz, w_jaccard, w_focal, w_cce)
        net.ops.append(self)

    def prepare(self):
        pass

    def set_weights(self, w_bce, w_dice, w_iou, w_lovasz=0.0, w_jaccard=0.0, w_focal=0.0, w_cce=0.0):
        self.spec = _lib.LossSpec(w_bce, w_dice, w_iou, w_jaccard, w_focal)
        self.w_lovasz = float(w_lovasz)
        self.w_cce = float(w_cce)   # categorical_crossentropy on softmax probabilities (stp_softmax_cce_fwd)
        if self.w_cce != 0.0 and self.head.classes < 2:
            raise ValueError("categorical_crossentropy needs classes >= 2")
        if self.w_lovasz != 0.0:
            # classes > 1: the reference's K.squeeze(..., -1) is undefined; one hinge per (image, class), averaged
            # (oracle/losses.py lovasz_loss)
            if self.lov_ws is None:
                h = self.head
                nbytes = self.net.L.lovasz_workspace(h.out_n * h.classes, h.out_hw)
                self.lov_ws = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=self.net.device)

    def fwd(self):
        if self.enabled:
            L, h = self.net.L, self.head
            L.loss_fwd(self.head.logits.data_ptr(), self.mask.storage.data_ptr(), self.count, C.byref(self.spec),
                       self.lpartial.data_ptr(), self.result.data_ptr(), _stream())
            if self.w_lovasz != 0.0:
                L.lovasz_fwd_mc(self.head.logits.data_ptr(), self.mask.storage.data_ptr(), h.out_n, h.out_hw, h.classes,
                                int(self.lovasz_act == "elu"), self.w_lovasz, 1, self.lov_ws.data_ptr(),
                                self.lov_ws.numel(), self.result.data_ptr(), _stream())
            if self.w_cce != 0.0:
                L.softmax_cce_fwd(self.head.logits.data_ptr(), self.mask.storage.data_ptr(), h.out_n * h.out_hw, h.classes,
                                  self.w_cce, 1, self.lpartial.data_ptr(), self.result.data_ptr(), _stream())

    def bwd(self):
        L, h = self.net.L, self.head
        L.loss_bwd(self.head.logits.data_ptr(), self.mask.storage.data_ptr(), self.count, C.byref(self.spec),
                   self.result.data_ptr(), self.head.dlogits.data_ptr(), _stream())
        if self.w_lovasz != 0.0:
            L.lovasz_bwd(self.lov_ws.data_ptr(), self.lov_ws.numel(), h.out_n * h.classes, h.out_hw, self.w_lovasz, 1,
                         self.head.dlogits.data_ptr(), _stream())
        if self.w_cce != 0.0:
            L.softmax_cce_bwd(self.head.logits.data_ptr(), self.mask.storage.data_ptr(), h.out_n * h.out_hw, h.classes,
                              self.w_cce, 1, self.head.dlogits.data_ptr(), _stream())
