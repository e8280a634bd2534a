"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU
tests), pure data parallelism -- the only partition this path has (reference FAQ.md:108-112 `cfg.gpus = N` ->
keras.utils.multi_gpu_model [DEP]: batch split across towers, per-tower BatchNorm statistics, gradients summed).

Per step each rank trains on its own shard of the sample stream and the flat fp32 gradient buffer (one contiguous
tensor for the whole network, engine.Net.flat_g) is all-reduced ONCE; the 1/world mean is folded into the optimizer
kernel (stp_grad_xform.scale), BatchNorm statistics stay local like the reference's towers."""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: Optional[str] = None, device: Optional[torch.device] = None):
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_indices(indices: Sequence[int], rank: int, world: int, batch: int) -> np.ndarray:
    """Rank's share of an epoch's (already shuffled) sample order: global batch g = [g*world*batch, (g+1)*world*batch)
    is cut into `world` contiguous per-rank batches -- the same split keras multi_gpu_model makes of one global batch.
    Every rank gets the same number of full batches (the tail that does not fill a global batch is dropped, like
    fit_generator's steps = len // batch)."""
    idx = np.asarray(indices)
    gb = world * batch
    n_global = len(idx) // gb
    out = []
    for g in range(n_global):
        lo = g * gb + rank * batch
        out.append(idx[lo:lo + batch])
    return np.concatenate(out) if out else idx[:0]


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the flat gradient buffer (no-op for world_size 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def allreduce_sum_async(flat: torch.Tensor, group=None):
    """Enqueue a SUM all-reduce that runs on the backend's own stream, ordered after what is already enqueued on the
    current stream; returns a handle whose wait() orders the current stream after the collective (no host block on
    NCCL).  None for world_size 1.  Lets the all-reduce of the late layers' gradients overlap the rest of the backward."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return None


def bucket_split(param_offsets_sizes, min_tail_fraction: float = 0.5) -> int:
    """Float offset that cuts the flat gradient buffer (parameters in forward order) into [head | tail]: the tail holds at
    least `min_tail_fraction` of the floats and starts at a parameter boundary.  The tail's gradients are complete first
    (backward runs in reverse layer order), so its all-reduce can start while the head's layers are still back-propagating."""
    items = sorted(param_offsets_sizes)
    if not items:
        return 0
    end = max(off + sz for off, sz in items)
    best = 0
    for off, _ in items:
        if end - off >= end * min_tail_fraction:
            best = max(best, off)
    return best


def broadcast_(flat: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Make every rank start from rank `src`'s parameters."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def max_over_ranks(values: List[float], device="cpu", group=None) -> List[float]:
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t.cpu()]


def mean_over_ranks_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks (no-op for world_size 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= dist.get_world_size(group)
    return t


def sync_buffers_mean_(buffers, group=None):
    """Average a dict of small fp32 tensors (the BatchNorm moving statistics, local per rank like the reference's per-tower
    BatchNorm) across ranks through ONE flat all-reduce, so that every rank validates -- and therefore takes every
    callback decision (EarlyStopping, ReduceLROnPlateau, best-weights checkpoint) -- on identical numbers."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1) or not buffers:
        return
    keys = sorted(buffers)
    flat = torch.cat([buffers[k].reshape(-1).float() for k in keys])
    mean_over_ranks_(flat, group)
    off = 0
    for k in keys:
        n = buffers[k].numel()
        buffers[k].copy_(flat[off:off + n].view_as(buffers[k]))
        off += n


def barrier(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)
