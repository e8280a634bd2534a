"""world_size-2 gloo test of the data-parallel plumbing (ddp.py): the flat-gradient SUM all-reduce + 1/world mean and
the parameter broadcast, run as two real processes on CPU.  Also checks the DP identity the engine relies on:
mean over ranks of per-shard mean-loss gradients == gradient of the global-batch mean loss (BatchNorm-free model)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from segmentation_training_pipeline_b200 import ddp
    r, lr, w = ddp.init(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    wgt = torch.randn(37, 5)                          # identical "parameters" on every rank (seeded)
    if rank == 1:
        wgt += 1.0                                    # ... until rank 1 drifts; broadcast must repair it
    ddp.broadcast_(wgt, 0)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(8, 37, generator=g)
    t = torch.randn(8, 5, generator=g)
    order = np.arange(8)
    mine = ddp.shard_indices(order, rank, world, 4)
    wl = wgt.clone().requires_grad_(True)
    loss = ((x[mine] @ wl - t[mine]) ** 2).mean()
    loss.backward()
    flat = wl.grad.reshape(-1).clone()
    ddp.allreduce_sum_(flat)
    flat *= 1.0 / world                               # what stp_grad_xform.scale does inside the optimizer kernel
    wf = wgt.clone().requires_grad_(True)
    ((x @ wf - t) ** 2).mean().backward()
    err = float((flat - wf.grad.reshape(-1)).abs().max())
    mx = ddp.max_over_ranks([float(rank + 1), 5.0 - rank])
    out.put((rank, err, float(wgt.sum()), mx))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_allreduce_and_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=100) for _ in ps)
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, e0, s0, m0), (r1, e1, s1, m1) = res
    assert e0 < 1e-6 and e1 < 1e-6                    # DP mean of shard gradients == global-batch gradient
    assert s0 == s1                                   # parameters identical after the broadcast
    assert m0 == m1 == [2.0, 5.0]                     # max over ranks (bench timing rule)


def _worker_buckets(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from segmentation_training_pipeline_b200 import ddp
    ddp.init(backend="gloo")
    # flat "gradient" of three layers in forward order; the late 90 % is reduced first (asynchronously), then the head --
    # the order Trainer._ddp_step enqueues them in -- and the result must equal ONE all-reduce of the whole buffer
    sizes = [(0, 8), (8, 40), (48, 120), (168, 832)]
    off = ddp.bucket_split(sizes, 0.9)
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    whole = g.clone()
    w1 = ddp.allreduce_sum_async(g[off:])
    w2 = ddp.allreduce_sum_async(g[:off])
    for w in (w1, w2):
        w.wait()
    ddp.allreduce_sum_(whole)
    # BatchNorm moving statistics (local per rank) averaged through one flat all-reduce before validation (fit.py)
    bufs = {"bn1/moving_mean": torch.full((3,), float(rank)), "bn0/moving_variance": torch.full((2, 2), 10.0 * (rank + 1))}
    ddp.sync_buffers_mean_(bufs)
    ok_bufs = bool(torch.allclose(bufs["bn1/moving_mean"], torch.full((3,), 0.5)) and
                   torch.allclose(bufs["bn0/moving_variance"], torch.full((2, 2), 15.0)))
    m = ddp.mean_over_ranks_(torch.tensor([float(rank), 4.0], dtype=torch.float64))
    ddp.barrier()
    out.put((rank, off, bool(torch.equal(g, whole)) and ok_bufs and m.tolist() == [0.5, 4.0], float(g[999])))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_allreduce_equals_single_allreduce():
    """the two-bucket (late layers first, asynchronous) gradient all-reduce of the data-parallel step, world_size 2 on gloo."""
    from segmentation_training_pipeline_b200 import ddp
    assert ddp.bucket_split([(0, 8), (8, 40), (48, 120), (168, 832)], 0.9) == 48    # largest boundary with tail >= 90 %
    assert ddp.bucket_split([(0, 8), (8, 40), (48, 120), (168, 832)], 0.5) == 168
    assert ddp.bucket_split([], 0.9) == 0 and ddp.bucket_split([(0, 10)], 0.9) == 0
    assert ddp.allreduce_sum_async(torch.zeros(4)) is None                           # single process: nothing to do
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_buckets, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=100) for _ in ps)
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, off, same, last in res:
        assert off == 48 and same and last == 999.0 * 3
