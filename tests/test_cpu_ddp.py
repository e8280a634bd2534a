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
