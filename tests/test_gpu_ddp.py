"""Data-parallel step on TWO GPUs against the oracle (SURVEY.md section 4 "N-rank step == oracle with per-slice BatchNorm";
VERDICT r1 item 6).  Needs >= 2 CUDA devices: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`; on a
one-GPU box the test is skipped (the driver's round-end GPU tier has one GPU; the log of the 2-GPU run is committed under
profiles/r2_ddp_2gpu_test.log)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path, precision):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import torch.distributed as dist
    from segmentation_training_pipeline_b200 import ddp
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer
    from tests.test_gpu_model import _data
    torch.cuda.set_device(rank)
    ddp.init(device=torch.device("cuda:%d" % rank))
    n, size, lr = 2, 64, 0.1
    net = SegNet("resnet18", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:%d" % rank, seed=0, loss=(1.0, 1.0, 0.0),
                 precision=precision)
    W0 = net.get_weights()
    img, mask = _data(n * world, size, size, seed=21)
    tr = Trainer(net, optimizer="SGD", lr=lr, world_size=world)
    ddp.broadcast_(net.flat_p, 0)
    tr.set_pool(img[rank * n:(rank + 1) * n], mask[rank * n:(rank + 1) * n])   # this rank's slice of the global batch
    tr.capture()                                                                # three graphs + two NCCL all-reduces (_capture_ddp)
    tr.step()
    torch.cuda.synchronize()
    flat = net.flat_p.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    in_sync = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    loss = torch.tensor([tr.loss_value()], device=flat.device, dtype=torch.float64)
    losses = [torch.zeros_like(loss) for _ in range(world)]
    dist.all_gather(losses, loss)
    if rank == 0:
        W1 = net.get_weights()
        np.savez(out_path, in_sync=in_sync, losses=np.array([float(l) for l in losses]),
                 **{"w0/" + k: v for k, v in W0.items()}, **{"w1/" + k: v for k, v in W1.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_rank_step_equals_oracle_with_per_slice_batchnorm(tmp_path, precision):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from oracle import losses as OL
    from oracle.models import SegModel
    from tests.test_gpu_model import _data
    world, n, size, lr = 2, 2, 64, 0.1
    out = str(tmp_path / "ddp.npz")
    mp.spawn(_worker, args=(world, 29731 + (precision == "bf16"), out, precision), nprocs=world, join=True)
    z = np.load(out)
    assert bool(z["in_sync"]), "ranks hold different parameters after the all-reduced step"
    W0 = {k[3:]: z[k] for k in z.files if k.startswith("w0/")}
    W1 = {k[3:]: z[k] for k in z.files if k.startswith("w1/")}
    img, mask = _data(n * world, size, size, seed=21)

    # oracle: each slice is its own tower (own BatchNorm statistics, keras multi_gpu_model), gradients averaged, one SGD step
    def oracle(storage):
        grads, losses = None, []
        for r in range(world):
            om = SegModel("Unet", "resnet18", classes=1, input_shape=(size, size, 3), storage=storage, update_moving=False)
            om.load_numpy(W0)
            y = om(img[r * n:(r + 1) * n].float())
            t = mask[r * n:(r + 1) * n].float()
            lo = OL.binary_crossentropy(t, y) + OL.dice_loss(t, y)
            lo.backward()
            losses.append(float(lo.detach()))
            g = {k: p.grad.double().numpy() / world for k, p in om.params.items()}
            grads = g if grads is None else {k: grads[k] + g[k] for k in g}
        return grads, losses

    g_ref, l_ref = oracle("fp64" if precision == "fp32" else "bf16")
    g_floor, _ = oracle("fp32")
    assert np.allclose(z["losses"], l_ref, rtol=1e-5 if precision == "fp32" else 2e-3), (z["losses"], l_ref)
    worst = ("", 0.0)
    for k, g in g_ref.items():
        ge = (W0[k].astype(np.float64) - W1[k].astype(np.float64)) / lr          # the averaged gradient the step applied
        den = np.linalg.norm(g) + 1e-30
        e = float(np.linalg.norm(ge - g) / den)
        floor = float(np.linalg.norm(g_floor[k] - g) / den)
        if e > worst[1]:
            worst = (k, e)
        if precision == "fp32":
            # W1 = W0 - lr*g in fp32: the difference quotient carries eps*|W|/lr of rounding on top of the gradient noise
            tol = max(1e-4, 5.0 * floor) + 2e-7 * float(np.linalg.norm(W0[k])) / (lr * den)
        else:
            tol = max(3e-2, 2.0 * floor) + 2e-7 * float(np.linalg.norm(W0[k])) / (lr * den)
        assert e < tol, (k, e, tol, floor)
    print("2-rank %s step: losses %s vs oracle %s; worst averaged-gradient error %s" % (precision, z["losses"], l_ref, worst))


def _comm_worker(rank, world, id_path, out_path):
    import ctypes as C
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from segmentation_training_pipeline_b200 import lib
    torch.cuda.set_device(rank)
    L = lib.Lib()
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        L.comm_unique_id(buf)
        with open(id_path + ".tmp", "wb") as f:
            f.write(bytes(buf))
        os.replace(id_path + ".tmp", id_path)
    else:
        for _ in range(600):
            if os.path.exists(id_path):
                break
            time.sleep(0.05)
    uid = open(id_path, "rb").read()
    comm = C.c_void_p()
    L.comm_init(world, rank, uid, C.byref(comm))
    x = torch.full((1 << 20,), float(rank + 1), device="cuda:%d" % rank)
    x[:8] = torch.arange(8, device=x.device, dtype=torch.float32) * (rank + 1)
    L.allreduce(comm, x.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    L.comm_destroy(comm)
    np.save(out_path % rank, x[:16].cpu().numpy())


def test_c_abi_nccl_wrappers(tmp_path):
    """stp_comm_unique_id / stp_comm_init / stp_allreduce / stp_comm_destroy (include/stp.h K13): one process per GPU, the
    id handed over out of band (a file), in-place SUM all-reduce of an fp32 buffer."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    out = str(tmp_path / "r%d.npy")
    mp.spawn(_comm_worker, args=(world, str(tmp_path / "nccl.id"), out), nprocs=world, join=True)
    want = np.full(16, 3.0, np.float32)
    want[:8] = np.arange(8) * 3.0
    for r in range(world):
        assert np.array_equal(np.load(out % r), want), r
