"""End-to-end GPU parity of the DeepLabV3+ / MobileNetV2 graph (the reference's in-tree impl/deeplab/model.py, architecture
`DeepLabV3` of every example config) against the CPU oracle: forward probabilities, loss, every parameter gradient, with the
engine's dropout mask reproduced by the oracle; inference mode; a short training run under the whole-step CUDA graph.
Tolerance model as tests/test_gpu_model.py: the engine must be as close to the bf16-storage oracle as bf16 storage itself is
to fp32 (the deviation between the two oracles is the noise floor)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_model import _data, _perturb

pytestmark = pytest.mark.gpu


def _build(n, size, classes, loss, activation, dropout=None, channels=3):
    from segmentation_training_pipeline_b200.models import SegNet
    return SegNet("mobilenetv2", classes=classes, input_shape=(size, size, channels), batch=n, device="cuda:0", seed=0, loss=loss,
                  architecture="DeepLabV3", activation=activation, dropout=dropout)


def _masks(mask, classes, onehot, size):
    if classes > 1:
        mask = torch.cat([torch.roll(mask, c * size // 8, dims=2) for c in range(classes)], dim=3).contiguous()
    if onehot:
        lab = torch.zeros(mask.shape[:3], dtype=torch.long)
        for c in range(classes - 1, 0, -1):
            lab[mask[..., c] > 0] = c
        mask = torch.nn.functional.one_hot(lab, classes).to(torch.uint8).contiguous()
    return mask


@pytest.mark.parametrize("size,classes,loss,activation,dropout,n", [
    (64, 1, (1.0, 1.0, 0.0), "sigmoid", 0.1, 4),
    (96, 2, (1.0, 0.0, 0.0), "sigmoid", 0.0, 4),
    (64, 3, (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0), "softmax", 0.1, 4),
    (320, 1, (1.0, 0.0, 0.0), "sigmoid", 0.1, 10),     # the reference's example experiment as written: shape 320, batch 10, binary_crossentropy
])
def test_deeplab_forward_backward_parity(cuda, size, classes, loss, activation, dropout, n):
    from oracle import losses as OL
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200 import lib
    from segmentation_training_pipeline_b200.trainer import Trainer

    onehot = activation == "softmax"
    net = _build(n, size, classes, loss, activation, dropout)
    W = _perturb(net.get_weights())
    net.set_weights(W)
    tr = Trainer(net)
    img, mask = _data(n, size, size)
    mask = _masks(mask, classes, onehot, size)
    tr.set_batch(img.cuda(), mask.cuda())
    net.d_step.fill_(5)
    net.prep_weights()
    net.forward()
    net.backward()
    torch.cuda.synchronize()
    res = net.loss.result.cpu().numpy()
    lg = net.head.logits.cpu().view(n, size, size, classes)
    prob = torch.sigmoid(lg) if activation == "sigmoid" else torch.softmax(lg, dim=-1)
    grads = net.get_grads()

    def run(storage):
        om = SegModel("DeepLabV3", "mobilenetv2", classes=classes, activation=activation, input_shape=(size, size, 3), storage=storage,
                      update_moving=False, dropout=(dropout, net.seed, 0xD0, 5) if dropout else None)
        assert set(om.params.keys()) == set(net.params.keys()), sorted(set(om.params.keys()) ^ set(net.params.keys()))
        om.load_numpy(W)
        y = om(img.float())
        t = mask.float()
        if onehot:
            lo = loss[6] * OL.categorical_crossentropy(t, y)
        else:
            lo = loss[0] * OL.binary_crossentropy(t, y) + loss[1] * OL.dice_loss(t, y) + loss[2] * OL.iou_loss(t, y)
        lo.backward()
        return y.detach(), float(lo.detach()), {k: p.grad.numpy().copy() for k, p in om.params.items()}

    y, lo, go_all = run("bf16")
    y32, lo32, go32 = run("fp32")
    floor = float((y - y32).norm() / y32.norm())
    err = float((prob - y).norm() / y.norm())
    print("probabilities rel err", err, "bf16-vs-fp32 floor", floor, "loss", float(res[lib.L_LOSS]), lo, lo32)
    assert err < max(5e-3, 0.8 * floor)
    # bf16 realisation noise of the loss: the same graph run with different (equally exact) kernel selections for its 1x1 layers
    # lands on either side of the bf16 oracle -- 0.60428 / 0.60683 / within 1e-3, oracle 0.60551, fp32 oracle 0.60521 at 320^2 batch
    # 10 (profiles/r2_s10_deeplab_loss_realisations.txt) -- while the probabilities stay at 0.36x the bf16-vs-fp32 floor in every
    # case.  The bound is therefore the probability floor carried to the loss (|dL| <= mean|dL/dp| * |dp|, a few 1e-3 here), not
    # 1e-3; the graph's SEMANTICS are held to 2e-7 on the loss by the fp32 parity mode (test_deeplab_fp32_parity_mode).
    assert abs(float(res[lib.L_LOSS]) - lo) < max(2.5e-3 * max(1.0, abs(lo)), 1.5 * abs(lo - lo32))
    worst = ("", 0.0)
    for k, go in go_all.items():
        ge = grads[k]
        assert go.shape == ge.shape, (k, go.shape, ge.shape)
        e = float(np.linalg.norm(ge - go) / (np.linalg.norm(go) + 1e-12))
        fl = float(np.linalg.norm(go - go32[k]) / (np.linalg.norm(go32[k]) + 1e-12))
        if fl > 0.25:     # gradients that are near-total cancellations (e.g. a bias / beta in front of a BatchNorm)
            continue
        if e / max(fl, 1e-3) > worst[1]:
            worst = (k, e / max(fl, 1e-3))
        assert e < max(2e-2, 1.5 * fl), (k, e, fl)
    print("worst grad err / bf16 floor", worst)


def test_deeplab_weight_names_match_keras_layers(cuda):
    """parameter names / Keras layouts of impl/deeplab/model.py (weights exchangeable by name with model.load_weights(by_name))"""
    net = _build(2, 64, 1, (1.0, 0.0, 0.0), "sigmoid")
    w = net.get_weights()
    assert w["Conv/kernel"].shape == (3, 3, 3, 32)
    assert w["expanded_conv_depthwise/depthwise_kernel"].shape == (3, 3, 32, 1)
    assert w["expanded_conv_1_expand/kernel"].shape == (1, 1, 16, 96)
    assert w["expanded_conv_16_project/kernel"].shape == (1, 1, 960, 320)
    assert w["image_pooling/kernel"].shape == (1, 1, 320, 256)
    assert w["concat_projection/kernel"].shape == (1, 1, 512, 256)
    assert w["custom_logits_semantic/kernel"].shape == (1, 1, 256, 1) and w["custom_logits_semantic/bias"].shape == (1,)
    assert w["Conv_BN/moving_variance"].shape == (32,)
    assert sum(v.size for k, v in w.items() if not k.startswith(("Conv_BN/moving", )) and "moving_" not in k) == 2108417 + 0
    w2 = {k: (v + 0.01).astype(np.float32) for k, v in w.items()}
    net.set_weights(w2)
    w3 = net.get_weights()
    for k in w2:
        assert np.array_equal(w2[k], w3[k]), k
    enc = set(net.encoder_param_names)
    assert "expanded_conv_16_project_BN/beta" in enc and "aspp0/kernel" not in enc


def test_deeplab_inference_matches_oracle(cuda):
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size = 2, 64
    net = _build(n, size, 1, (1.0, 0.0, 0.0), "sigmoid")
    rng = np.random.default_rng(3)
    W = _perturb(net.get_weights())
    for k in W:
        if k.endswith("moving_mean"):
            W[k] = rng.normal(0, 0.2, W[k].shape).astype(np.float32)
        if k.endswith("moving_variance"):
            W[k] = rng.uniform(0.5, 1.5, W[k].shape).astype(np.float32)
    net.set_weights(W)
    tr = Trainer(net)
    img, mask = _data(n, size, size)
    tr.set_batch(img.cuda(), mask.cuda())
    net.training = False
    net.prep_weights()
    net.forward()
    torch.cuda.synchronize()
    prob = torch.sigmoid(net.head.logits.cpu().view(n, size, size, 1))
    outs = {}
    for storage in ("bf16", "fp32"):
        om = SegModel("DeepLabV3", "mobilenetv2", classes=1, input_shape=(size, size, 3), storage=storage)
        om.load_numpy(W)
        om.training = False
        with torch.no_grad():
            outs[storage] = om(img.float())
    floor = float((outs["bf16"] - outs["fp32"]).norm() / outs["fp32"].norm())
    err = float((prob - outs["bf16"]).norm() / outs["bf16"].norm())
    print("inference rel err", err, "floor", floor)
    assert err < max(5e-3, 1.5 * floor)


def test_deeplab_trains_under_cuda_graph(cuda):
    """the reference example's shape of run (architecture DeepLabV3, backbone mobilenetv2, binary_crossentropy, Adam) at small
    size: loss decreases; whole-step CUDA graph replay == eager (the dropout mask follows the device step counter)."""
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size, steps = 4, 64, 30
    img, mask = _data(n * 2, size, size, seed=3)
    curves = []
    for graph in (True, False):
        net = _build(n, size, 1, (1.0, 0.0, 0.0), "sigmoid")
        tr = Trainer(net, optimizer="Adam", lr=1e-3)
        tr.set_pool(img, mask)
        if graph:
            tr.capture()
        c = []
        for s in range(steps if graph else 6):
            tr.step()
            c.append(tr.loss_value())
        curves.append(c)
    print(curves[0])
    assert curves[0][-1] < 0.8 * curves[0][0]
    assert np.isfinite(curves[0]).all()
    for a, b in zip(curves[0][:6], curves[1]):
        assert abs(a - b) < 1e-5 * max(1.0, abs(b)), (curves[0][:6], curves[1])


def test_deeplab_depth_profile(cuda):
    """Layer by layer along the MobileNetV2 body (scripts/deeplab_diag.py, profiles/r2_deeplab_depth_profile.txt): the first layers
    match the bf16-storage oracle to rounding, and at EVERY tapped layer the engine is closer to the bf16 oracle than bf16 storage
    itself is to fp32 -- the deviation at the output is the depth-amplified bf16 noise of this random-init network (bf16 vs fp32
    oracle: 0.4 % after the stem, ~30 % after block 16), not a step at some layer."""
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size = 4, 64
    net = _build(n, size, 1, (1.0, 0.0, 0.0), "sigmoid", dropout=0.0)
    W = _perturb(net.get_weights())
    net.set_weights(W)
    tr = Trainer(net)
    img, mask = _data(n, size, size)
    tr.set_batch(img.cuda(), mask.cuda())
    net.prep_weights()
    net.forward()
    torch.cuda.synchronize()
    oms = {}
    for st in ("bf16", "fp32"):
        om = SegModel("DeepLabV3", "mobilenetv2", classes=1, input_shape=(size, size, 3), storage=st, update_moving=False)
        om.load_numpy(W)
        with torch.no_grad():
            om(img.float())
        oms[st] = om
    seen = 0
    for name, b in oms["bf16"].taps.items():
        if name not in net.bufs or net.bufs[name].torch().shape != b.permute(0, 2, 3, 1).shape:
            continue
        e = net.bufs[name].torch().float().cpu()
        b = b.permute(0, 2, 3, 1)
        f = oms["fp32"].taps[name].permute(0, 2, 3, 1)
        err, floor = float((e - b).norm() / b.norm()), float((b - f).norm() / f.norm())
        assert err < max(2e-3, 0.7 * floor), (name, err, floor)
        seen += 1
    assert seen >= 18
    first = net.bufs["Conv_Relu6"].torch().float().cpu()
    assert float((first - oms["bf16"].taps["Conv_Relu6"].permute(0, 2, 3, 1)).norm() / first.norm()) < 1e-3


@pytest.mark.parametrize("size,n,classes,activation,dropout", [(64, 4, 1, "sigmoid", 0.1), (96, 4, 3, "softmax", 0.0)])
def test_deeplab_fp32_parity_mode(cuda, size, n, classes, activation, dropout):
    """PARITY MODE of the DeepLabV3 graph (SegNet(precision="fp32"): fp32 activations / gradients / weights, CUDA-core kernels
    with double accumulators behind the same C-ABI entry points, csrc/f32_path.cu + f32_deeplab.cu).  Anchor: the oracle in DOUBLE
    precision.  The output probabilities within 1e-4 relative L2 (or twice the fp32 oracle's own distance from the anchor), the
    loss within 1e-5, every parameter gradient as close to the anchor as the fp32 oracle is (each tensor within x5, median ratio
    below 2) -- where bf16 storage is 30-40 % away (test_deeplab_depth_profile)."""
    from oracle import losses as OL
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200 import lib
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer

    onehot = activation == "softmax"
    loss = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0) if onehot else (1.0, 1.0, 0.0)
    net = SegNet("mobilenetv2", classes=classes, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=loss,
                 architecture="DeepLabV3", activation=activation, dropout=dropout, precision="fp32")
    W = _perturb(net.get_weights())
    net.set_weights(W)
    tr = Trainer(net)
    img, mask = _data(n, size, size)
    mask = _masks(mask, classes, onehot, size)
    tr.set_batch(img.cuda(), mask.cuda())
    net.d_step.fill_(3)
    before = net.L.tc_launch_count()
    net.prep_weights()
    net.forward()
    net.backward()
    torch.cuda.synchronize()
    assert net.L.tc_launch_count() == before      # no bf16 tensor-core kernel ran
    res = net.loss.result.cpu().numpy()
    lg = net.head.logits.cpu().view(n, size, size, classes).double()
    prob = torch.sigmoid(lg) if activation == "sigmoid" else torch.softmax(lg, dim=-1)
    grads = net.get_grads()

    def run(storage):
        om = SegModel("DeepLabV3", "mobilenetv2", classes=classes, activation=activation, input_shape=(size, size, 3), storage=storage,
                      update_moving=False, dropout=(dropout, net.seed, 0xD0, 3) if dropout else None)
        om.load_numpy(W)
        y = om(img.float())
        t = mask.float()
        lo = loss[6] * OL.categorical_crossentropy(t, y) if onehot else OL.binary_crossentropy(t, y) + OL.dice_loss(t, y)
        lo.backward()
        return y.detach().double(), float(lo.detach()), {k: p.grad.double().numpy().copy() for k, p in om.params.items()}

    y64, lo64, g64 = run("fp64")
    y32, lo32, g32 = run("fp32")
    err = float((prob - y64).norm() / y64.norm())
    err32 = float((y32 - y64).norm() / y64.norm())
    print("DeepLabV3 fp32 parity mode vs fp64 anchor: probabilities rel err %.3e (fp32 oracle %.3e), loss %.7f vs %.7f" %
          (err, err32, float(res[lib.L_LOSS]), lo64))
    assert err < max(1e-4, 2.0 * err32)
    assert abs(float(res[lib.L_LOSS]) - lo64) < 1e-5 * max(1.0, abs(lo64))
    worst, ratios = ("", 0.0, 0.0), []
    for k, go in g64.items():
        ge = grads[k].astype(np.float64)
        assert go.shape == ge.shape, (k, go.shape, ge.shape)
        if np.linalg.norm(go) < 1e-9 * go.size ** 0.5:
            continue
        den = np.linalg.norm(go) + 1e-30
        e, floor = float(np.linalg.norm(ge - go) / den), float(np.linalg.norm(g32[k] - go) / den)
        if e > worst[1]:
            worst = (k, e, floor)
        ratios.append(e / max(floor, 1e-7))
        assert e < max(2e-4, 5.0 * floor), (k, e, floor)
    print("worst gradient: %s engine-vs-fp64 %.3e, fp32-oracle-vs-fp64 %.3e; median engine/oracle error ratio %.2f" %
          (worst + (float(np.median(ratios)),)))
    assert float(np.median(ratios)) < 2.0, float(np.median(ratios))


@pytest.mark.parametrize("lr,mom", [(3e-3, 0.9)])
def test_deeplab_loss_curve_fp32_parity_mode(cuda, lr, mom):
    """Loss curve of the DeepLabV3 graph in the fp32 parity mode: 40 SGD-momentum steps under the whole-step CUDA graph with
    Dropout(0.1) ON (the oracle regenerates each step's mask from the device step counter) against the fp32 and fp64 oracles from
    the same weights and batches.  The first step (pure forward parity) agrees to 1e-7.  After that THIS run is chaotic for any fp32
    implementation: the fp32 ORACLE is 4e-3 from the fp64 oracle at step 2 and 8e-2 by step 40 (batch 4: the image-pooling
    BatchNorm normalises over four values per channel, and every block re-normalises), so north_star's 1e-3 cannot be asserted
    between two fp32 runs of it; asserted instead: the engine stays as close to the fp64 anchor as the fp32 oracle does (x2).
    Measured: engine 6.6e-2, fp32 oracle 8.2e-2 (lr 3e-3, momentum 0.9); 6.1e-2 vs 9.8e-2 with plain SGD lr 1e-3."""
    from oracle import losses as OL, optim as OO
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer

    n, size, steps, pool = 4, 64, 40, 8
    net = SegNet("mobilenetv2", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(1.0, 0.0, 0.0),
                 architecture="DeepLabV3", precision="fp32")
    W = net.get_weights()
    img, mask = _data(pool, size, size, seed=11)
    tr = Trainer(net, optimizer="SGD", lr=lr, momentum=mom)
    tr.set_pool(img, mask)
    tr.capture()
    curve, dsteps = [], []
    for s in range(steps):
        dsteps.append(int(net.d_step.item()))
        tr.step()
        curve.append(tr.loss_value())

    def oracle_curve(storage):
        om = SegModel("DeepLabV3", "mobilenetv2", classes=1, input_shape=(size, size, 3), storage=storage)
        om.load_numpy(W)
        opt = OO.SGD(om.params, lr=lr, momentum=mom)
        out = []
        for s in range(steps):
            idx = [(s * n + j) % pool for j in range(n)]
            om.dropout = (0.1, net.seed, 0xD0, dsteps[s])
            y = om(img[idx].float())
            lo = OL.binary_crossentropy(mask[idx].float(), y)
            for p in om.params.values():
                p.grad = None
            lo.backward()
            opt.step({k: p.grad for k, p in om.params.items()})
            out.append(float(lo.detach()))
        return np.array(out)

    c, r32, r64 = np.array(curve), oracle_curve("fp32"), oracle_curve("fp64")
    dev = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d_engine, d_oracle, d_pair = dev(c, r64), dev(r32, r64), dev(c, r32)
    print("engine fp32", np.round(c[::5], 5))
    print("oracle fp32", np.round(r32[::5], 5))
    print("oracle fp64", np.round(r64[::5], 5))
    print("max deviation from the fp64 anchor: engine %.3e, fp32 oracle %.3e; engine vs fp32 oracle %.3e" %
          (d_engine.max(), d_oracle.max(), d_pair.max()))
    print("per-step deviation engine vs fp32 oracle", np.round(d_pair[:6], 6), "fp32 oracle vs fp64", np.round(d_oracle[:6], 6))
    assert d_pair[0] < 1e-5, d_pair[0]                 # the first step is pure forward parity
    assert d_engine.max() < max(1e-3, 2.0 * d_oracle.max()), (d_engine.max(), d_oracle.max())
    assert d_pair.max() < max(1e-3, 2.0 * d_oracle.max()), (d_pair.max(), d_oracle.max())
    assert c[-5:].mean() < c[:3].mean()


@pytest.mark.parametrize("OS", [16, 8])
def test_deeplab_xception_fp32_parity_mode(cuda, OS):
    """The other backbone of the reference's in-tree model (impl/deeplab/model.py:339-383: modified aligned Xception, 41 M
    parameters; ASPP with three atrous separable branches, decoder with the 1/4-resolution skip): fp32 parity mode against the fp64
    oracle -- probabilities, loss, every parameter gradient -- for output stride 16 (schema default) and 8."""
    from oracle import losses as OL
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200 import lib
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer

    n, size, dropout = 2, 64, 0.1
    net = SegNet("xception", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(1.0, 1.0, 0.0),
                 architecture="DeepLabV3", dropout=dropout, precision="fp32", OS=OS)
    W = _perturb(net.get_weights())
    net.set_weights(W)
    tr = Trainer(net)
    img, mask = _data(n, size, size)
    tr.set_batch(img.cuda(), mask.cuda())
    net.prep_weights()
    net.forward()
    net.backward()
    torch.cuda.synchronize()
    res = net.loss.result.cpu().numpy()
    prob = torch.sigmoid(net.head.logits.cpu().view(n, size, size, 1).double())
    grads = net.get_grads()

    def run(storage):
        om = SegModel("DeepLabV3", "xception", classes=1, input_shape=(size, size, 3), storage=storage, update_moving=False,
                      dropout=(dropout, net.seed, 0xD0, 0), OS=OS)
        assert set(om.params.keys()) == set(net.params.keys())
        om.load_numpy(W)
        y = om(img.float())
        lo = OL.binary_crossentropy(mask.float(), y) + OL.dice_loss(mask.float(), y)
        lo.backward()
        return y.detach().double(), float(lo.detach()), {k: p.grad.double().numpy().copy() for k, p in om.params.items()}

    y64, lo64, g64 = run("fp64")
    y32, lo32, g32 = run("fp32")
    err, err32 = float((prob - y64).norm() / y64.norm()), float((y32 - y64).norm() / y64.norm())
    print("xception OS%d fp32 parity mode vs fp64 anchor: probabilities rel err %.3e (fp32 oracle %.3e), loss %.7f vs %.7f" %
          (OS, err, err32, float(res[lib.L_LOSS]), lo64))
    assert err < max(1e-4, 2.0 * err32)
    assert abs(float(res[lib.L_LOSS]) - lo64) < 1e-5 * max(1.0, abs(lo64))
    ratios, worst = [], ("", 0.0, 0.0)
    for k, go in g64.items():
        ge = grads[k].astype(np.float64)
        if np.linalg.norm(go) < 1e-9 * go.size ** 0.5:
            continue
        den = np.linalg.norm(go) + 1e-30
        e, floor = float(np.linalg.norm(ge - go) / den), float(np.linalg.norm(g32[k] - go) / den)
        if e > worst[1]:
            worst = (k, e, floor)
        ratios.append(e / max(floor, 1e-7))
        assert e < max(2e-4, 5.0 * floor), (k, e, floor)
    print("worst gradient: %s engine-vs-fp64 %.3e, fp32-oracle-vs-fp64 %.3e; median ratio %.2f" % (worst + (float(np.median(ratios)),)))
    assert float(np.median(ratios)) < 2.0


def test_deeplab_xception_bf16_step(cuda):
    """product path (bf16, tcgen05 where the shapes allow): forward / loss against the bf16-storage oracle within the bf16 floor,
    and a few Adam steps under the CUDA graph reduce the loss"""
    from oracle import losses as OL
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200 import lib
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size = 4, 64
    net = SegNet("xception", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(1.0, 0.0, 0.0),
                 architecture="DeepLabV3", dropout=0.0)
    W = _perturb(net.get_weights())
    net.set_weights(W)
    tr = Trainer(net, optimizer="Adam", lr=1e-3)
    img, mask = _data(n, size, size)
    tr.set_batch(img.cuda(), mask.cuda())
    net.prep_weights()
    net.forward()
    torch.cuda.synchronize()
    prob = torch.sigmoid(net.head.logits.cpu().view(n, size, size, 1))
    loss0 = float(net.loss.result.cpu()[lib.L_LOSS])
    ys = {}
    for st in ("bf16", "fp32"):
        om = SegModel("DeepLabV3", "xception", classes=1, input_shape=(size, size, 3), storage=st, update_moving=False)
        om.load_numpy(W)
        with torch.no_grad():
            ys[st] = om(img.float())
    floor = float((ys["bf16"] - ys["fp32"]).norm() / ys["fp32"].norm())
    err = float((prob - ys["bf16"]).norm() / ys["bf16"].norm())
    lo = float(OL.binary_crossentropy(mask.float(), ys["bf16"]))
    print("xception bf16: probabilities rel err %.4f, bf16-vs-fp32 floor %.4f, loss %.5f vs %.5f" % (err, floor, loss0, lo))
    assert err < max(5e-3, 0.8 * floor)
    tr.set_pool(img, mask)
    tr.capture()
    c = []
    for _ in range(25):
        tr.step()
        c.append(tr.loss_value())
    assert np.isfinite(c).all() and c[-1] < 0.9 * c[0], c
