"""End-to-end plumbing on the GPU through the reference-facing API: segmentation.parse(cfg).fit(SimplePNGMaskDataSet)
writes the files the reference writes (weights/, metrics/, summary.yaml), refuses to re-run a finished experiment,
and the training loss goes down."""
import csv
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _make_dataset(root, n=8, size=64):
    import cv2
    os.makedirs(os.path.join(root, "img"))
    os.makedirs(os.path.join(root, "mask"))
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:size, 0:size]
    for k in range(n):
        cy, cx, r = rng.integers(size // 4, 3 * size // 4, 2).tolist() + [size // 5]
        m = (((yy - cy) ** 2 + (xx - cx) ** 2) < r * r).astype(np.uint8)
        img = (rng.integers(0, 120, (size, size, 3)) + m[:, :, None] * 120).astype(np.uint8)
        cv2.imwrite(os.path.join(root, "img", "%02d.png" % k), img)
        cv2.imwrite(os.path.join(root, "mask", "%02d.png" % k), m * 255)


def test_fit_writes_reference_artifacts(cuda, tmp_path):
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet
    _make_dataset(str(tmp_path))
    cfgp = str(tmp_path / "exp" / "config.yaml")
    os.makedirs(os.path.dirname(cfgp))
    shutil.copy(os.path.join(HERE, "golden", "configs", "c1_plumbing.yaml"), cfgp)
    cfg = segmentation.parse(cfgp)
    ds = SimplePNGMaskDataSet(str(tmp_path / "img"), str(tmp_path / "mask"))
    res = cfg.fit(ds)
    exp = os.path.dirname(cfgp)
    assert os.path.exists(os.path.join(exp, "summary.yaml"))
    assert len(res) == 4                                  # 2 folds x 2 stages
    for fold in range(2):
        for stage, epochs in ((0, 2), (1, 1)):
            assert os.path.exists(os.path.join(exp, "weights", "best-%d.%d.weights.npz" % (fold, stage)))
            rows = list(csv.DictReader(open(os.path.join(exp, "metrics", "metrics-%d.%d.csv" % (fold, stage)))))
            assert len(rows) == epochs
            for r in rows:
                for k in ("loss", "val_loss", "binary_accuracy", "val_binary_accuracy", "iou", "val_iou", "dice"):
                    assert np.isfinite(float(r[k])), (k, r)
                assert 0.0 <= float(r["val_binary_accuracy"]) <= 1.0
    with pytest.raises(ValueError, match="already finished"):
        cfg.fit(ds)
    net = cfg.load_model(0, 1)
    assert net.get_weights()["conv0/kernel"].shape == (7, 7, 3, 64)
    assert len(cfg.info()) == 4
    # ---- inference verbs on the trained folds (reference segmentation.py:62-91): flip-TTA, fold ensembling, scale back
    import cv2
    big = str(tmp_path / "big")
    os.makedirs(big)
    src = cv2.imread(str(tmp_path / "img" / "00.png"))
    cv2.imwrite(os.path.join(big, "a.png"), cv2.resize(src, (96, 80)))          # not the network resolution
    cv2.imwrite(os.path.join(big, "b.png"), src)
    out = str(tmp_path / "pred")
    assert cfg.predict_to_directory(big, out, fold=[0, 1], stage=1, ttflips=True) == 2
    pa, pb = cv2.imread(os.path.join(out, "a.png"), cv2.IMREAD_GRAYSCALE), cv2.imread(os.path.join(out, "b.png"), cv2.IMREAD_GRAYSCALE)
    assert pa.shape == (80, 96) and pb.shape == (64, 64) and pa.dtype == np.uint8   # arr*255: an 8-bit probability image
    seen = {}
    cfg.predict_in_directory(big, 0, 1, lambda id_, m, data: data.__setitem__(id_, m.shape), seen, ttflips=False)
    assert seen == {"a.png": (80, 96, 1), "b.png": (64, 64, 1)}
    # ttflips is an average over flips: a horizontally flipped image gives the flipped prediction
    from segmentation_training_pipeline_b200.predict import predict_arrays
    x = np.stack([cv2.cvtColor(src, cv2.COLOR_BGR2RGB)] * 2)
    p0 = predict_arrays(net, x, ttflips=True)
    p1 = predict_arrays(net, x[:, :, ::-1], ttflips=True)
    assert np.allclose(p0, p1[:, :, ::-1], atol=2e-3)
    batches = list(cfg.evaluateAll(ds, fold=0, stage=1))
    assert sum(len(b.data) for b in batches) == len(ds) and batches[0].results[0].shape == (64, 64, 1)


def test_lovasz_step_in_cuda_graph(cuda):
    """lovasz_loss (radix sort + scan + Jaccard-gradient kernels) inside the captured training step: loss goes down."""
    import torch
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size = 4, 64
    g = torch.Generator().manual_seed(0)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    mask = (((yy - 30) ** 2 + (xx - 34) ** 2) < 220).to(torch.uint8)[None, :, :, None].repeat(n, 1, 1, 1)
    img = (torch.randint(0, 100, (n, size, size, 3), generator=g) + mask * 120).to(torch.uint8)
    net = SegNet("resnet18", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(0.0, 0.0, 0.0, 1.0))
    tr = Trainer(net, optimizer="Adam", lr=1e-3)
    tr.set_pool(img, mask)
    tr.capture()
    losses = []
    for _ in range(30):
        tr.step()
        losses.append(tr.loss_value())
    assert all(np.isfinite(losses)) and losses[-1] < 0.7 * losses[0], losses


def test_loss_decreases_on_a_fixed_batch(cuda):
    import torch
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size = 4, 64
    g = torch.Generator().manual_seed(0)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    mask = (((yy - 32) ** 2 + (xx - 28) ** 2) < 200).to(torch.uint8)[None, :, :, None].repeat(n, 1, 1, 1)
    img = (torch.randint(0, 100, (n, size, size, 3), generator=g) + mask * 120).to(torch.uint8)
    net = SegNet("resnet18", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(1.0, 1.0, 0.0))
    tr = Trainer(net, optimizer="Adam", lr=1e-3)
    tr.enable_host_feed()
    hi, hm = img.pin_memory(), mask.pin_memory()
    losses = [tr.step_from_host(hi, hm)["loss"] for _ in range(30)]
    assert all(np.isfinite(losses))
    assert losses[-1] < 0.7 * losses[0], losses


def test_fit_config0_unet_vgg16(cuda, tmp_path):
    """BASELINE.json configs[0] as written: 4 synthetic 128x128 png pairs, U-Net/VGG16, 2 folds x 2 epochs."""
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet
    _make_dataset(str(tmp_path), n=4, size=128)
    cfgp = str(tmp_path / "exp" / "config.yaml")
    os.makedirs(os.path.dirname(cfgp))
    shutil.copy(os.path.join(HERE, "golden", "configs", "c1_unet_vgg16.yaml"), cfgp)
    cfg = segmentation.parse(cfgp)
    res = cfg.fit(SimplePNGMaskDataSet(str(tmp_path / "img"), str(tmp_path / "mask")))
    exp = os.path.dirname(cfgp)
    assert len(res) == 2 and os.path.exists(os.path.join(exp, "summary.yaml"))
    for fold in range(2):
        assert os.path.exists(os.path.join(exp, "weights", "best-%d.0.weights.npz" % fold))
        rows = list(csv.DictReader(open(os.path.join(exp, "metrics", "metrics-%d.0.csv" % fold))))
        assert len(rows) == 2 and all(np.isfinite(float(r["loss"])) and np.isfinite(float(r["val_loss"])) for r in rows)
    w = cfg.load_model(0, 0).get_weights()
    assert w["block1_conv1/kernel"].shape == (3, 3, 3, 64) and w["block5_conv3/bias"].shape == (512,)


def test_fit_linknet_resnet34_cyclic_lr(cuda, tmp_path):
    """Linknet / ResNet-34 through the YAML surface with the full device augmentation block and a CyclicLR callback."""
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet
    _make_dataset(str(tmp_path), n=8, size=64)
    cfgp = str(tmp_path / "exp" / "config.yaml")
    os.makedirs(os.path.dirname(cfgp))
    shutil.copy(os.path.join(HERE, "golden", "configs", "c5_linknet_resnet34.yaml"), cfgp)
    cfg = segmentation.parse(cfgp)
    res = cfg.fit(SimplePNGMaskDataSet(str(tmp_path / "img"), str(tmp_path / "mask")))
    assert len(res) == 2
    rows = list(csv.DictReader(open(os.path.join(os.path.dirname(cfgp), "metrics", "metrics-0.0.csv"))))
    assert len(rows) == 2 and all(np.isfinite(float(r["loss"])) for r in rows)
    lrs = [float(r["lr"]) for r in rows]
    assert all(0.0005 - 1e-9 <= v <= 0.002 + 1e-9 for v in lrs) and lrs[0] != lrs[1]
    w = cfg.load_model(0, 0).get_weights()
    assert "decoder_stage0_conv3/kernel" in w and w["decoder_stage4_conv3/kernel"].shape == (1, 1, 16, 16)


class _ThreeClassDisks:
    """Dataset protocol (README.md:417-427): __len__, __getitem__ -> PredictionItem(id, x uint8 HxWx3, y uint8 HxWx3)."""

    def __init__(self, n=8, size=128):
        from segmentation_pipeline.impl.datasets import PredictionItem
        rng = np.random.default_rng(7)
        yy, xx = np.mgrid[0:size, 0:size]
        self.items = []
        for k in range(n):
            y = np.zeros((size, size, 3), np.uint8)
            x = rng.integers(0, 90, (size, size, 3)).astype(np.uint8)
            for c in range(3):
                cy, cx = rng.integers(size // 4, 3 * size // 4, 2).tolist()
                y[:, :, c] = ((yy - cy) ** 2 + (xx - cx) ** 2) < (size // 6) ** 2
                x[:, :, c] += y[:, :, c] * 150
            self.items.append(PredictionItem("s%02d" % k, x, y))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def test_fit_config2_fpn_resnet50_lovasz_3class(cuda, tmp_path):
    """BASELINE.json configs[2] through the YAML surface (shape / batch / folds shrunk): FPN / ResNet-50, 3-class masks,
    lovasz_loss on logits with `activation: softmax` shaping the predictions."""
    import yaml
    from segmentation_pipeline import segmentation
    spec = yaml.safe_load(open(os.path.join(HERE, "golden", "configs", "c3_fpn_resnet50.yaml")))
    spec.update({"shape": [128, 128, 3], "batch": 2, "folds_count": 2, "stages": [{"epochs": 3}]})
    cfgp = str(tmp_path / "exp" / "config.yaml")
    os.makedirs(os.path.dirname(cfgp))
    yaml.safe_dump(spec, open(cfgp, "w"))
    cfg = segmentation.parse(cfgp)
    ds = _ThreeClassDisks()
    res = cfg.fit(ds)
    assert len(res) == 2
    rows = list(csv.DictReader(open(os.path.join(os.path.dirname(cfgp), "metrics", "metrics-0.0.csv"))))
    assert len(rows) == 3 and all(np.isfinite(float(r["loss"])) and np.isfinite(float(r["val_loss"])) for r in rows)
    assert float(rows[-1]["loss"]) < float(rows[0]["loss"])
    net = cfg.load_model(0, 0)
    w = net.get_weights()
    assert w["head_conv/kernel"].shape == (3, 3, 512, 3) and w["pyramid_stage_0_conv1x1/kernel"].shape == (1, 1, 2048, 256)
    from segmentation_training_pipeline_b200.predict import predict_arrays
    p = predict_arrays(net, np.stack([ds[0].x, ds[1].x]))
    assert p.shape == (2, 128, 128, 3) and np.allclose(p.sum(axis=-1), 1.0, atol=1e-5)


def test_fit_training_controls(cuda, tmp_path):
    """freeze_encoder / unfreeze_encoder, negatives / validation_negatives, initial_weights, extra_train_data,
    setAllowResume and lr_find through the YAML / API surface (reference README.md:281-304, 385-415, 455-470; FAQ.md:3-12, 50-61)."""
    import cv2
    import yaml
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet
    _make_dataset(str(tmp_path), n=10, size=64)
    for k in (1, 4, 7):   # three negative examples
        cv2.imwrite(str(tmp_path / "mask" / ("%02d.png" % k)), np.zeros((64, 64), np.uint8))
    extra_root = tmp_path / "extra"
    _make_dataset(str(extra_root), n=2, size=64)
    ds = SimplePNGMaskDataSet(str(tmp_path / "img"), str(tmp_path / "mask"))
    segmentation.extra_train["more"] = SimplePNGMaskDataSet(str(extra_root / "img"), str(extra_root / "mask"))
    spec = {"architecture": "Unet", "backbone": "resnet18", "classes": 1, "activation": "sigmoid", "shape": [64, 64, 3],
            "batch": 2, "folds_count": 2, "optimizer": "Adam", "lr": 0.001, "loss": "binary_crossentropy+dice_loss",
            "metrics": ["binary_accuracy"], "primary_metric": "val_loss", "freeze_encoder": True, "extra_train_data": "more",
            "stages": [{"epochs": 1, "negatives": "none", "validation_negatives": "real"},
                       {"epochs": 1, "unfreeze_encoder": True, "negatives": 1, "initial_weights": "./weights/best-0.0.weights"}]}
    cfgp = str(tmp_path / "exp" / "config.yaml")
    os.makedirs(os.path.dirname(cfgp))
    yaml.safe_dump(spec, open(cfgp, "w"))
    cfg = segmentation.parse(cfgp)
    w0 = cfg.createNet().get_weights()            # initial weights (same seed as the folds' networks)
    res = cfg.fit(ds)
    assert len(res) == 4
    exp = os.path.dirname(cfgp)
    wa = dict(np.load(os.path.join(exp, "weights", "best-0.0.weights.npz")))
    wb = dict(np.load(os.path.join(exp, "weights", "best-0.1.weights.npz")))
    enc = ["conv0/kernel", "stage2_unit1_conv1/kernel", "bn0/gamma", "stage4_unit2_bn1/beta"]
    dec = ["decoder_stage0_conv1/kernel", "final_conv/kernel"]
    for k in enc:
        assert np.array_equal(wa[k], w0[k]), k            # frozen encoder: bit-identical after stage 0
        assert not np.array_equal(wb[k], w0[k]), k        # unfrozen in stage 1
    for k in dec:
        assert not np.array_equal(wa[k], w0[k]), k
    # resume: nothing is re-run once weights + metrics of every (fold, stage) exist
    os.remove(os.path.join(exp, "summary.yaml"))
    stamp = os.path.getmtime(os.path.join(exp, "weights", "best-1.1.weights.npz"))
    cfg.setAllowResume(True)
    res2 = cfg.fit(ds)
    assert len(res2) == 4 and all(r.get("resumed") for r in res2)
    assert os.path.getmtime(os.path.join(exp, "weights", "best-1.1.weights.npz")) == stamp
    # learning-rate range test
    finder = cfg.lr_find(ds, start_lr=1e-5, end_lr=1e-1, epochs=2)
    assert len(finder.lrs) >= 3 and len(finder.lrs) == len(finder.losses)
    assert abs(finder.lrs[0] - 1e-5) < 1e-12 and all(b > a for a, b in zip(finder.lrs, finder.lrs[1:]))
    assert all(np.isfinite(v) for v in finder.losses[:-1]) and 1e-5 <= finder.best_lr() <= 1e-1
    del segmentation.extra_train["more"]


def test_fit_and_predict_on_crops(cuda, tmp_path):
    """`crops: 2` (README.md:471-491): 128x128 images are trained on as four 64x64 cells and predicted cell by cell, the
    assembled mask coming back at the image's own size."""
    import cv2
    import yaml
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet
    _make_dataset(str(tmp_path), n=4, size=128)
    spec = {"architecture": "Unet", "backbone": "resnet18", "classes": 1, "activation": "sigmoid", "shape": [64, 64, 3],
            "batch": 4, "folds_count": 2, "crops": 2, "loss": "binary_crossentropy+dice_loss", "metrics": ["binary_accuracy"],
            "primary_metric": "val_loss", "stages": [{"epochs": 2}]}
    cfgp = str(tmp_path / "exp" / "config.yaml")
    os.makedirs(os.path.dirname(cfgp))
    yaml.safe_dump(spec, open(cfgp, "w"))
    cfg = segmentation.parse(cfgp)
    ds = SimplePNGMaskDataSet(str(tmp_path / "img"), str(tmp_path / "mask"))
    res = cfg.fit(ds)
    assert len(res) == 2
    rows = list(csv.DictReader(open(os.path.join(os.path.dirname(cfgp), "metrics", "metrics-0.0.csv"))))
    assert len(rows) == 2 and all(np.isfinite(float(r["val_loss"])) for r in rows)
    out = str(tmp_path / "pred")
    assert cfg.predict_to_directory(str(tmp_path / "img"), out, fold=0, stage=0) == 4
    m = cv2.imread(os.path.join(out, "00.png"), cv2.IMREAD_GRAYSCALE)
    assert m.shape == (128, 128) and m.dtype == np.uint8
    b = next(iter(cfg.evaluateAll(ds, fold=0, stage=0)))
    assert b.results[0].shape == (128, 128, 1)


def test_fit_reference_example_config_verbatim(cuda, tmp_path):
    """The reference's own example experiment examples/people/ds_1.yaml, VERBATIM (DeepLabV3 / mobilenetv2, shape 320, batch 10,
    Fliplr+Flipud+Rotate90, binary_crossentropy, Adam, EarlyStopping + ReduceLROnPlateau on val_iou_coef, `datasets:` +
    `fit_with:` pointing at ../../picsart/train): parsed, built, trained and validated through cfg.fit() with no dataset
    argument, then used for prediction.  Only deviations: the stage's 100 epochs are cut to 2 after parsing (test time) and
    `encoder_weights: pascal_voc` is served by a local .npz twin of the file the reference downloads (no network)."""
    import cv2
    from segmentation_pipeline import segmentation
    root = tmp_path
    exp = root / "examples" / "people"
    os.makedirs(exp)
    cfgp = str(exp / "ds_1.yaml")
    shutil.copy(os.path.join(HERE, "golden", "configs", "reference_examples", "ds_1.yaml"), cfgp)
    os.makedirs(root / "picsart" / "train")
    os.makedirs(root / "picsart" / "train_mask")
    rng = np.random.default_rng(0)
    for k in range(25):
        h, w = int(rng.integers(90, 140)), int(rng.integers(90, 140))          # ragged sizes: "everything will be resized to fit"
        yy, xx = np.mgrid[0:h, 0:w]
        cy, cx, r = int(rng.integers(h // 4, 3 * h // 4)), int(rng.integers(w // 4, 3 * w // 4)), min(h, w) // 4
        m = (((yy - cy) ** 2 + (xx - cx) ** 2) < r * r).astype(np.uint8)
        img = (rng.integers(0, 120, (h, w, 3)) + m[:, :, None] * 120).astype(np.uint8)
        cv2.imwrite(str(root / "picsart" / "train" / ("%02d.jpg" % k)), img)
        cv2.imwrite(str(root / "picsart" / "train_mask" / ("%02d.png" % k)), m * 255)
    # the .npz twin of deeplabv3_mobilenetv2_tf_dim_ordering_tf_kernels.h5 (by-name load; the class layer is not in it)
    probe = segmentation.parse(cfgp)
    probe.encoder_weights = None
    w0 = probe.createNet(batch=2).get_weights()
    marker = {k: v for k, v in w0.items() if not k.startswith("custom_logits")}
    marker["Conv_BN/moving_mean"] = np.full_like(marker["Conv_BN/moving_mean"], 0.125)
    np.savez(str(exp / "deeplabv3_mobilenetv2_tf_dim_ordering_tf_kernels.npz"), **marker)
    del probe

    cfg = segmentation.parse(cfgp)
    assert cfg.architecture == "DeepLabV3" and cfg.backbone == "mobilenetv2" and cfg.batch == 10 and cfg.encoder_weights == "pascal_voc"
    assert float(cfg.createNet(batch=2).get_weights()["Conv_BN/moving_mean"][0]) == 0.125       # pascal_voc twin was loaded by name
    cfg.stages[0]["epochs"] = 2
    res = cfg.fit(foldsToExecute=[0])
    assert len(res) == 1
    rows = list(csv.DictReader(open(str(exp / "metrics" / "metrics-0.0.csv"))))
    assert len(rows) == 2
    for r in rows:
        for k in ("loss", "val_loss", "binary_accuracy", "val_binary_accuracy", "iou", "val_iou"):
            assert np.isfinite(float(r[k])), (k, r)
    assert float(rows[-1]["loss"]) < float(rows[0]["loss"])
    assert os.path.exists(str(exp / "weights" / "best-0.0.weights.npz"))
    out = str(root / "pred")
    assert cfg.predict_to_directory(str(root / "picsart" / "train"), out, fold=0, stage=0, ttflips=True) == 25
    p = cv2.imread(os.path.join(out, "00.png"), cv2.IMREAD_GRAYSCALE)
    src = cv2.imread(str(root / "picsart" / "train" / "00.jpg"))
    assert p.shape == src.shape[:2]
