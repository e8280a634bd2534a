"""CPU suite: the reference-facing host API (config parsing, loss / augmentation mapping, dataset protocol, folds,
rank sharding) -- no GPU, no libstp compute calls."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = os.path.join(HERE, "golden", "configs")


def test_import_paths_of_the_reference_work():
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import PredictionItem, SimplePNGMaskDataSet
    assert callable(segmentation.parse) and isinstance(segmentation.custom_models, dict)
    assert PredictionItem("a", 1, 2).id == "a" and SimplePNGMaskDataSet is not None


def test_parse_baseline_config():
    from segmentation_pipeline import segmentation
    cfg = segmentation.parse(os.path.join(CFG, "c2_unet_resnet34.yaml"))
    assert cfg.path.endswith("c2_unet_resnet34.yaml")
    assert (cfg.architecture, cfg.backbone, cfg.classes, cfg.batch) == ("Unet", "resnet34", 1, 16)
    assert cfg.shape == [512, 512, 3] and cfg.folds_count == 5 and cfg.random_state == 33
    assert segmentation.parse_loss(cfg.loss) == (1.0, 1.0, 0.0)
    a = segmentation.parse_augmentation(cfg.augmentation)
    assert a.fliplr == 0.5 and a.flipud == 0.5 and a.affine and a.scale == (0.8, 1.5)
    assert a.translate_x == (-0.2, 0.2) and a.rotate == (-16.0, 16.0) and a.multiply == (0.8, 1.2) and a.add == (-10, 10)
    c = a.to_c()
    assert c.affine == 1 and c.has_mul == 1 and c.has_add == 1 and abs(c.shear_hi - 16.0) < 1e-12


@pytest.mark.parametrize("expr,want", [("binary_crossentropy", (1, 0, 0)), ("dice_loss", (0, 1, 0)),
                                       ("binary_crossentropy+0.1*dice_loss", (1, 0.1, 0)),
                                       ("0.5*binary_crossentropy + iou_loss*2", (0.5, 0, 2)),
                                       ("binary_crossentropy+dice_loss+iou_loss", (1, 1, 1))])
def test_composite_loss_strings(expr, want):
    from segmentation_pipeline.segmentation import parse_loss
    assert parse_loss(expr) == pytest.approx(want)


def test_lovasz_loss_string():
    from segmentation_pipeline.segmentation import parse_loss
    assert parse_loss("lovasz_loss") == (0.0, 0.0, 0.0, 1.0)
    assert parse_loss("0.5*lovasz_loss") == (0.0, 0.0, 0.0, 0.5)
    with pytest.raises(ValueError):
        parse_loss("binary_crossentropy+lovasz_loss")
    assert parse_loss("binary_crossentropy+0.5*jaccard_loss") == (1.0, 0.0, 0.0, 0.0, 0.5)
    assert parse_loss("focal_loss") == (0.0, 0.0, 0.0, 0.0, 0.0, 1.0)


def test_loss_and_augmenter_errors_are_loud():
    from segmentation_pipeline.segmentation import parse_augmentation, parse_loss
    assert parse_loss("categorical_crossentropy") == (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)    # fused softmax kernel (round 2)
    with pytest.raises(ValueError, match="cannot be combined"):
        parse_loss("categorical_crossentropy+dice_loss")                                     # dice fuses a sigmoid
    with pytest.raises(ValueError):
        parse_loss("no_such_loss")
    with pytest.raises(NotImplementedError):
        parse_augmentation({"ElasticTransformation": {"alpha": 1.0}})


def test_unknown_architecture_and_backbone(tmp_path, capsys):
    from segmentation_pipeline import segmentation
    p = tmp_path / "c.yaml"
    p.write_text("architecture: Nope\nbackbone: resnet34\nclasses: 1\nshape: [64,64,3]\n")
    with pytest.raises(ValueError, match="Unknown architecture"):
        segmentation.parse(str(p)).createNet()
    p.write_text("architecture: Unet\nbackbone: nope\nclasses: 1\nshape: [64,64,3]\n")
    with pytest.raises(ValueError, match="Unknown backbone"):
        segmentation.parse(str(p)).createNet()
    assert "Known backbones" in capsys.readouterr().out


def test_custom_models_plugin_hook(tmp_path):
    from segmentation_pipeline import segmentation
    seen = {}

    def factory(**kw):
        seen.update(kw)
        return "model"

    segmentation.custom_models["MyNet"] = factory
    try:
        p = tmp_path / "c.yaml"
        p.write_text("architecture: MyNet\nbackbone: resnet18\nclasses: 1\nshape: [64,64,3]\n")
        assert segmentation.parse(str(p)).createNet() == "model"
        assert seen["backbone"] == "resnet18" and seen["input_shape"] == (64, 64, 3)
    finally:
        del segmentation.custom_models["MyNet"]


def test_simple_png_mask_dataset(tmp_path):
    import cv2
    from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet
    (tmp_path / "i").mkdir()
    (tmp_path / "m").mkdir()
    rng = np.random.default_rng(0)
    for k in range(3):
        img = rng.integers(0, 256, (20, 24, 3), dtype=np.uint8)
        m = np.zeros((20, 24), np.uint8)
        m[5:10, 3:9] = 255 if k else 0
        cv2.imwrite(str(tmp_path / "i" / ("%d.png" % k)), img)
        cv2.imwrite(str(tmp_path / "m" / ("%d.png" % k)), m)
    ds = SimplePNGMaskDataSet(str(tmp_path / "i"), str(tmp_path / "m"))
    assert len(ds) == 3
    it = ds[1]
    assert it.x.shape == (20, 24, 3) and it.x.dtype == np.uint8 and it.y.shape == (20, 24, 1)
    assert set(np.unique(it.y)) == {0, 1} and it.y.sum() == 30 and it.id == "1"
    assert ds.isPositive(1) and not ds.isPositive(0)


def test_kfold_is_sklearn_shuffled_and_seeded(tmp_path):
    from sklearn.model_selection import KFold
    from segmentation_pipeline import segmentation
    cfg = segmentation.parse(os.path.join(CFG, "c1_plumbing.yaml"))
    f1, f2 = cfg.kfold(11), cfg.kfold(11)
    want = list(KFold(n_splits=2, shuffle=True, random_state=7).split(np.arange(11)))
    for (a, b), (c, d), (e, f) in zip(f1, f2, want):
        assert np.array_equal(a, c) and np.array_equal(b, d) and np.array_equal(a, e) and np.array_equal(b, f)


def test_shard_indices_partitions_global_batches():
    from segmentation_training_pipeline_b200.ddp import shard_indices
    order = np.random.default_rng(1).permutation(103)
    world, batch = 4, 5
    shards = [shard_indices(order, r, world, batch) for r in range(world)]
    assert all(len(s) == (103 // 20) * 5 for s in shards)
    allv = np.concatenate(shards)
    assert len(set(allv.tolist())) == len(allv)                      # disjoint
    # global batch g of the single-process run == concatenation of the ranks' g-th batches
    for g in range(103 // 20):
        cat = np.concatenate([s[g * batch:(g + 1) * batch] for s in shards])
        assert np.array_equal(cat, order[g * 20:(g + 1) * 20])
    assert len(shard_indices(order[:7], 0, 4, 5)) == 0               # ragged tail dropped, empty is fine


class _FakeTrainer:
    def __init__(self, lr):
        self.lr = lr

    def set_lr(self, lr):
        self.lr = lr

    def get_lr(self):
        return self.lr


def test_callbacks_known_answers():
    """keras 2.2.4 EarlyStopping / ReduceLROnPlateau and Kenstler's CyclicLR (reference callbacks.raml:22-48), by hand."""
    from segmentation_training_pipeline_b200 import callbacks as CB
    t = _FakeTrainer(0.01)
    cyc = CB.CyclicLR(base_lr=0.001, max_lr=0.006, step_size=4, mode="triangular")
    cyc.on_train_begin(t)
    lrs = []
    for it in range(17):
        cyc.on_batch_begin(t, it)
        lrs.append(t.lr)
    assert lrs[0] == pytest.approx(0.001) and lrs[4] == pytest.approx(0.006) and lrs[8] == pytest.approx(0.001)
    assert lrs[2] == pytest.approx(0.0035) and lrs[6] == pytest.approx(0.0035) and lrs[12] == pytest.approx(0.006)
    cyc2 = CB.CyclicLR(base_lr=0.001, max_lr=0.006, step_size=4, mode="triangular2")
    cyc2.on_train_begin(t)
    l2 = []
    for it in range(13):
        cyc2.on_batch_begin(t, it)
        l2.append(t.lr)
    assert l2[4] == pytest.approx(0.006) and l2[12] == pytest.approx(0.001 + 0.005 / 2)
    # ReduceLROnPlateau: patience 2, factor 0.5, cooldown 1, monitor val_loss (auto -> min, min_delta 1e-4)
    t = _FakeTrainer(0.01)
    rl = CB.ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=2, cooldown=1)
    hist = []
    for e, v in enumerate([1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.5]):
        rl.on_epoch_end(t, e, {"val_loss": v})
        hist.append(t.lr)
    assert hist == pytest.approx([0.01, 0.01, 0.01, 0.005, 0.005, 0.0025, 0.0025, 0.0025])
    # EarlyStopping: auto mode is max for accuracies; stops after `patience` epochs without improvement
    es = CB.EarlyStopping(monitor="val_binary_accuracy", patience=2)
    es.on_train_begin(t)
    stops = []
    for e, v in enumerate([0.5, 0.6, 0.6, 0.55, 0.7]):
        es.on_epoch_end(t, e, {"val_binary_accuracy": v})
        stops.append(es.stop_training)
    assert stops == [False, False, False, True, True]
    assert [type(c).__name__ for c in CB.build({"EarlyStopping": {"patience": 3}}, {"CyclicLR": None})] == ["EarlyStopping", "CyclicLR"]
    with pytest.raises(NotImplementedError):
        CB.build({"TensorBoard": {}})


def test_rle_known_answers_and_round_trip():
    """Kaggle column-major RLE (reference impl/rle.py:10-35): 1-based starts, top-to-bottom then left-to-right."""
    from segmentation_pipeline.impl.rle import masks_as_image, multi_rle_encode, rle_decode, rle_encode
    m = np.zeros((4, 3), np.uint8)
    m[0:3, 0] = 1          # pixels 1..3 (first column)
    m[1:3, 2] = 1          # third column: pixels 10, 11
    from segmentation_pipeline.impl.rle import rle_decode_hw
    assert rle_encode(m) == "1 3 10 2"
    assert np.array_equal(rle_decode_hw("1 3 10 2", (4, 3)), m)
    rng = np.random.default_rng(0)
    for shape in ((1, 1), (5, 7), (64, 48)):
        a = (rng.random(shape) > 0.6).astype(np.uint8)
        assert np.array_equal(rle_decode_hw(rle_encode(a), shape), a)
    for shape in ((1, 1), (7, 7), (32, 32)):   # square: the reference-exact decode is the true inverse as well
        a = (rng.random(shape) > 0.6).astype(np.uint8)
        assert np.array_equal(rle_decode(rle_encode(a), shape), a)
    assert rle_encode(np.zeros((3, 3), np.uint8)) == "" and rle_decode("", (3, 3)).sum() == 0
    parts = multi_rle_encode(m[:, :, None])
    assert sorted(parts) == ["1 3", "10 2"]
    sq = np.zeros((4, 4), np.uint8)
    sq[0:3, 0] = 1
    sq[1:3, 2] = 1
    assert np.array_equal(masks_as_image(multi_rle_encode(sq[:, :, None]), (4, 4))[:, :, 0], sq)


def test_rle_matches_reference_golden_vectors():
    """tests/golden/rle_golden.npz holds outputs of the REFERENCE's own impl/rle.py (generated by make_rle_golden.py, which
    imports it): encode strings, decode arrays -- including the reference's (w, h)-shaped result for non-square shapes --
    masks_as_image / masks_as_images.  Integer work: bit exact."""
    import os
    from segmentation_pipeline.impl.rle import masks_as_image, masks_as_images, rle_decode, rle_encode
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rle_golden.npz"))
    for k in range(int(g["n"])):
        a, enc, shape = g["mask_%d" % k], str(g["enc_%d" % k]), tuple(int(v) for v in g["shape_%d" % k])
        assert rle_encode(a) == enc, (k, shape)
        if "dec_%d" % k in g.files:
            d = rle_decode(enc, shape)
            want = g["dec_%d" % k]
            assert d.shape == want.shape and d.dtype == want.dtype and np.array_equal(d, want), (k, shape, d.shape, want.shape)
    l = [str(g["mai_in_0"]), str(g["mai_in_1"]), float("nan")]
    got = masks_as_image(l, (9, 9))
    assert got.shape == g["mai_out"].shape and got.dtype == g["mai_out"].dtype and np.array_equal(got, g["mai_out"])
    gots = np.stack(masks_as_images(l, (9, 9)))
    assert gots.dtype == g["mais_out"].dtype and np.array_equal(gots, g["mais_out"])


def test_negatives_selection_and_extra_train_concat():
    """stage keys negatives / validation_negatives (README.md:385-415) and the extra_train_data concatenation (FAQ.md:50-61)."""
    from segmentation_training_pipeline_b200.fit import _Concat, _select_negatives

    class DS:
        def __init__(self, flags):
            self.flags = flags

        def __len__(self):
            return len(self.flags)

        def isPositive(self, i):
            return self.flags[i]

        def __getitem__(self, i):
            return ("item", i)

    ds = DS([True, False, True, False, False, False, True, False])
    idx = np.arange(8)
    rng = np.random.default_rng(0)
    assert list(_select_negatives(ds, idx, "real", rng)) == list(range(8))
    assert list(_select_negatives(ds, idx, None, rng)) == list(range(8))
    assert list(_select_negatives(ds, idx, "none", rng)) == [0, 2, 6]
    one = _select_negatives(ds, idx, 1, rng)
    assert len(one) == 6 and {0, 2, 6} <= set(one) and len(set(one)) == 6
    assert list(_select_negatives(ds, idx, 5, rng)) == list(range(8))   # more negatives asked for than exist
    cat = _Concat(ds, DS([False, True]))
    assert len(cat) == 10 and cat[9] == ("item", 1) and cat[3] == ("item", 3)
    assert cat.isPositive(9) and not cat.isPositive(8) and cat.isPositive(0)


def test_augmentation_block_order_is_checked():
    """imgaug Sequential runs augmenters in YAML order; the fused kernel's order is fixed for the geometric part and free
    for the colour part -- anything else must raise instead of silently computing a different pipeline."""
    from segmentation_training_pipeline_b200.segmentation import parse_augmentation
    c = parse_augmentation({"Rotate90": True, "Fliplr": 0.5, "Affine": {"rotate": [-5, 5]}, "Invert": 0.25, "Add": [-3, 3], "Multiply": [0.9, 1.1]})
    assert c.rot90 and c.invert == 0.25 and c.color_order == (2, 1, 0) and c.affine and c.enabled()
    assert parse_augmentation({"Multiply": [0.9, 1.1]}).color_order == (0, 1, 2)
    with pytest.raises(NotImplementedError, match="order"):
        parse_augmentation({"Multiply": [0.9, 1.1], "Fliplr": 0.5})
    with pytest.raises(NotImplementedError, match="order"):
        parse_augmentation({"Affine": {}, "Flipud": 0.5})
    with pytest.raises(NotImplementedError, match="not fused"):
        parse_augmentation({"ElasticTransformation": {"alpha": 0.5, "sigma": 0.25}})
    assert parse_augmentation({"DirectedEdgeDetect": {"alpha": 0.5, "direction": [0.0, 0.5]}}).colour_seq[0][1][:5] == (6, 0.5, 0.5, 0.0, 0.5)
    # neighbourhood augmenters (csrc/augment_nb.cu) join the colour block in YAML order, interleaved with the pixel-wise ones
    nb = parse_augmentation({"Fliplr": 0.5, "Multiply": [0.9, 1.1], "GaussianBlur": {"sigma": [0.0, 3.0]}, "Add": [-5, 5],
                             "OneOf": {"AverageBlur": {"k": [2, 7]}, "MedianBlur": {"k": [3, 5]}},
                             "Sharpen": {"alpha": [0, 1.0], "lightness": [0.75, 1.5]}, "Emboss": {"alpha": 0.5, "strength": [0, 2.0]},
                             "EdgeDetect": 0.25})
    assert [t for t, _ in nb.colour_seq] == ["pix", "nb", "pix", "nb", "nb", "nb", "nb", "nb"]
    runs = nb.colour_runs()
    assert [(r[0], r[1]) for r in runs] == [("pix", 0), ("nb", 1), ("pix", 2), ("nb", 3), ("nb", 4), ("nb", 5), ("nb", 6), ("nb", 7)]
    assert runs[1][2][:3] == (0, 0.0, 3.0) and runs[3][2][5:] == (0, 2, 0) and runs[4][2][5:] == (0, 2, 1)     # OneOf group 0, 2 members
    assert runs[5][2][:5] == (3, 0.0, 1.0, 0.75, 1.5) and runs[7][2][:3] == (5, 0.25, 0.25)
    with pytest.raises(ValueError, match="sigma"):
        parse_augmentation({"GaussianBlur": {"sigma": [0.0, 12.0]}})
    with pytest.raises(ValueError, match="MedianBlur"):
        parse_augmentation({"MedianBlur": {"k": [3, 11]}})
    with pytest.raises(ValueError, match="alpha"):
        parse_augmentation({"Sharpen": {"alpha": [0.0, 1.5]}})
    # Rotate90 / Fliplr / Flipud in any order among themselves (the reference's examples list the flips first)
    c = parse_augmentation({"Fliplr": 0.5, "Flipud": 0.5, "Rotate90": True})
    assert c.rot90 and c.fliplr == 0.5 and c.flipud == 0.5 and c.flip_before_rot90 == 3
    assert parse_augmentation({"Fliplr": 0.5, "Rotate90": True, "Flipud": 0.5}).flip_before_rot90 == 1
    assert parse_augmentation({"Rotate90": True, "Fliplr": 0.5, "Flipud": 0.5}).flip_before_rot90 == 0
    assert parse_augmentation({"Flipud": 0.5, "Rotate90": True}).to_c().flip_before_rot90 == 2


def test_oracle_flip_rotate_composition_is_the_literal_sequence():
    """the algebra behind stp_aug_spec.flip_before_rot90, checked on the CPU oracle: flips listed before an odd np.rot90
    equal the OTHER flips applied after it."""
    from oracle import augment as OA
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (6, 6, 3), dtype=np.uint8)
    msk = rng.integers(0, 2, (6, 6, 1), dtype=np.uint8)
    for pre in range(4):
        for k in range(4):
            for a in (False, True):
                for b in (False, True):
                    lit = OA.SampleParams(a, b, np.eye(2, 3), False, 1.0, False, 0, k, False, (0, 1, 2), pre)
                    lr_pre, ud_pre = bool(pre & 1), bool(pre & 2)
                    if k & 1:
                        a2 = (False if lr_pre else a) ^ (b if ud_pre else False)
                        b2 = (False if ud_pre else b) ^ (a if lr_pre else False)
                    else:
                        a2, b2 = a, b
                    dev = OA.SampleParams(a2, b2, np.eye(2, 3), False, 1.0, False, 0, k, False, (0, 1, 2), 0)
                    li, lm = OA.apply(img, msk, lit)
                    di, dm = OA.apply(img, msk, dev)
                    assert np.array_equal(li, di) and np.array_equal(lm, dm), (pre, k, a, b)


def test_reference_example_configs_parse_and_build_their_augmenter():
    """all five experiment files the reference ships (examples/people/*.yaml, copied verbatim as fixtures under
    tests/golden/configs/reference_examples/) parse, expose the reference's keys and build the fused device augmenter --
    their block order is Fliplr, Flipud, Rotate90."""
    import glob
    import os
    from segmentation_pipeline import segmentation
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "configs", "reference_examples", "*.yaml")))
    assert len(files) == 5
    for f in files:
        cfg = segmentation.parse(f)
        assert cfg.architecture == "DeepLabV3" and cfg.backbone == "mobilenetv2" and cfg.classes == 1
        assert cfg.shape == [320, 320, 3] and cfg.activation == "sigmoid" and cfg.optimizer == "Adam"
        aug = segmentation.parse_augmentation(cfg.augmentation)
        assert aug.fliplr == 0.5 and aug.enabled()
        if "Rotate90" in cfg.augmentation:
            assert aug.rot90 and aug.flipud == 0.5 and aug.flip_before_rot90 == 3
        from segmentation_training_pipeline_b200 import callbacks as CB
        assert isinstance(CB.build(cfg.callbacks, None), list)


def test_lr_variator_schedule():
    """musket's LRVariator (README.md:454): fromVal -> toVal over relSize epochs' worth of batches, then constant."""
    from segmentation_training_pipeline_b200 import callbacks as CB

    class T:
        steps_per_epoch = 10
        lr = 0.01

        def get_lr(self):
            return self.lr

        def set_lr(self, v):
            self.lr = v

    t = T()
    (cb,) = CB.build({"LRVariator": {"relSize": 0.5, "toVal": 0.001, "style": "linear"}})
    cb.on_train_begin(t)
    seen = []
    for it in range(8):
        cb.on_batch_begin(t, it)
        seen.append(t.lr)
    assert abs(seen[0] - 0.01) < 1e-12 and abs(seen[5] - 0.001) < 1e-12 and abs(seen[7] - 0.001) < 1e-12
    assert all(b <= a + 1e-15 for a, b in zip(seen, seen[1:])) and abs(seen[1] - (0.01 - 0.009 / 5)) < 1e-12
    for style in CB.LRVariator.STYLES:
        v = CB.LRVariator(style=style)
        assert abs(v.shape(0.0)) < 1e-12 and abs(v.shape(1.0) - 1.0) < 1e-12
        assert all(0.0 <= v.shape(k / 10) <= 1.0 for k in range(11))
    with pytest.raises(ValueError):
        CB.LRVariator(style="zigzag")


def test_crops_cells_and_reassembly():
    """`crops: N` (README.md:471-491): N x N cells per image for training, cell-wise prediction assembled back."""
    from segmentation_pipeline.impl.datasets import PredictionItem
    from segmentation_training_pipeline_b200.crops import CellDataSet, cell_bounds, predict_image_by_cells
    b = cell_bounds(10, 7, 3)
    assert len(b) == 9 and b[0][:2] == (0, 3) and b[-1] == (7, 10, 5, 7)
    cover = np.zeros((10, 7), int)
    for y0, y1, x0, x1 in b:
        cover[y0:y1, x0:x1] += 1
    assert (cover == 1).all()                     # the cells tile the image exactly
    rng = np.random.default_rng(0)
    items = [PredictionItem("a%d" % k, rng.integers(0, 255, (12, 12, 3), dtype=np.uint8),
                            (rng.random((12, 12, 1)) > 0.5).astype(np.uint8)) for k in range(3)]
    cd = CellDataSet(items, 2)
    assert len(cd) == 12 and cd[5].id == "a1_1" and cd[5].x.shape == (6, 6, 3)
    assert np.array_equal(cd[5].x, items[1].x[0:6, 6:12]) and np.array_equal(cd[7].y, items[1].y[6:12, 6:12])
    assert list(cd.expand([2, 0])) == [8, 9, 10, 11, 0, 1, 2, 3]
    # an "identity network" (probability = red channel / 255 at the network shape == cell size) is reassembled exactly
    fn = lambda x: x[..., :1].astype(np.float32) / 255.0
    p = predict_image_by_cells(fn, items[0].x, 2, (6, 6), batch=3)
    assert p.shape == (12, 12, 1) and np.allclose(p[..., 0], items[0].x[..., 0] / 255.0)


@pytest.mark.parametrize("arch,backbone,classes", [("Unet", "resnet34", 1), ("FPN", "resnet18", 1), ("FPN", "resnet50", 3),
                                                   ("Linknet", "resnet18", 1), ("Unet", "vgg16", 1)])
def test_engine_graph_matches_oracle_parameters(arch, backbone, classes):
    """Graph construction needs no GPU (device='cpu' allocates the buffers on the host, nothing is launched): the engine's
    parameter names / Keras-layout shapes must be exactly the oracle's, weights must round-trip through get/set_weights, and
    the structural decisions the step relies on must hold."""
    from oracle.models import SegModel
    from segmentation_training_pipeline_b200 import engine as E
    from segmentation_training_pipeline_b200.models import SegNet
    net = SegNet(backbone, classes=classes, input_shape=(64, 64, 3), batch=2, device="cpu", seed=0, architecture=arch)
    om = SegModel(arch, backbone, classes=classes, input_shape=(64, 64, 3))
    W = net.get_weights()
    assert set(om.params) == set(net.params)
    for k, p in om.params.items():
        assert tuple(p.shape) == W[k].shape, (k, tuple(p.shape), W[k].shape)
    W2 = {k: (v + 1.0).astype(np.float32) for k, v in W.items()}
    net.set_weights(W2)
    W3 = net.get_weights()
    assert all(np.array_equal(W3[k], W2[k]) for k in W2)
    # encoder parameters form the leading range of the flat buffers (freeze_encoder = an offset into the optimizer launch)
    enc_end = net.encoder_floats()
    assert 0 < enc_end < net.n_flat
    for name, p in net.params.items():
        assert (name in net.encoder_param_names) == (p.offset < enc_end), name
    # data-parallel bucket split: ops[i:] own exactly the parameters at offsets >= off, and that tail is >= 90 % of the floats
    i, off = net.split_for_overlap(0.9)
    if off:
        assert (net.n_flat - off) >= 0.9 * net.n_flat * 0.99
        for j, op in enumerate(net.ops):
            for a in ("w", "b", "gamma", "beta"):
                p = getattr(op, a, None)
                if isinstance(p, E.Param):
                    assert (j >= i) == (p.offset >= off), (j, a)
    # every BatchNorm gets its forward statistics from the epilogue of the conv that feeds it (when exactly one conv does)
    bns = [o for o in net.ops if isinstance(o, E.BNRelu)]
    assert bns == [] or sum(o.stats_from_conv for o in bns) >= len(bns) - 2


def test_create_net_key_validation(tmp_path):
    """createNet (reference segmentation.py:96-155): unknown names raise the reference's ValueErrors; keys whose kernels are
    not built raise NotImplementedError naming the key (nothing is silently ignored)."""
    import yaml
    from segmentation_pipeline import segmentation

    def cfg(**kw):
        spec = {"architecture": "Unet", "backbone": "resnet18", "classes": 1, "activation": "sigmoid", "shape": [64, 64, 3], "batch": 2}
        spec.update(kw)
        p = tmp_path / ("c%d.yaml" % len(list(tmp_path.iterdir())))
        yaml.safe_dump(spec, open(p, "w"))
        c = segmentation.parse(str(p))
        c.device = "cpu"
        return c

    with pytest.raises(ValueError, match="Unknown architecture"):
        cfg(architecture="DeepLabV4").createNet()
    # DeepLabV3: the reference's in-tree model over mobilenetv2 (impl/deeplab/model.py); segmentation backbones do not apply
    with pytest.raises(ValueError, match="Unknown backbone"):
        cfg(architecture="DeepLabV3").createNet()
    with pytest.raises(NotImplementedError, match="OS"):
        cfg(architecture="DeepLabV3", backbone="xception", OS=32).createNet()
    xc = cfg(architecture="DeepLabV3", backbone="xception", OS=8).createNet()     # the other backbone of impl/deeplab/model.py
    wx = xc.get_weights()
    assert wx["middle_flow_unit_16_separable_conv3_pointwise/kernel"].shape == (1, 1, 728, 728)
    assert wx["exit_flow_block2_separable_conv3_depthwise/depthwise_kernel"].shape == (3, 3, 1536, 1)
    assert wx["feature_projection0/kernel"].shape == (1, 1, 256, 48) and wx["decoder_conv1_pointwise/kernel"].shape == (1, 1, 256, 256)
    assert sum(v.size for k, v in wx.items() if "moving_" not in k) == 41050273
    with pytest.raises(NotImplementedError, match="alpha"):
        cfg(architecture="DeepLabV3", backbone="mobilenetv2", alpha=0.5).createNet()
    with pytest.raises(NotImplementedError, match="cannot be downloaded"):
        cfg(architecture="DeepLabV3", backbone="mobilenetv2", encoder_weights="pascal_voc").createNet()
    dl = cfg(architecture="DeepLabV3", backbone="mobilenetv2").createNet()
    wd = dl.get_weights()
    assert wd["expanded_conv_16_project/kernel"].shape == (1, 1, 960, 320) and wd["custom_logits_semantic/kernel"].shape == (1, 1, 256, 1)
    assert wd["expanded_conv_3_depthwise/depthwise_kernel"].shape == (3, 3, 144, 1) and dl.activation == "sigmoid"
    assert sum(v.size for k, v in wd.items() if "moving_" not in k) == 2108417      # == the Keras model's trainable parameter count
    # `encoder_weights: pascal_voc`: the .npz twin of the reference's download, found next to the config, is loaded by layer name
    pv = {k: np.full_like(v, 0.25) for k, v in wd.items() if not k.startswith("custom_logits")}
    np.savez(str(tmp_path / "deeplabv3_mobilenetv2_tf_dim_ordering_tf_kernels.npz"), **pv)
    dl2 = cfg(architecture="DeepLabV3", backbone="mobilenetv2", encoder_weights="pascal_voc").createNet()
    wd2 = dl2.get_weights()
    assert float(wd2["aspp0/kernel"].min()) == 0.25 and float(wd2["Conv_BN/moving_mean"].max()) == 0.25
    assert not np.allclose(wd2["custom_logits_semantic/kernel"], 0.25)
    with pytest.raises(ValueError, match="divisible by 6"):
        cfg(architecture="PSPNet").createNet()                       # 64 is not a multiple of 6 * downsample_factor
    with pytest.raises(NotImplementedError, match="psp_pooling_type"):
        cfg(architecture="PSPNet", shape=[96, 96, 3], psp_pooling_type="max").createNet()
    psp = cfg(architecture="PSPNet", shape=[96, 96, 3], psp_conv_filters=64).createNet()
    wp = psp.get_weights()
    assert wp["psp_level6_conv/kernel"].shape == (1, 1, 128, 64) and wp["psp_conv/kernel"].shape == (1, 1, 128 + 4 * 64, 512)
    assert "stage3_unit1_bn1/gamma" in wp and "stage3_unit1_conv1/kernel" not in wp and "stage4_unit1_bn1/gamma" not in wp
    assert wp["final_conv/kernel"].shape == (3, 3, 512, 1)
    with pytest.raises(ValueError, match="Unknown backbone"):
        cfg(backbone="efficientnetb4").createNet()
    with pytest.raises(NotImplementedError, match="softmax"):
        cfg(activation="softmax", classes=3).createNet()
    with pytest.raises(NotImplementedError, match="classes"):
        cfg(classes=7).createNet()
    net = cfg(architecture="FPN", classes=3, activation="softmax", loss="lovasz_loss", pyramid_block_filters=64,
              segmentation_block_filters=32).createNet()
    assert net.activation == "softmax" and net.get_weights()["pyramid_stage_1_conv1x1/kernel"].shape == (1, 1, 256, 64)
    assert net.get_weights()["head_conv/kernel"].shape == (3, 3, 128, 3)
    assert cfg(crops=3).crops == 3
    # encoder_weights: a local .npz of Keras-named encoder arrays (no download here); only encoder layers are taken from it
    base = cfg().createNet()
    w = base.get_weights()
    np.savez(str(tmp_path / "enc.npz"), **{"conv0/kernel": w["conv0/kernel"] + 1.0, "bn0/moving_mean": w["bn0/moving_mean"] + 2.0,
                                            "final_conv/kernel": w["final_conv/kernel"] + 3.0})
    w2 = cfg(encoder_weights="enc.npz").createNet().get_weights()
    assert np.array_equal(w2["conv0/kernel"], w["conv0/kernel"] + 1.0) and np.array_equal(w2["bn0/moving_mean"], w["bn0/moving_mean"] + 2.0)
    assert np.array_equal(w2["final_conv/kernel"], w["final_conv/kernel"])          # decoder arrays of the file are ignored
    with pytest.raises(NotImplementedError, match="encoder_weights"):
        cfg(encoder_weights="imagenet").createNet()
    with pytest.raises(NotImplementedError, match="transforms"):
        cfg(transforms={"Fliplr": 1.0}).fit([])
    with pytest.raises(NotImplementedError, match="dataset_augmenter"):
        cfg(dataset_augmenter={"name": "x"}).fit([])
    segmentation.custom_objects["my_loss"] = lambda t, p: 0.0
    try:
        with pytest.raises(NotImplementedError, match="custom_objects"):
            segmentation.parse_loss("binary_crossentropy+0.5*my_loss")
    finally:
        del segmentation.custom_objects["my_loss"]


def test_prediction_maps_and_ansemble_predictions(tmp_path):
    """What prediction callbacks receive (reference segmentation.py:81-91, README.md:505-513: `img.arr > threshold`) and
    segmentation.ansemblePredictions (README.md:745-754): averaging the .npy predictions of several runs."""
    import cv2
    from segmentation_pipeline import segmentation
    from segmentation_training_pipeline_b200.predict import SegmentationMapOnImage, _scaled_map
    p = np.linspace(0, 1, 64 * 64, dtype=np.float32).reshape(64, 64, 1)
    m = SegmentationMapOnImage(p)
    assert m.shape == (64, 64, 1) and m.arr.dtype == np.float32
    assert np.array_equal(np.asarray(m), (p > 0.5).astype(np.uint8)) and m.get_arr_int(0.25).sum() > m.get_arr_int(0.75).sum()
    big = _scaled_map(p, np.zeros((80, 96, 3), np.uint8))
    assert big.shape == (80, 96, 1) and 0.0 <= big.arr.min() and big.arr.max() <= 1.0
    assert _scaled_map(p, np.zeros((64, 64, 3), np.uint8)).arr is not None
    # two runs' predictions of two files, weights 3:1
    src, a, b = tmp_path / "src", tmp_path / "a", tmp_path / "b"
    for d in (src, a, b):
        d.mkdir()
    for name, va, vb in (("x", 0.8, 0.0), ("y", 0.2, 0.6)):
        cv2.imwrite(str(src / (name + ".png")), np.zeros((8, 8, 3), np.uint8))
        np.save(str(a / name), np.full((8, 8, 1), va, np.float32))
        np.save(str(b / name), np.full((8, 8, 1), vb, np.float32))
    got = {}
    segmentation.ansemblePredictions(str(src), [str(a), str(b)], lambda f, mm, data: data.__setitem__(f, float(mm.arr.mean())), got)
    assert got == pytest.approx({"x.png": 0.4, "y.png": 0.4})
    got2 = {}
    segmentation.ansemblePredictions(str(src), [str(a), str(b)], lambda f, mm, data: data.__setitem__(f, float(mm.arr.mean())), got2,
                                     weights=[3, 1])
    assert got2 == pytest.approx({"x.png": 0.6, "y.png": 0.3})


@pytest.mark.parametrize("crops", [0, 2])
def test_prediction_verbs_host_logic(tmp_path, monkeypatch, crops):
    """predict_to_directory / predict_in_directory / evaluateAll host logic (file walking, scale-back, writers, batch fields,
    fold ensembling, crops reassembly) with the device forward replaced by a stand-in: probability = red channel / 255."""
    import cv2
    from segmentation_pipeline.impl.datasets import PredictionItem
    from segmentation_training_pipeline_b200 import predict as P

    class Net:
        batch, classes, input_shape = 4, 1, (64, 64, 3)

    class Cfg:
        shape, folds_count = [64, 64, 3], 2

        def load_model(self, fold, stage):
            return Net()

    cfg = Cfg()
    cfg.crops = crops
    monkeypatch.setattr(P, "predict_arrays", lambda net, x, ttflips=False: x[..., :1].astype(np.float32) / 255.0)
    src = tmp_path / "src"
    src.mkdir()
    rng = np.random.default_rng(0)
    imgs = {"a.png": rng.integers(0, 256, (80, 96, 3), dtype=np.uint8), "b.png": rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)}
    for k, v in imgs.items():
        cv2.imwrite(str(src / k), cv2.cvtColor(v, cv2.COLOR_RGB2BGR))
    out = tmp_path / "out"
    assert P.predict_to_directory(cfg, str(src), str(out), fold=[0, 1], stage=0) == 2
    pb = cv2.imread(str(out / "b.png"), cv2.IMREAD_GRAYSCALE)
    assert pb.shape == (64, 64)
    if not crops:   # arr*255 of red/255; with crops every cell goes through a cubic up- and a bilinear down-sampling
        assert np.abs(pb.astype(int) - imgs["b.png"][..., 0].astype(int)).max() <= 1
    else:
        assert np.corrcoef(pb.reshape(-1).astype(float), imgs["b.png"][..., 0].reshape(-1).astype(float))[0, 1] > 0.5
    assert cv2.imread(str(out / "a.png"), cv2.IMREAD_GRAYSCALE).shape == (80, 96)
    arrs = tmp_path / "arr"
    assert P.predict_to_directory(cfg, str(src), str(arrs), binaryArray=True) == 2
    assert np.load(str(arrs / "a.npy")).shape == (80, 96, 1) and np.load(str(arrs / "b.npy")).dtype == np.float32
    seen = {}
    P.predict_in_directory(cfg, str(src), 0, 0, lambda f, m, data: data.__setitem__(f, (m.shape, float((m.arr > 0.25).mean()))), seen)
    assert seen["a.png"][0] == (80, 96, 1) and seen["b.png"][0] == (64, 64, 1) and 0.6 < seen["b.png"][1] < 0.9
    ds = [PredictionItem("i%d" % k, v, (v[..., :1] > 127).astype(np.uint8)) for k, v in enumerate(imgs.values())]
    batches = list(P.evaluate_all(cfg, ds, fold=0, stage=0))
    assert sum(len(b.data) for b in batches) == 2
    b0 = batches[0]
    assert b0.results[0].shape == (80, 96, 1) and b0.predicted_maps_aug[0].shape == (80, 96, 1)
    assert b0.segmentation_maps[0].arr.shape == (80, 96, 1) and b0.images[0].shape == (80, 96, 3)
    if not crops:   # PipelineConfig.evaluate: heat maps of the validation items of a fold at network resolution
        cfg.kfold = lambda n: [(np.array([0]), np.array([1])), (np.array([1]), np.array([0]))]
        ev = list(P.evaluate(cfg, ds, 0, 0, limit=16))
        assert len(ev) == 1 and len(ev[0].data) == 1 and ev[0].heatmaps_aug[0].shape == (64, 64, 1) and ev[0].images_aug[0].shape == (64, 64, 3)


@pytest.mark.parametrize("workers", [0, 3])
def test_host_loader_prefetch_ring(workers):
    """loader.HostLoader: batches decoded / resized by a thread pool two batches ahead into a ring of four host buffers; every
    yielded pair must hold exactly its batch (in order, ragged sizes resized) while the consumer still looks at the previous
    one, and a worker's exception must surface in the consumer."""
    from segmentation_pipeline.impl.datasets import PredictionItem
    from segmentation_training_pipeline_b200.loader import HostLoader

    class DS:
        def __len__(self):
            return 23

        def __getitem__(self, i):
            if i == 22:
                raise IOError("broken file")
            h, w = (16, 16) if i % 3 else (20, 12)                     # every third item has another size
            x = np.full((h, w, 3), i, np.uint8)
            y = np.full((h, w, 1), i % 2, np.uint8)
            return PredictionItem(str(i), x, y)

    ld = HostLoader(DS(), (16, 16, 3), 1, batch=4, workers=workers, pin=False)
    batches = [[(4 * s + j) % 22 for j in range(4)] for s in range(9)]
    prev = None
    for k, (img, mask) in enumerate(ld.iterate(batches)):
        assert img.shape == (4, 16, 16, 3) and mask.shape == (4, 16, 16, 1)
        assert [int(img[j, 0, 0, 0]) for j in range(4)] == batches[k]
        assert [int(mask[j, 5, 5, 0]) for j in range(4)] == [i % 2 for i in batches[k]]
        if prev is not None:   # the previous pair is still intact while this one is being consumed
            assert [int(prev[0][j, 0, 0, 0]) for j in range(4)] == batches[k - 1]
        prev = (img, mask)
    with pytest.raises(IOError):
        list(ld.iterate([[0, 1, 2, 3], [4, 5, 22, 6]]))
    ld.close()


def test_run_fit_control_flow_with_stub_engine(tmp_path, monkeypatch):
    """The k-fold x stage loop of fit() (reference generic fit, README.md:116-205) end to end on the CPU with the device engine
    replaced by a stub: files written, stage keys (negatives, initial_weights, freeze / unfreeze, callbacks), host loader,
    resume -- everything except the kernels."""
    import csv
    import yaml
    import torch
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import PredictionItem
    from segmentation_training_pipeline_b200 import fit as F, trainer as TR

    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    log = {"trainers": [], "steps": 0, "set_weights": 0}

    class Net:
        batch, classes, training = 2, 1, True
        device = torch.device("cpu")
        buffers = {}

        class loss:
            @staticmethod
            def set_weights(*w):
                log["loss_w"] = w

        def __init__(self):
            self.w = {"conv0/kernel": np.zeros(3, np.float32)}

        def get_weights(self):
            return {k: v.copy() for k, v in self.w.items()}

        def set_weights(self, d, strict=True):
            log["set_weights"] += 1
            self.w.update({k: np.asarray(v) for k, v in d.items() if k in self.w})

        def prep_weights(self):
            pass

        def forward(self):
            pass

    class StubTrainer:
        def __init__(self, net, **kw):
            self.net, self.kw, self.lr, self.k = net, kw, float(kw.get("lr") or 1e-3), 0
            log["trainers"].append(kw)

        def enable_host_feed(self):
            pass

        def set_lr(self, v):
            self.lr = v

        def get_lr(self):
            return self.lr

        def step_from_host_pipelined(self, img, mask):
            assert img.shape == (2, 32, 32, 3) and mask.shape == (2, 32, 32, 1) and img.dtype == torch.uint8
            log["steps"] += 1
            self.net.w["conv0/kernel"] += 1.0
            self.k += 1
            return None if self.k == 1 else {"loss": 1.0 / self.k, "binary_accuracy": 0.5}

        def flush_host_pipeline(self):
            return {"loss": 0.25, "binary_accuracy": 0.75}

        def set_batch(self, img, mask):
            pass

        def metrics(self):
            return {"loss": 0.5 / (1 + log["steps"]), "binary_accuracy": 0.9}

    monkeypatch.setattr(TR, "Trainer", StubTrainer)
    spec = {"architecture": "Unet", "backbone": "resnet18", "classes": 1, "shape": [32, 32, 3], "batch": 2, "folds_count": 2,
            "metrics": ["binary_accuracy"], "primary_metric": "val_loss", "freeze_encoder": True, "random_state": 3,
            "callbacks": {"EarlyStopping": {"patience": 5, "monitor": "val_loss"}},
            "stages": [{"epochs": 2, "negatives": "none"},
                       {"epochs": 1, "unfreeze_encoder": True, "lr": 0.01, "initial_weights": "weights/best-0.0.weights",
                        "extra_callbacks": {"LRVariator": {"relSize": 1.0, "toVal": 0.001}}}]}
    cfgp = tmp_path / "exp" / "config.yaml"
    cfgp.parent.mkdir()
    yaml.safe_dump(spec, open(cfgp, "w"))
    cfg = segmentation.parse(str(cfgp))
    monkeypatch.setattr(cfg, "createNet", lambda *a, **k: Net())
    rng = np.random.default_rng(0)
    ds = [PredictionItem("s%d" % i, rng.integers(0, 255, (40, 36, 3), dtype=np.uint8),
                         np.full((40, 36, 1), 0 if i in (2, 5) else 1, np.uint8)) for i in range(8)]
    res = cfg.fit(ds)
    exp = str(cfgp.parent)
    assert len(res) == 4 and os.path.exists(os.path.join(exp, "summary.yaml"))
    assert [t["freeze_encoder"] for t in log["trainers"]] == [True, False, True, False]
    assert log["trainers"][1]["lr"] == 0.01 and log["set_weights"] == 2          # initial_weights loaded once per fold
    rows = list(csv.DictReader(open(os.path.join(exp, "metrics", "metrics-0.0.csv"))))
    assert len(rows) == 2 and set(rows[0]) >= {"epoch", "loss", "binary_accuracy", "val_loss", "val_binary_accuracy", "lr"}
    # 8 items, 2 folds -> 4 training items per fold, minus the negatives of that fold ("negatives: none") -> 1 or 2 steps per epoch
    assert 2 * 2 * 1 + 2 * 2 <= log["steps"] <= 2 * 2 * 2 + 2 * 2
    rows1 = list(csv.DictReader(open(os.path.join(exp, "metrics", "metrics-1.1.csv"))))
    assert len(rows1) == 1 and 0.001 <= float(rows1[0]["lr"]) <= 0.01
    assert os.path.exists(os.path.join(exp, "weights", "best-1.1.weights.npz"))
    with pytest.raises(ValueError, match="already finished"):
        cfg.fit(ds)
    os.remove(os.path.join(exp, "summary.yaml"))
    steps_before = log["steps"]
    cfg.setAllowResume(True)
    res2 = cfg.fit(ds)
    assert all(r.get("resumed") for r in res2) and log["steps"] == steps_before


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's CPU path = the oracle port, timed on the host cores) prints ONE JSON line with
    the contract's keys; reduced size so the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--size", "64", "--ref-batch", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config"):
        assert k in d, k


def test_schema_typed_model_kwargs_are_honoured_or_rejected_by_name():
    """reference createNet1 (segmentation.py:119-129) forwards every non-custom schema property to the architecture's
    constructor; here each such key (schemas/segmentation.raml:158-248, read by schema.py) is either built or raises naming
    the key -- none is dropped silently (Linknet + upsample_layer: transpose used to pass)."""
    from segmentation_training_pipeline_b200 import schema
    from segmentation_training_pipeline_b200.segmentation import PipelineConfig
    assert schema.architectures() == ["Unet", "FPN", "Linknet", "PSPNet", "DeepLabV3"]
    mk = schema.model_keys("Unet")
    assert mk["use_batchnorm"] == ("decoder_use_batchnorm", True) and mk["shape"][0] == "input_shape" and mk["backbone"][0] == "backbone_name"
    assert "loss" not in mk and "batch" not in mk and "augmentation" not in mk      # (meta.custom) keys never reach the constructor
    assert schema.model_keys("FPN")["last_upsample"] == ("last_upsample", 4)
    base = dict(backbone="resnet18", classes=1, activation="sigmoid", shape=[64, 64, 3])
    ok = PipelineConfig(architecture="Linknet", **base)._model_kwargs("Linknet")
    assert ok["upsample_layer"] == "upsampling" and ok["use_batchnorm"] is True
    for arch, key, val in [("Linknet", "upsample_layer", "transpose"), ("Linknet", "decoder_use_batchnorm", False),
                           ("Linknet", "n_upsample_blocks", 4), ("Linknet", "upsample_kernel_size", [2, 2]),
                           ("Unet", "n_upsample_blocks", 4), ("Unet", "upsample_rates", [2, 2, 2, 2, 4]),
                           ("FPN", "interpolation", "nearest"), ("FPN", "last_upsample", 2), ("FPN", "upsample_rates", [2, 2, 4]),
                           ("FPN", "use_batchnorm", False), ("FPN", "dropout", 0.3)]:
        cfg = PipelineConfig(architecture=arch, **base, **{key: val})
        name = {"decoder_use_batchnorm": "use_batchnorm"}.get(key, key)
        with pytest.raises(NotImplementedError, match=name):
            cfg._model_kwargs(arch)
    # schema defaults / supported values pass; decoder_use_batchnorm: false is BUILT for the Unet upsampling decoder
    assert PipelineConfig(architecture="Unet", **base, decoder_use_batchnorm=False)._model_kwargs("Unet")["use_batchnorm"] is False
    assert PipelineConfig(architecture="Unet", **base, use_batchnorm=False, n_upsample_blocks=5)._model_kwargs("Unet")["use_batchnorm"] is False
    with pytest.raises(NotImplementedError, match="use_batchnorm"):
        PipelineConfig(architecture="Unet", **base, use_batchnorm=False, decoder_block_type="transpose")._model_kwargs("Unet")
    assert PipelineConfig(architecture="FPN", **base, dropout=0, interpolation="bilinear")._model_kwargs("FPN")["last_upsample"] == 4


def test_crops_build_the_network_at_cell_resolution():
    """reference createNet1 (segmentation.py:131-132): with `crops: N` the model's input_shape is (H//N, W//N, C)."""
    from segmentation_training_pipeline_b200.segmentation import PipelineConfig
    cfg = PipelineConfig(architecture="Unet", backbone="resnet18", classes=1, shape=[256, 384, 3], crops=2)
    assert cfg.net_shape() == [128, 192, 3]
    assert PipelineConfig(architecture="Unet", shape=[256, 256, 3]).net_shape() == [256, 256, 3]


def test_declarative_datasets_with_bindings(tmp_path):
    """`datasets:` entries in the bindings form of the reference's examples (ds_2.yaml:43-66, ds_3.yaml:43-76): channels picked
    from several folders / readers, binary masks, colour-keyed masks; the short {input_path, output_path} form yields the same
    items; fit_with resolves either."""
    import cv2
    import shutil
    import yaml
    from segmentation_pipeline import segmentation
    from segmentation_pipeline.impl.datasets import BoundDataSet, SimplePNGMaskDataSet, dataset_from_spec
    rng = np.random.default_rng(0)
    img_dir, mask_dir, rgba_dir = tmp_path / "img", tmp_path / "mask", tmp_path / "rgba"
    for d in (img_dir, mask_dir, rgba_dir):
        d.mkdir()
    colors = [[56, 37, 13], [123, 12, 11], [56, 56, 68], [255, 255, 255], [43, 37, 21]]
    imgs = {}
    for i in range(3):
        rgb = rng.integers(0, 256, (24, 20, 3), dtype=np.uint8)
        m = (rng.random((24, 20)) > 0.6).astype(np.uint8) * 255
        cm = np.zeros((24, 20, 3), np.uint8)
        cm[:, :] = [1, 2, 3]
        cm[2:6, 3:9] = colors[3]
        cm[10:12, 0:4] = colors[1]
        cv2.imwrite(str(img_dir / ("s%d.png" % i)), rgb[:, :, ::-1])
        cv2.imwrite(str(mask_dir / ("s%d.png" % i)), m)
        cv2.imwrite(str(rgba_dir / ("s%d.png" % i)), cm[:, :, ::-1])
        imgs[i] = (rgb, m, cm)
    # ds_2.yaml form: identical to the short form
    spec2 = {"inputs": [{"name": "default", "data_type": "uint8", "bindings": [
                {"reader": "RGBA", "path": "img", "bind": [0, 1, 2], "treat": {'type"': "as_is"}}]}],
             "outputs": [{"name": "default", "data_type": "bool", "bindings": [
                {"reader": "monochrome", "path": "mask", "bind": [0], "treat": {"type": "binary_mask"}}]}]}
    a = dataset_from_spec(spec2, str(tmp_path))
    b = dataset_from_spec({"input_path": "img", "output_path": "mask"}, str(tmp_path))
    assert isinstance(a, BoundDataSet) and isinstance(b, SimplePNGMaskDataSet) and len(a) == len(b) == 3
    for i in range(3):
        assert a[i].id == b[i].id == "s%d" % i
        assert np.array_equal(a[i].x, imgs[i][0]) and np.array_equal(b[i].x, imgs[i][0])
        assert np.array_equal(a[i].y, (imgs[i][1] > 0).astype(np.uint8)[:, :, None]) and np.array_equal(a[i].y, b[i].y)
    assert a.isPositive(0)
    # ds_3.yaml form: channels 0,1 from one binding, channel 2 from another; output = 4th channel of the colour-keyed tensor
    spec3 = {"inputs": [{"name": "default", "data_type": "uint8", "bindings": [
                {"reader": "RGBA", "path": "img", "bind": [0, 1], "treat": {'type"': "as_is"}},
                {"reader": "RGBA", "path": "img", "bind": [2], "treat": {'type"': "as_is"}}]}],
             "outputs": [{"name": "default", "data_type": "bool", "bindings": [
                {"reader": "RGBA", "path": "rgba", "bind": [3], "treat": {"type": "binary_mask", "colors": colors}}]}]}
    c = dataset_from_spec(spec3, str(tmp_path))
    assert np.array_equal(c[1].x, imgs[1][0])
    want = np.zeros((24, 20, 1), np.uint8)
    want[2:6, 3:9] = 1
    assert np.array_equal(c[1].y, want)
    with pytest.raises(IndexError):
        dataset_from_spec({**spec3, "outputs": [{"bindings": [{"reader": "monochrome", "path": "mask", "bind": [2]}]}]}, str(tmp_path))[0]
    with pytest.raises(NotImplementedError, match="reader"):
        dataset_from_spec({**spec3, "outputs": [{"bindings": [{"reader": "hsv", "path": "mask", "bind": [0]}]}]}, str(tmp_path))[0]
    # fit_with picks the entry by name; the list-valued `composite:` key of the examples is not a dataset
    cfgd = {"architecture": "Unet", "backbone": "resnet18", "classes": 1, "shape": [32, 32, 3], "fit_with": "simple_1",
            "datasets": {"simple": {"input_path": "img", "output_path": "mask"}, "composite": ["default"], "simple_1": spec2}}
    (tmp_path / "c.yaml").write_text(yaml.safe_dump(cfgd))
    cfg = segmentation.parse(str(tmp_path / "c.yaml"))
    assert isinstance(cfg._resolve_dataset(None), BoundDataSet)
    cfg.fit_with = "simple"
    assert isinstance(cfg._resolve_dataset(None), SimplePNGMaskDataSet)
    cfg.fit_with = "composite"
    with pytest.raises(ValueError, match="not a dataset"):
        cfg._resolve_dataset(None)


def test_crop_pad_augmenters_parse():
    """Pad / PadToFixedSize / CropToFixedSize / CropAndPad (schemas/augmenters.raml:72-87, 113-116) lead the block and are
    composed into one window; keep_size ops must come last of the family."""
    from segmentation_training_pipeline_b200.segmentation import parse_augmentation
    c = parse_augmentation({"PadToFixedSize": {"width": 600, "height": 600}, "CropToFixedSize": {"width": 512, "height": 512},
                            "Fliplr": 0.5, "Multiply": [0.9, 1.1]})
    assert c.crop_pad == ((2, 0, 600, 600, 0, 0), (3, 0, 512, 512, 0, 0)) and c.fliplr == 0.5 and c.enabled()
    assert parse_augmentation({"Pad": {"px": [1, 2, 3, 4]}}).crop_pad == ((1, 0, 1, 2, 3, 4),)
    assert parse_augmentation({"Pad": 5}).crop_pad == ((1, 0, 5, 5, 5, 5),)
    assert parse_augmentation({"CropAndPad": {"percent": [-0.1, 0.2]}}).crop_pad == ((4, 1, -0.1, 0.2, 0.0, 0.0),)
    assert parse_augmentation({"CropAndPad": {"percent": [0.1, 0.0, -0.1, 0.05]}}).crop_pad[0][:2] == (4, 0)
    spec = parse_augmentation({"CropToFixedSize": {"width": 8, "height": 8}, "CropAndPad": {"percent": 0.1}}).croppad_c()
    assert spec.n_ops == 2 and spec.ops[0].kind == 3 and spec.ops[1].kind == 4 and abs(spec.ops[1].a - 0.1) < 1e-7
    with pytest.raises(NotImplementedError, match="order"):
        parse_augmentation({"Fliplr": 0.5, "CropToFixedSize": {"width": 8, "height": 8}})
    with pytest.raises(NotImplementedError, match="keep_size"):
        parse_augmentation({"Pad": 3, "CropToFixedSize": {"width": 8, "height": 8}})
    with pytest.raises(ValueError):
        parse_augmentation({"CropAndPad": {"percent": [0.1, 0.2, 0.3]}})


def test_oracle_crop_pad_window_composition():
    from oracle import augment as OA
    # PadToFixedSize then CropToFixedSize back to the original size: a pure shift, never larger than the pad
    for sample in range(6):
        vy0, vx0, vh, vw = OA.crop_pad_window([(2, 0, 80, 80, 0, 0), (3, 0, 64, 64, 0, 0)], 1, 0, sample, 64, 64)
        assert (vh, vw) == (64, 64) and -16 <= vy0 <= 16 and -16 <= vx0 <= 16
    assert OA.crop_pad_window([(1, 0, 1, 2, 3, 4)], 0, 0, 0, 10, 20) == (-1, -4, 14, 26)
    assert OA.crop_pad_window([(4, 0, 0.1, 0.1, 0.1, 0.1)], 0, 0, 0, 100, 50) == (-10, -5, 120, 60)
    assert OA.crop_pad_window([(4, 0, -0.25, 0.0, 0.0, 0.0)], 0, 0, 0, 100, 50) == (25, 0, 75, 50)


def test_raw_batch_packing_for_device_ingest():
    """loader.RawBatch (on-device ingest, `device_resize: true`): samples of different sizes packed back to back, one
    stp_resize_item each; HostLoader hands them out through the same iterate() contract (no GPU needed: unpinned)."""
    import ctypes as C
    from segmentation_training_pipeline_b200 import lib
    from segmentation_training_pipeline_b200.impl.datasets import PredictionItem
    from segmentation_training_pipeline_b200.loader import HostLoader, RawBatch
    rng = np.random.default_rng(0)
    samples = [(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), rng.integers(0, 2, (h, w, 1), dtype=np.uint8)) for h, w in [(5, 7), (8, 8), (3, 11)]]
    rb = RawBatch(64, 16, 4, pin=False)       # deliberately too small: grows
    rb.pack(samples, 3, 1)
    assert rb.n == 3 and rb.used_img >= sum(x.size for x, _ in samples)
    host = bytes(rb.items_img.numpy())
    hostm = bytes(rb.items_mask.numpy())
    for j, (x, y) in enumerate(samples):
        it = lib.ResizeItem.from_buffer_copy(host, j * C.sizeof(lib.ResizeItem))
        assert (it.sh, it.sw, it.vy0, it.vx0, it.vh, it.vw) == (x.shape[0], x.shape[1], 0, 0, x.shape[0], x.shape[1]) and it.src_off % 16 == 0
        assert np.array_equal(rb.arena_img.numpy()[it.src_off:it.src_off + x.size].reshape(x.shape), x)
        im = lib.ResizeItem.from_buffer_copy(hostm, j * C.sizeof(lib.ResizeItem))
        assert np.array_equal(rb.arena_mask.numpy()[im.src_off:im.src_off + y.size].reshape(y.shape), y)

    class DS:
        def __len__(self):
            return 3

        def __getitem__(self, i):
            return PredictionItem(str(i), samples[i][0], samples[i][1][:, :, 0])

    ld = HostLoader(DS(), (32, 32, 3), 1, 2, workers=0, pin=False, device_resize=True)
    out = list(ld.iterate([[0, 1], [2, 0]]))
    assert len(out) == 2 and all(m is None and b.n == 2 for b, m in out)


def test_nchannel_encoder_weight_adaptation(tmp_path):
    """reference createNet1 (segmentation.py:138-153): a >3-channel input with pretrained 3-channel encoder weights -- first
    conv widened (RGB planes copied, extra planes = their mean), result cached next to the config as .mdl-nchannel."""
    from segmentation_training_pipeline_b200.segmentation import PipelineConfig
    cfg = PipelineConfig(architecture="Unet", backbone="resnet18", classes=1, shape=[64, 64, 4])
    cfg.path = str(tmp_path / "c.yaml")

    class FakeNet:
        def get_weights(self):
            return {"conv0/kernel": np.zeros((7, 7, 4, 64), np.float32), "bn_data/beta": np.zeros(4, np.float32),
                    "bn_data/moving_variance": np.ones(4, np.float32), "bn0/gamma": np.ones(64, np.float32)}

    rng = np.random.default_rng(0)
    w3 = {"conv0/kernel": rng.normal(size=(7, 7, 3, 64)).astype(np.float32), "bn_data/beta": np.array([1, 2, 3], np.float32),
          "bn_data/moving_variance": np.array([4, 5, 6], np.float32), "bn0/gamma": rng.normal(size=64).astype(np.float32)}
    out = cfg._adapt_input_channels(FakeNet(), w3)
    assert out["conv0/kernel"].shape == (7, 7, 4, 64)
    assert np.array_equal(out["conv0/kernel"][:, :, :3], w3["conv0/kernel"])
    assert np.allclose(out["conv0/kernel"][:, :, 3], w3["conv0/kernel"].mean(axis=2))
    assert np.allclose(out["bn_data/beta"], [1, 2, 3, 2]) and np.allclose(out["bn_data/moving_variance"], [4, 5, 6, 5])
    assert np.array_equal(out["bn0/gamma"], w3["bn0/gamma"])
    assert (tmp_path / "c.yaml.mdl-nchannel.npz").exists()
    again = cfg._adapt_input_channels(FakeNet(), {})           # served from the cache
    assert np.array_equal(again["conv0/kernel"], out["conv0/kernel"])
    cfg3 = PipelineConfig(architecture="Unet", backbone="resnet18", classes=1, shape=[64, 64, 3])
    assert cfg3._adapt_input_channels(FakeNet(), w3) is w3


def test_pixelwise_augmenters_and_control_flow_parse():
    """AddElementwise / MultiplyElementwise / Dropout / AdditiveGaussianNoise / Grayscale (schemas/augmenters.raml:54-60, 88-96,
    120-122), Sequential (spliced in place) and OneOf (one member per sample) build the pixel-wise colour stage in YAML order."""
    from segmentation_training_pipeline_b200 import lib
    from segmentation_training_pipeline_b200.segmentation import parse_augmentation
    c = parse_augmentation({"Fliplr": 0.5, "Sequential": [{"Multiply": [0.9, 1.1]}, {"AdditiveGaussianNoise": {"scale": [0, 12.75], "per_channel": 0.5}}],
                            "OneOf": [{"Dropout": {"p": [0, 0.1]}}, {"Grayscale": {"alpha": [0, 1]}}, {"AddElementwise": [-10, 10]}]})
    kinds = [o[0] for o in c.pix_ops]
    assert kinds == [lib.PIX_KINDS[n] for n in ("Multiply", "AdditiveGaussianNoise", "Dropout", "Grayscale", "AddElementwise")]
    assert c.multiply == (0.9, 1.1) and c.fliplr == 0.5 and c.enabled()
    assert [o[4:] for o in c.pix_ops[2:]] == [(0, 3, 0), (0, 3, 1), (0, 3, 2)] and c.pix_ops[1][1] == 0.5
    spec = c.pix_c()
    assert spec.n_ops == 5 and spec.ops[1].kind == 6 and abs(spec.ops[1].b - 12.75) < 1e-6 and spec.ops[4].group_member == 2
    assert parse_augmentation({"Multiply": [0.9, 1.1], "Add": [-3, 3]}).pix_ops == ()        # the fused colour stage serves these alone
    assert parse_augmentation({"Dropout": 0.1}).pix_ops == ((5, 0.0, 0.1, 0.1, 0, 0, 0),)
    with pytest.raises(NotImplementedError, match="OneOf"):
        parse_augmentation({"OneOf": [{"Fliplr": 0.5}, {"Flipud": 0.5}]})
    with pytest.raises(NotImplementedError, match="twice"):
        parse_augmentation({"Add": [-3, 3], "Sequential": [{"Add": [-1, 1]}]})
    with pytest.raises(NotImplementedError, match="not fused"):
        parse_augmentation({"Sequential": [{"PiecewiseAffine": 0.05}]})
    assert parse_augmentation({"Sequential": [{"GaussianBlur": 1.0}]}).colour_seq[0][0] == "nb"
    with pytest.raises(NotImplementedError, match="order"):
        parse_augmentation({"Dropout": 0.1, "Fliplr": 0.5})
