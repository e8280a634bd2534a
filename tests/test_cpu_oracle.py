"""CPU suite: pins the oracle against what CAN be pinned here (cv2 4.13, Random123 KATs, closed forms, sklearn)
and checks graph bookkeeping.  The reference ships no tests/golden vectors."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import augment as OA
from oracle import losses as OL
from oracle import nn as ON
from oracle import optim as OO
from oracle import philox
from oracle.models import SegModel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_random123_known_answers():
    assert philox.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox.philox4x32((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox.philox4x32((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


@pytest.mark.parametrize("hw", [(128, 128), (97, 131), (256, 256)])
def test_fixedpoint_warp_equals_cv2(hw):
    """the numpy restatement of cv2.warpAffine's fixed-point rule (what the CUDA kernel implements) is bit exact."""
    H, W = hw
    rng = np.random.default_rng(0)
    spec = OA.AugSpec(affine=True, scale=(0.8, 1.5), translate_x=(-0.2, 0.2), translate_y=(-0.2, 0.2),
                      rotate=(-16, 16), shear=(-16, 16))
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    msk = rng.integers(0, 2, (H, W, 1), dtype=np.uint8)
    for s in range(10):
        p = OA.draw_params(spec, 7, s, 3, H, W)
        assert np.array_equal(OA.warp_cv2(img, p.matrix, False), OA.warp_fixedpoint(img, p.matrix, False))
        assert np.array_equal(OA.warp_cv2(msk, p.matrix, True), OA.warp_fixedpoint(msk, p.matrix, True))


def test_augment_golden_fixture():
    """committed golden vectors (tests/golden/make_augment_golden.py) reproduce."""
    d = np.load(os.path.join(GOLD, "augment_golden.npz"))
    spec = OA.AugSpec(fliplr=0.5, flipud=0.5, affine=True, scale=(0.8, 1.5), translate_x=(-0.2, 0.2),
                      translate_y=(-0.2, 0.2), rotate=(-16, 16), shear=(-16, 16), multiply=(0.8, 1.2), add=(-10, 10))
    oi, om = OA.augment_batch(d["images"], d["masks"], spec, int(d["seed"]), int(d["step"]))
    assert np.array_equal(oi, d["out_images"]) and np.array_equal(om, d["out_masks"])
    for n in range(d["images"].shape[0]):
        p = OA.draw_params(spec, int(d["seed"]), int(d["step"]), n, 64, 64)
        assert np.allclose(p.matrix, d["matrices"][n], rtol=0, atol=1e-12)


def test_identity_affine_is_identity():
    img = np.random.default_rng(1).integers(0, 256, (33, 47, 3), dtype=np.uint8)
    M = np.array([[1.0, 0, 0], [0, 1.0, 0]])
    assert np.array_equal(OA.warp_fixedpoint(img, M, False), img)
    assert np.array_equal(OA.warp_fixedpoint(img, M, True), img)


def test_loss_known_answers():
    t = torch.tensor([0.0, 1.0, 0.0, 1.0]).view(1, 2, 2, 1)
    p = torch.full((1, 2, 2, 1), 0.5)
    assert abs(float(OL.binary_crossentropy(t, p)) - math.log(2)) < 1e-6
    assert abs(float(OL.dice_loss(t, t))) < 1e-6
    assert abs(float(OL.dice(t, p)) - (2 * 1.0 + 1) / (2 + 2 + 1)) < 1e-6
    assert abs(float(OL.iou(t, t)) - 1.0) < 1e-6
    # Lovasz hinge, 4-pixel hand example (Berman, relu): labels 1,0,1,0 ; logits 2,-2,-0.5,0.5
    lab = torch.tensor([1.0, 0.0, 1.0, 0.0])
    lg = torch.tensor([2.0, -2.0, -0.5, 0.5])
    # errors = 1 - lg*sign = [-1,-1,1.5,1.5]; sorted desc: 1.5(gt1),1.5(gt0),-1,-1 ; jaccard grads: 0.5, 1/6, ...
    val = float(OL.lovasz_hinge_flat(lg, lab, act="relu"))
    assert abs(val - (1.5 * 0.5 + 1.5 * (2.0 / 3 - 0.5))) < 1e-6
    assert OL.parse_composite("binary_crossentropy+0.1*dice_loss") == [(1.0, "binary_crossentropy"), (0.1, "dice_loss")]


def test_keras_adam_first_step():
    p = {"w": torch.tensor([1.0, -2.0, 3.0])}
    o = OO.Adam(p, lr=1e-3)
    g = torch.tensor([0.5, -0.25, 1e-3])
    o.step({"w": g})
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expect = torch.tensor([1.0, -2.0, 3.0]) - lr_t * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-7)
    assert torch.allclose(p["w"], expect, atol=1e-9)
    assert abs(float(p["w"][0]) - (1.0 - 1e-3)) < 1e-6  # ~ lr * sign(g)


def test_same_padding_rule_and_transpose():
    assert ON.keras_same_pad(512, 3, 1) == (1, 1)
    assert ON.keras_same_pad(512, 3, 2) == (0, 1)
    assert ON.keras_same_pad(512, 7, 2) == (2, 3)
    x = torch.randn(1, 4, 6, 6)
    w4 = torch.randn(4, 4, 5, 4)  # (kh,kw,Cout,Cin)
    assert ON.conv2d_transpose(x, w4, None, 2).shape == (1, 5, 12, 12)
    w3 = torch.randn(3, 3, 5, 4)
    assert ON.conv2d_transpose(x, w3, None, 2).shape == (1, 5, 12, 12)


def test_tf1_bilinear_legacy_rule():
    x = torch.arange(4.0).view(1, 1, 1, 4)
    y = ON.resize_bilinear_tf1(x, 1, 8)
    assert torch.allclose(y.flatten(), torch.tensor([0, 0.5, 1, 1.5, 2, 2.5, 3, 3]))
    ya = ON.resize_bilinear_tf1(x, 1, 7, align_corners=True)
    assert torch.allclose(ya.flatten(), torch.tensor([0, 0.5, 1, 1.5, 2, 2.5, 3]))


@pytest.mark.parametrize("arch,bb,expected", [("Unet", "resnet34", 24421456), ("Unet", "vgg16", 19030608),
                                              ("FPN", "resnet50", 28571328)])
def test_model_conv_param_counts(arch, bb, expected):
    """24.42 M / 19.03 M / 28.58 M conv parameters."""
    m = SegModel(arch, bb, classes=1)
    assert sum(p.numel() for k, p in m.params.items() if k.endswith("kernel")) == expected


def test_unet_resnet34_layer_inventory():
    m = SegModel("Unet", "resnet34", classes=1)
    convs = [k for k in m.params if k.endswith("/kernel")]
    assert len(convs) == 48
    assert m.params["conv0/kernel"].shape == (7, 7, 3, 64)
    assert m.params["decoder_stage0_conv1/kernel"].shape == (3, 3, 768, 256)
    assert m.params["decoder_stage3_conv1/kernel"].shape == (3, 3, 128, 32)
    assert m.params["decoder_stage4_conv1/kernel"].shape == (3, 3, 32, 16)
    assert "bn_data/gamma" not in m.params and "bn_data/beta" in m.params


def test_bf16_storage_mode_close_to_fp32():
    torch.manual_seed(0)
    a = SegModel("Unet", "resnet18", classes=1, storage="fp32")
    b = SegModel("Unet", "resnet18", classes=1, storage="bf16")
    b.load_numpy(a.state_numpy())
    x = torch.rand(2, 64, 64, 3) * 255
    with torch.no_grad():
        ya, yb = a(x), b(x)
    # bf16 storage is NOT within 1e-3 of fp32 through 40+ layers (tiny-batch BN amplifies it): this is why the
    # engine is compared with the bf16-emulating oracle; here only a sanity bound on the mean deviation.
    assert float((ya - yb).abs().mean()) < 0.02


def test_kfold_split_is_sklearn():
    from sklearn.model_selection import KFold
    folds = list(KFold(n_splits=5, shuffle=True, random_state=33).split(np.arange(20)))
    assert len(folds) == 5 and all(len(te) == 4 for _, te in folds)


def test_resize_oracle_is_pinned_against_real_cv2():
    """oracle/resize.py restates cv::resize's scalar arithmetic; the real cv2 of this image (IPP off) agrees except at rounding
    ties of its SIMD vertical pass (fp32, round-half-even vs integer round-half-up): <= 0.05 % of pixels, |d| <= 1.  Nearest is
    bit exact."""
    import cv2
    from oracle import resize as OR
    rng = np.random.default_rng(5)
    prev = cv2.ipp.useIPP() if hasattr(cv2, "ipp") else None
    try:
        if prev is not None:
            cv2.ipp.setUseIPP(False)
        tot = bad = 0
        for (h, w, H, W) in [(37, 53, 64, 64), (300, 200, 128, 160), (64, 64, 200, 333), (100, 100, 37, 41), (96, 96, 192, 192),
                             (128, 96, 64, 48), (33, 65, 512, 512)]:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            ref = cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC)
            got = OR.resize_cubic_u8(img, H, W)
            d = np.abs(got.astype(int) - ref.astype(int))
            assert d.max() <= 1, (h, w, H, W, d.max())
            tot += d.size
            bad += int((d != 0).sum())
            m = (rng.random((h, w)) > 0.6).astype(np.uint8)
            assert np.array_equal(OR.resize_nearest_u8(m, H, W), cv2.resize(m, (W, H), interpolation=cv2.INTER_NEAREST)), (h, w, H, W)
        assert bad <= 5e-4 * tot, (bad, tot)
        # smooth images (photographs, not noise) hit fewer ties still; padding window helper == np.pad / slicing
        img = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
        assert np.array_equal(OR.window(img, -3, -2, 28, 40), np.pad(img, ((3, 5), (2, 8), (0, 0)))[:28, :40])
        assert np.array_equal(OR.window(img, 4, 5, 10, 12), img[4:14, 5:17])
    finally:
        if prev is not None:
            cv2.ipp.setUseIPP(prev)


def test_deeplab_oracle_graph_follows_the_reference_source():
    """DeepLabV3+/MobileNetV2 restated from the reference's IN-TREE impl/deeplab/model.py: Keras layer names, the Keras model's
    trainable parameter count, output = align_corners resize of the low-resolution probabilities, every parameter reached by the
    gradient."""
    import torch
    from oracle.models import SegModel
    m = SegModel("DeepLabV3", "mobilenetv2", classes=1, input_shape=(64, 64, 3), seed=1)
    assert sum(p.numel() for p in m.params.values()) == 2108417
    for k, shape in {"Conv/kernel": (3, 3, 3, 32), "expanded_conv_depthwise/depthwise_kernel": (3, 3, 32, 1),
                     "expanded_conv_1_expand/kernel": (1, 1, 16, 96), "expanded_conv_16_project/kernel": (1, 1, 960, 320),
                     "image_pooling/kernel": (1, 1, 320, 256), "aspp0/kernel": (1, 1, 320, 256),
                     "concat_projection/kernel": (1, 1, 512, 256), "custom_logits_semantic/kernel": (1, 1, 256, 1)}.items():
        assert tuple(m.params[k].shape) == shape, k
    assert "expanded_conv_expand/kernel" not in m.params            # block 0 has no expansion (model.py:241-251)
    assert "expanded_conv_16_project_BN" in m.P.encoder_names and "aspp0" not in m.P.encoder_names
    x = torch.rand(2, 64, 64, 3) * 255
    y = m(x)
    assert y.shape == (2, 64, 64, 1) and float(y.min()) > 0 and float(y.max()) < 1
    z = m.taps["logits_small"]
    assert z.shape == (2, 1, 8, 8)                                   # output stride 8
    ref = torch.nn.functional.interpolate(torch.sigmoid(z), size=(64, 64), mode="bilinear", align_corners=True)
    assert float((y.permute(0, 3, 1, 2) - ref).abs().max()) < 1e-6   # activation BEFORE the resize (model.py:499-500)
    y.mean().backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.params.values())
    # moving statistics use momentum 0.999 in the backbone, 0.99 in the ASPP head
    assert abs(float(m.buffers["Conv_BN/moving_variance"].mean()) - 1.0) > 1e-6
    m2 = SegModel("DeepLabV3", "mobilenetv2", classes=3, activation="softmax", input_shape=(32, 48, 3))
    p = m2(torch.rand(1, 32, 48, 3) * 255)
    assert p.shape == (1, 32, 48, 3) and float((p.sum(-1) - 1).abs().max()) < 1e-5


def test_deeplab_xception_oracle_graph():
    """modified aligned Xception + ASPP + decoder (impl/deeplab/model.py:339-383, 457-500): parameter count of the Keras model
    (41 050 273 trainable for one class), Keras layer names, explicit (1, 1) padding of the stride-2 separable convs"""
    import torch
    from oracle.models import SegModel
    m = SegModel("DeepLabV3", "xception", classes=1, input_shape=(64, 64, 3), seed=1)
    assert sum(p.numel() for p in m.params.values()) == 41050273
    for k in ("entry_flow_conv1_1/kernel", "entry_flow_block2_shortcut/kernel", "middle_flow_unit_16_separable_conv3_pointwise/kernel",
              "exit_flow_block2_separable_conv3_pointwise_BN/gamma", "aspp3_depthwise/depthwise_kernel", "feature_projection0_BN/beta",
              "decoder_conv1_pointwise/kernel", "custom_logits_semantic/bias"):
        assert k in m.params, k
    y = m(torch.rand(1, 64, 64, 3) * 255)
    assert y.shape == (1, 64, 64, 1) and m.taps["logits_small"].shape == (1, 1, 16, 16)      # decoder works at 1/4 resolution
    assert m.taps["exit_flow_block1"].shape[-1] == 4                                           # output stride 16
    m8 = SegModel("DeepLabV3", "xception", classes=1, input_shape=(64, 64, 3), OS=8)
    m8(torch.rand(1, 64, 64, 3) * 255)
    assert m8.taps["exit_flow_block1"].shape[-1] == 8


def test_depthwise_same_padding_and_dropout_mask():
    import torch
    from oracle import nn as L
    from oracle.philox import dropout_keep_mask
    w = torch.ones(3, 3, 8, 1)
    for size, stride, rate, out in ((16, 2, 1, 8), (15, 2, 1, 8), (12, 1, 4, 12)):
        y = L.depthwise_conv2d(torch.ones(1, 8, size, size), w, stride, rate)
        assert y.shape[-1] == out
    # even size, stride 2: TF pads 0 before / 1 after -> the first output sees the full 3x3 window, the last one loses a row+column
    y = L.depthwise_conv2d(torch.ones(1, 8, 16, 16), w, 2, 1)
    assert float(y[0, 0, 0, 0]) == 9.0 and float(y[0, 0, -1, -1]) == 4.0
    a = dropout_keep_mask(50, 16, 0.1, 7, 0xD0, 3)
    assert np.array_equal(a, dropout_keep_mask(50, 16, 0.1, 7, 0xD0, 3))
    assert not np.array_equal(a, dropout_keep_mask(50, 16, 0.1, 7, 0xD0, 4))
    assert abs(dropout_keep_mask(4000, 16, 0.1, 1, 2, 3).mean() - 0.9) < 0.01


def test_neighbourhood_arithmetic_is_pinned_against_real_cv2():
    """The cv2 calls behind imgaug's GaussianBlur / AverageBlur / MedianBlur / Sharpen / Emboss / EdgeDetect, restated in
    oracle/augment.py (and in csrc/augment_nb.cu), are bit exact against the real cv2 of this image, with Intel IPP on and off:
    cv2.GaussianBlur's 8.8 fixed-point separable kernel (error-diffused coefficients, (sum + 2^15) >> 16), cv2.blur for k = 2..17
    (round half up, one residue earlier for powers of two), cv2.medianBlur (k = 3, 5, 7; replicated border), cv2.filter2D with
    a float32 3x3 matrix (row-major float32 accumulation, rint)."""
    import cv2
    from oracle import augment as OA
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    img[:8, :8] = 255
    img[-8:, -8:] = 0
    ipp0 = cv2.ipp.useIPP()
    try:
        for ipp in (True, False):
            cv2.ipp.setUseIPP(ipp)
            for sigma in (0.4, 1.0, 1.9, 2.99, 3.0, 3.5, 4.99, 5.5, 8.8):
                ks = OA.gaussian_ksize_imgaug(sigma)
                ref = cv2.GaussianBlur(img, (ks, ks), sigmaX=sigma, sigmaY=sigma, borderType=cv2.BORDER_REFLECT_101)
                assert np.array_equal(OA.gaussian_blur_u8(img, ks, sigma), ref), ("gauss", sigma, ks, ipp)
            for k in range(2, 18):
                assert np.array_equal(OA.average_blur_u8(img, k), cv2.blur(img, (k, k))), ("blur", k, ipp)
            for k in (3, 5, 7):
                assert np.array_equal(OA.median_blur_u8(img, k), cv2.medianBlur(img, k)), ("median", k, ipp)
            for kind, alpha, second in ((3, 0.3, 1.2), (3, 1.0, 0.75), (4, 0.8, 1.7), (4, 0.05, 0.3), (5, 0.55, 0.0), (5, 0.0, 0.0)):
                act, par = OA.neighbourhood_params((kind, alpha, alpha, second, second, 2, 0, 0, 0), 1, 2, 3)
                assert act and par[0] == "filter"
                ref = np.stack([cv2.filter2D(np.ascontiguousarray(img[..., c]), -1, par[1]) for c in range(3)], -1)
                assert np.array_equal(OA.filter2d_3x3_u8(img, par[1]), ref), ("filter2D", kind, alpha, ipp)
    finally:
        cv2.ipp.setUseIPP(ipp0)
    assert OA.gaussian_kernel_fixed(5, 1.0).tolist() == [14, 62, 104, 62, 14]      # sum 256, centre takes the remainder
    assert [OA.gaussian_ksize_imgaug(s) for s in (0.5, 2.0, 3.0, 4.0, 6.0)] == [5, 7, 9, 11, 15]


@pytest.mark.parametrize("shape,out", [((2, 3, 7, 9), (20, 31)), ((1, 4, 40, 40), (320, 320)), ((1, 2, 33, 17), (9, 5)), ((2, 1, 1, 1), (6, 4))])
def test_align_corners_resize_is_torch_interpolate(shape, out):
    """the align_corners=True bilinear resize of impl/deeplab/model.py:92-100 (the network's final upsampling) against torch's
    independent implementation of the same published rule (scale = (in - 1) / (out - 1))"""
    import torch.nn.functional as F
    x = torch.randn(shape, generator=torch.Generator().manual_seed(sum(shape)))
    got = ON.resize_bilinear_tf1(x, out[0], out[1], align_corners=True)
    want = F.interpolate(x, size=out, mode="bilinear", align_corners=True)
    assert float((got - want).abs().max()) < 2e-6 * (1 + float(want.abs().max()))


def test_batchnorm_restatement_is_torch_batch_norm():
    """training mode normalises with the BIASED batch variance, inference mode with the moving statistics: torch.nn.functional.batch_norm"""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, 12, 9, 7, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(12, generator=g) + 0.5, torch.randn(12, generator=g)
    y, mean, var = ON.batchnorm_train(x, gamma, beta, 1e-3)
    want = F.batch_norm(x, None, None, gamma, beta, training=True, eps=1e-3)
    assert float((y - want).abs().max()) < 2e-6 * (1 + float(want.abs().max()))
    assert torch.allclose(mean, x.mean(dim=(0, 2, 3)), atol=1e-6) and torch.allclose(var, x.var(dim=(0, 2, 3), unbiased=False), atol=1e-6)
    mm, mv = torch.randn(12, generator=g), torch.rand(12, generator=g) + 0.5
    yi = ON.batchnorm_infer(x, gamma, beta, mm, mv, 1e-5)
    wi = F.batch_norm(x, mm, mv, gamma, beta, training=False, eps=1e-5)
    assert float((yi - wi).abs().max()) < 2e-6 * (1 + float(wi.abs().max()))


def test_tf_same_padding_rule_is_the_one_hf_transformers_ports():
    """keras 'same' padding (total = max((ceil(n/s) - 1)*s + k_eff - n, 0), the smaller half BEFORE) against the rule Hugging Face
    ported from TensorFlow for its TF-checkpoint models (apply_tf_padding), on even / odd sizes, strides and atrous rates"""
    mod = pytest.importorskip("transformers.models.mobilenet_v2.modeling_mobilenet_v2")
    import torch.nn as nn
    for n in (7, 8, 15, 16, 33, 64):
        for k, s, d in ((3, 1, 1), (3, 2, 1), (3, 1, 2), (3, 1, 4), (1, 1, 1), (1, 2, 1)):
            conv = nn.Conv2d(1, 1, k, stride=s, dilation=d)
            padded = mod.apply_tf_padding(torch.zeros(1, 1, n, n), conv)
            x = torch.zeros(1, 1, n, n)
            x[0, 0, 0, 0] = 1.0                                   # where the first input pixel lands = the padding BEFORE
            first = int(torch.nonzero(mod.apply_tf_padding(x, conv))[0][2])
            total = padded.shape[2] - n
            before, after = ON.keras_same_pad(n, (k - 1) * d + 1, s)
            assert (first, total - first) == (before, after), (n, k, s, d, (first, total - first), (before, after))


@pytest.mark.parametrize("cin,cout,G,w", [(16, 16, 4, 16), (32, 16, 4, 24), (16, 32, 4, 8), (32, 32, 2, 10)])
def test_pixel_merged_view_prototype(cin, cout, G, w):
    """the block-sparse expanded weights of the planned narrow-layer kernel (scripts/quadview_proto.py): a 3x3 'same' conv
    equals the 3x3 conv over the [N, H, W/G, G*C] VIEW with them; every K = 16 step feeds a contiguous run of output-pixel blocks"""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("quadview_proto", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "quadview_proto.py"))
    qp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(qp)
    assert qp.check(n=2, h=7, w=w, cin=cin, cout=cout, G=G, seed=cin + cout) < 1e-10
    sched = qp.mma_schedule(cin, cout, G)
    assert all(1 <= nb <= 3 for steps in sched.values() for _, _, nb in steps)
    assert len(sched[-1]) == len(sched[1]) == cin // 16        # only the nearest pixel of a neighbouring group contributes
