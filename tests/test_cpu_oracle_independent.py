"""Pins the oracle's DeepLabV3 / MobileNetV2 graph and layer arithmetic against an INDEPENDENT implementation that is present in
this image: Hugging Face `transformers.MobileNetV2ForSemanticSegmentation` (a port of the same TF-slim MobileNetV2 + DeepLabV3
checkpoints that the reference's `impl/deeplab/model.py` ports for Keras).

With identical weights the two must agree layer by layer: TF 'same' padding of the stride-2 convolutions (asymmetric on even
sizes), the output-stride-8 schedule (block 6 keeps stride 1 at rate 1, blocks 7-13 rate 2, 14-16 rate 4 -- model.py:404-433),
depthwise convolutions with atrous rates, ReLU6, inference- and training-mode BatchNorm, the residual rule (stride 1 and equal
widths), the two-branch ASPP (image pooling broadcast + 1x1), `concat_projection` and the biased class layer at 1/8 resolution.
One constant differs between the two sources and is set from the reference's own file: the head BatchNorm epsilon
(1e-5, model.py:443-471; HF uses 1e-3 everywhere) -- nothing else is adjusted.  CPU only, no GPU and no network."""
import numpy as np
import pytest
import torch

def _hf_model(num_labels, train):
    pytest.importorskip("transformers")   # (imported lazily: collecting this module for a `-m gpu` run must stay cheap)
    from transformers import MobileNetV2Config, MobileNetV2ForSemanticSegmentation
    cfg = MobileNetV2Config(output_stride=8, num_labels=num_labels, classifier_dropout_prob=0.0, tf_padding=True, hidden_act="relu6",
                            depth_multiplier=1.0, first_layer_is_expansion=True, layer_norm_eps=1e-3)
    m = MobileNetV2ForSemanticSegmentation(cfg)
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.Conv2d):
                fan_in = mod.weight.shape[1] * mod.weight.shape[2] * mod.weight.shape[3]
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * (1.6 / np.sqrt(fan_in)))
                if mod.bias is not None:
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            elif isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.2)
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.3)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
        for name in ("conv_pool", "conv_aspp", "conv_projection"):   # model.py:443, 449, 471: epsilon=1e-5 in the DeepLab head
            getattr(m.segmentation_head, name).normalization.eps = 1e-5
    return m.train() if train else m.eval()


def _copy_into_oracle(m, om):
    """HF module -> the oracle's Keras-named parameters (kernel HWIO, depthwise_kernel (k, k, C, 1)) and BatchNorm buffers"""
    assigned = set()

    def put(key, value):
        assert tuple(om.params[key].shape) == tuple(value.shape), (key, om.params[key].shape, value.shape)
        om.params[key].data.copy_(value)
        assigned.add(key)

    def conv(layer, name, depthwise=False):
        w = layer.convolution.weight.detach()
        if depthwise:
            put(name + "/depthwise_kernel", w.permute(2, 3, 0, 1))
        else:
            put(name + "/kernel", w.permute(2, 3, 1, 0))
        if layer.convolution.bias is not None:
            put(name + "/bias", layer.convolution.bias.detach())
        bn = layer.normalization
        if bn is not None:
            put(name + "_BN/gamma", bn.weight.detach())
            put(name + "_BN/beta", bn.bias.detach())
            om.buffers[name + "_BN/moving_mean"].copy_(bn.running_mean)
            om.buffers[name + "_BN/moving_variance"].copy_(bn.running_var)
    mv = m.mobilenet_v2
    conv(mv.conv_stem.first_conv, "Conv")
    conv(mv.conv_stem.conv_3x3, "expanded_conv_depthwise", depthwise=True)
    conv(mv.conv_stem.reduce_1x1, "expanded_conv_project")
    for i, layer in enumerate(mv.layer):
        pre = "expanded_conv_%d_" % (i + 1)
        conv(layer.expand_1x1, pre + "expand")
        conv(layer.conv_3x3, pre + "depthwise", depthwise=True)
        conv(layer.reduce_1x1, pre + "project")
    h = m.segmentation_head
    conv(h.conv_pool, "image_pooling")
    conv(h.conv_aspp, "aspp0")
    conv(h.conv_projection, "concat_projection")
    conv(h.classifier, "custom_logits_semantic")
    assert assigned == set(om.params.keys()), sorted(assigned ^ set(om.params.keys()))   # every oracle parameter came from HF


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("hw", [(64, 96), (65, 97), (80, 48)])
@pytest.mark.parametrize("train", [False, True])
def test_oracle_deeplab_mobilenetv2_matches_hf_transformers(hw, train):
    from oracle.models import SegModel
    h, w = hw
    n = 3 if train else 1
    m = _hf_model(num_labels=1, train=train)
    om = SegModel("DeepLabV3", "mobilenetv2", classes=1, activation="sigmoid", input_shape=(h, w, 3), storage="fp32",
                  update_moving=False, dropout=None)
    _copy_into_oracle(m, om)
    om.training = train
    g = torch.Generator().manual_seed(h * 1000 + w)
    x = torch.randn(n, h, w, 3, generator=g)
    with torch.no_grad():
        hf = m(x.permute(0, 3, 1, 2).contiguous(), output_hidden_states=True)
        om(x)
    # backbone, block by block: hidden_states[i] is the output of HF layer i == the reference's block id i + 1
    worst = 0.0
    for i, hs in enumerate(hf.hidden_states):
        pre = "expanded_conv_%d_" % (i + 1)
        tap = om.taps.get(pre + "add", om.taps.get(pre + "project_BN"))
        assert tap is not None and tuple(tap.shape) == tuple(hs.shape), (i, None if tap is None else tap.shape, hs.shape)
        e = _rel(tap, hs)
        worst = max(worst, e)
        assert e < 1e-4, (i, e)   # fp32 round-off of two summation orders through up to 50 layers (measured <= 8e-5)
    # head: the class layer at 1/8 resolution (the reference applies its activation there and then upsamples, model.py:494-500)
    z = om.taps["logits_small"]
    assert tuple(z.shape) == tuple(hf.logits.shape) == (n, 1, -(-h // 8), -(-w // 8))
    print("worst block rel err %.2e, logits rel err %.2e" % (worst, _rel(z, hf.logits)))
    assert _rel(z, hf.logits) < 2e-4, _rel(z, hf.logits)
    assert float((z - hf.logits).abs().max()) < 2e-4 * (1.0 + float(hf.logits.abs().max()))


def test_parameter_count_matches_hf_minus_the_unused_classifier_stem():
    """same trainable tensors: HF's extra 1280-channel conv_1x1 (+ BN) is the ImageNet classifier stem the DeepLab graph never uses"""
    from oracle.models import SegModel
    m = _hf_model(num_labels=1, train=False)
    om = SegModel("DeepLabV3", "mobilenetv2", classes=1, activation="sigmoid", input_shape=(64, 64, 3), storage="fp32")
    hf_n = sum(p.numel() for nme, p in m.named_parameters() if not nme.startswith("mobilenet_v2.conv_1x1"))
    assert hf_n == sum(p.numel() for p in om.params.values()) == 2108417   # == the Keras model's trainable count


def test_oracle_vgg16_encoder_matches_torchvision():
    """BASELINE configs[0] (U-Net / VGG16): the restated keras.applications.VGG16 feature extractor -- 13 biased 3x3 'same' convolutions
    + ReLU in five blocks (64-64, 128-128, 256x3, 512x3, 512x3), a 2x2/2 max pool after each, the last activation of every block
    as a decoder skip -- against torchvision.models.vgg16's independent definition of the same network, with identical weights"""
    tv = pytest.importorskip("torchvision")
    from oracle.models import SegModel
    net = tv.models.vgg16(weights=None).features.eval()
    g = torch.Generator().manual_seed(7)
    convs = [m for m in net if isinstance(m, torch.nn.Conv2d)]
    assert len(convs) == 13
    with torch.no_grad():
        for c in convs:
            c.weight.copy_(torch.randn(c.weight.shape, generator=g) * (1.4 / np.sqrt(c.weight.shape[1] * 9)))
            c.bias.copy_(torch.randn(c.bias.shape, generator=g) * 0.05)
    om = SegModel("Unet", "vgg16", classes=1, activation="sigmoid", input_shape=(64, 96, 3), storage="fp32")
    names = ["block%d_conv%d" % (b + 1, i + 1) for b, n in enumerate((2, 2, 3, 3, 3)) for i in range(n)]
    for name, c in zip(names, convs):
        assert tuple(om.params[name + "/kernel"].shape) == tuple(c.weight.permute(2, 3, 1, 0).shape), name
        om.params[name + "/kernel"].data.copy_(c.weight.detach().permute(2, 3, 1, 0))
        om.params[name + "/bias"].data.copy_(c.bias.detach())
    x = torch.rand(2, 64, 96, 3, generator=g) * 255.0            # raw 0..255, as the reference feeds it
    with torch.no_grad():
        feat, skips = om._vgg16(x.permute(0, 3, 1, 2).contiguous())
        want_skips, h = [], x.permute(0, 3, 1, 2).contiguous()
        for m in net:
            if isinstance(m, torch.nn.MaxPool2d):
                want_skips.append(h)
            h = m(h)
    assert len(skips) == 5 and _rel(feat, h) < 1e-5
    for got, want in zip(skips, want_skips[::-1]):                # the oracle lists skips deepest first
        assert tuple(got.shape) == tuple(want.shape) and _rel(got, want) < 1e-5
