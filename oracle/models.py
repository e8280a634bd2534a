"""Model graphs the reference builds at segmentation.py:96-155 (createNet1) restated on the oracle layers.

`segmentation_models.Unet/FPN/Linknet` (==0.2.1, requires.txt:15) and `classification_models` ResNets
are un-vendored [DEP]; the graphs below follow SURVEY.md section 8 a-4/a-5 and Appendix A/B.
TEST INFRASTRUCTURE -- parity unpinned (no reference test pins these graphs), with one exception: the DeepLabV3 / MobileNetV2
graph (restated from the reference's in-tree impl/deeplab/model.py) is pinned block by block against Hugging Face transformers'
independent MobileNetV2 + DeepLabV3 implementation, and the VGG16 encoder against torchvision.models.vgg16 (tests/test_cpu_oracle_independent.py).

A model is (params: Dict[str, Tensor in Keras layout], forward(x_nhwc_float, training) -> y_nhwc).
Layer names are the Keras names so weight dicts are exchangeable with the engine.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import nn as L

ENC_BN_EPS = 2e-5  # classification_models get_bn_params(): epsilon=2e-5, momentum=0.99
DEC_BN_EPS = 1e-3  # keras.layers.BatchNormalization default epsilon
BN_MOMENTUM = 0.99

RESNET_REPS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3),
               "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
RESNET_BOTTLENECK = {"resnet18": False, "resnet34": False, "resnet50": True, "resnet101": True, "resnet152": True}
VGG16_BLOCKS = ((64, 2), (128, 2), (256, 3), (512, 3), (512, 3))


class ParamStore:
    """Creates parameters on first use, in graph order, from one numpy Generator (seeded)."""

    def __init__(self, seed: int = 0, enc_init="he_uniform", dec_init="glorot_uniform"):
        self.gen = np.random.default_rng(seed)
        self.params: Dict[str, torch.Tensor] = {}
        self.buffers: Dict[str, torch.Tensor] = {}  # BN moving stats
        self.enc_init, self.dec_init = enc_init, dec_init
        self.encoder_names: List[str] = []
        self._in_encoder = True

    def conv(self, name, kh, kw, cin, cout, bias, init=None):
        if name + "/kernel" not in self.params:
            init = init or (self.enc_init if self._in_encoder else self.dec_init)
            fn = L.he_uniform if init == "he_uniform" else L.glorot_uniform
            self.params[name + "/kernel"] = torch.from_numpy(fn((kh, kw, cin, cout), self.gen))
            if bias:
                self.params[name + "/bias"] = torch.zeros(cout)
            if self._in_encoder:
                self.encoder_names.append(name)
        return self.params[name + "/kernel"], self.params.get(name + "/bias")

    def depthwise(self, name, k, c):
        """keras DepthwiseConv2D kernel (k, k, C, 1), glorot_uniform (fan_in k*k*C, fan_out k*k)"""
        key = name + "/depthwise_kernel"
        if key not in self.params:
            lim = float(np.sqrt(6.0 / (k * k * c + k * k)))
            self.params[key] = torch.from_numpy(self.gen.uniform(-lim, lim, size=(k, k, c)).astype(np.float32)[..., None])
            if self._in_encoder:
                self.encoder_names.append(name)
        return self.params[key]

    def bn(self, name, c, scale=True):
        if name + "/beta" not in self.params:
            if scale:
                self.params[name + "/gamma"] = torch.ones(c)
            self.params[name + "/beta"] = torch.zeros(c)
            self.buffers[name + "/moving_mean"] = torch.zeros(c)
            self.buffers[name + "/moving_variance"] = torch.ones(c)
            if self._in_encoder:
                self.encoder_names.append(name)
        return self.params.get(name + "/gamma"), self.params[name + "/beta"]


class SegModel:
    """U-Net / FPN / Linknet over ResNet-{18,34,50,..} or VGG16, Keras semantics, NHWC at the boundary."""

    def __init__(self, architecture="Unet", backbone="resnet34", classes=1, activation="sigmoid",
                 input_shape=(512, 512, 3), seed=0, storage="fp32",
                 decoder_filters=(256, 128, 64, 32, 16), decoder_use_batchnorm=True,
                 decoder_block_type="upsampling", enc_init="he_uniform", dec_init="glorot_uniform",
                 pyramid_block_filters=256, segmentation_block_filters=128, fpn_dropout=None,
                 update_moving=True, downsample_factor=8, psp_conv_filters=512, dropout=None, OS=16):
        self.arch, self.backbone = architecture, backbone.lower()
        self.classes, self.activation = classes, activation
        self.storage = storage
        self.decoder_filters = tuple(decoder_filters)
        self.dec_bn = decoder_use_batchnorm
        self.block_type = decoder_block_type
        self.pyr, self.segf = pyramid_block_filters, segmentation_block_filters
        self.psp_factor, self.psp_filters = int(downsample_factor), int(psp_conv_filters)
        # DeepLabV3 only: None (no dropout) or (rate, seed, salt, step) -- the engine's Philox mask (oracle/philox.py)
        self.dropout = dropout
        self.OS = int(OS)   # DeepLabV3 / xception only (model.py:339-349); mobilenetv2 is always output stride 8
        if self.arch == "DeepLabV3":
            enc_init = dec_init = "glorot_uniform"   # keras defaults of impl/deeplab/model.py
        self.P = ParamStore(seed, enc_init, dec_init)
        self.training = True
        self.update_moving = update_moving
        self.taps: Dict[str, torch.Tensor] = {}
        # materialise parameters with a dry run on a tiny input of the right channel count
        h = 64 if self.backbone != "vgg16" else 32
        if self.arch == "PSPNet":
            h = 6 * self.psp_factor
        if self.arch == "DeepLabV3":
            h = 16
        um, self.update_moving = self.update_moving, False
        with torch.no_grad():
            self.forward(torch.zeros(1, h, h, input_shape[2]), emit_logits=False)
        self.update_moving = um
        # storage="fp64": the same graph in double precision -- the ANCHOR against which both the fp32 oracle and the engine's
        # fp32 parity mode are measured (how far apart two correct fp32 implementations of this training run can be)
        self.dtype = torch.float64 if storage == "fp64" else torch.float32
        if self.dtype is torch.float64:
            for d in (self.P.params, self.P.buffers):
                for k in list(d):
                    d[k] = d[k].double()
        for p in self.P.params.values():
            p.requires_grad_(True)

    # -- helpers ---------------------------------------------------------------------------
    @property
    def params(self):
        return self.P.params

    @property
    def buffers(self):
        return self.P.buffers

    def _bn(self, x, name, eps, scale=True, momentum=BN_MOMENTUM):
        gamma, beta = self.P.bn(name, x.shape[1], scale)
        if self.training:
            y, mean, var = L.batchnorm_train(x, gamma, beta, eps)
            if self.update_moving:
                with torch.no_grad():
                    m = x.numel() // x.shape[1]
                    # keras 2.2.x normalization.py: variance *= sample_size / (sample_size - (1.0 + epsilon))
                    unbiased = var * (m / (m - (1.0 + eps))) if m > 1 else var
                    mm, mv = self.P.buffers[name + "/moving_mean"], self.P.buffers[name + "/moving_variance"]
                    mm.mul_(momentum).add_(mean.detach() * (1 - momentum))
                    mv.mul_(momentum).add_(unbiased.detach() * (1 - momentum))
            return y
        return L.batchnorm_infer(x, gamma, beta, self.P.buffers[name + "/moving_mean"],
                                 self.P.buffers[name + "/moving_variance"], eps)

    def _bn_relu(self, x, name, eps, tap=None):
        y = L.rb(torch.relu(self._bn(x, name, eps)), self.storage)
        if tap:
            self.taps[tap] = y
        return y

    def _conv(self, x, name, k, cout, stride=1, padding="valid", bias=False, residual=None):
        w, b = self.P.conv(name, k, k, x.shape[1], cout, bias)
        y = L.conv2d(x, w, b, stride, padding, self.storage)
        if residual is not None:
            y = y + residual
        return L.rb(y, self.storage)

    # -- encoders --------------------------------------------------------------------------
    def _resnet(self, x):
        reps, bott = RESNET_REPS[self.backbone], RESNET_BOTTLENECK[self.backbone]
        x = L.rb(self._bn(x, "bn_data", ENC_BN_EPS, scale=False), self.storage)
        x = self._conv(x, "conv0", 7, 64, stride=2, padding=3)
        x = self._bn_relu(x, "bn0", ENC_BN_EPS, tap="relu0")
        x = L.maxpool(x, 3, 2, 1)
        psp_tap = {4: "stage2_unit1_relu1", 8: "stage3_unit1_relu1", 16: "stage4_unit1_relu1"}.get(self.psp_factor) \
            if self.arch == "PSPNet" else None
        for stage, rep in enumerate(reps):
            f = 64 * 2 ** stage
            for block in range(rep):
                pre = "stage%d_unit%d_" % (stage + 1, block + 1)
                first = block == 0
                stride = 2 if (first and stage > 0) else 1
                y = self._bn_relu(x, pre + "bn1", ENC_BN_EPS, tap=pre + "relu1")
                if psp_tap == pre + "relu1":
                    return y, []   # PSPNet: the Keras Model ends the encoder at the feature layer (later layers are not in the graph)
                if bott:
                    sc = self._conv(y, pre + "sc", 1, 4 * f, stride) if first else x
                    z = self._conv(y, pre + "conv1", 1, f)
                    z = self._bn_relu(z, pre + "bn2", ENC_BN_EPS)
                    z = self._conv(z, pre + "conv2", 3, f, stride, padding=1)
                    z = self._bn_relu(z, pre + "bn3", ENC_BN_EPS)
                    x = self._conv(z, pre + "conv3", 1, 4 * f, residual=sc)
                else:
                    sc = self._conv(y, pre + "sc", 1, f, stride) if first else x
                    z = self._conv(y, pre + "conv1", 3, f, stride, padding=1)
                    z = self._bn_relu(z, pre + "bn2", ENC_BN_EPS)
                    x = self._conv(z, pre + "conv2", 3, f, padding=1, residual=sc)
        x = self._bn_relu(x, "bn1", ENC_BN_EPS, tap="relu1")
        skips = ["stage4_unit1_relu1", "stage3_unit1_relu1", "stage2_unit1_relu1", "relu0"]
        return x, [self.taps[s] for s in skips]

    def _vgg16(self, x):
        """keras.applications.VGG16 [DEP]: 3x3 same conv + bias + ReLU, 2x2/2 max pool; raw 0..255 input."""
        x = L.rb(x, self.storage)
        skips = []
        for bi, (f, n) in enumerate(VGG16_BLOCKS):
            for ci in range(n):
                name = "block%d_conv%d" % (bi + 1, ci + 1)
                w, b = self.P.conv(name, 3, 3, x.shape[1], f, True, init="glorot_uniform")
                x = L.rb(torch.relu(L.conv2d(x, w, b, 1, "same", self.storage)), self.storage)
            skips.append(x)
            x = L.maxpool(x, 2, 2, 0)
        return x, skips[::-1]

    # -- decoders --------------------------------------------------------------------------
    def _conv_bn_relu(self, x, cname, bname, k, cout, use_bn=True):
        z = self._conv(x, cname, k, cout, 1, "same", bias=not use_bn)
        if use_bn:
            return self._bn_relu(z, bname, DEC_BN_EPS)
        return L.rb(torch.relu(z), self.storage)

    def _unet_decoder(self, x, skips):
        for i, f in enumerate(self.decoder_filters):
            pre = "decoder_stage%d_" % i
            skip = skips[i] if i < len(skips) else None
            if self.block_type == "transpose":
                w, b = self.P.conv(pre + "transpose", 4, 4, f, x.shape[1], not self.dec_bn)  # (kh,kw,Cout,Cin)
                z = L.rb(L.conv2d_transpose(x, w, b, 2, self.storage), self.storage)
                x = self._bn_relu(z, pre + "bn1", DEC_BN_EPS) if self.dec_bn else L.rb(torch.relu(z), self.storage)
                if skip is not None:
                    x = torch.cat([x, skip], dim=1)
                x = self._conv_bn_relu(x, pre + "conv2", pre + "bn2", 3, f, self.dec_bn)
            else:
                x = L.upsample_nearest(x, 2)
                if skip is not None:
                    x = torch.cat([x, skip], dim=1)
                x = self._conv_bn_relu(x, pre + "conv1", pre + "bn1", 3, f, self.dec_bn)
                x = self._conv_bn_relu(x, pre + "conv2", pre + "bn2", 3, f, self.dec_bn)
        return x

    def _fpn_decoder(self, x, skips):
        """segmentation_models 0.2.1 FPN [DEP]: SURVEY.md 8 a-5."""
        # pyramid: top + 3 skips (H/32 .. H/4)
        feats = [x] + skips[:3]
        p = None
        pyramid = []
        for i, c in enumerate(feats):
            # Add(UpSampling2D(2)(previous level), lateral): one rounding of the sum (the engine adds in the conv epilogue)
            p = self._conv(c, "pyramid_stage_%d_conv1x1" % i, 1, self.pyr, bias=True,
                           residual=L.upsample_nearest(p, 2) if p is not None else None)
            pyramid.append(p)
        outs = []
        rates = (8, 4, 2, 1)
        for i, p in enumerate(pyramid):
            s = self._conv_bn_relu(p, "segm_stage_%d_conv1" % i, "segm_stage_%d_bn1" % i, 3, self.segf, self.dec_bn)
            s = self._conv_bn_relu(s, "segm_stage_%d_conv2" % i, "segm_stage_%d_bn2" % i, 3, self.segf, self.dec_bn)
            if rates[i] > 1:
                s = L.rb(L.resize_bilinear_tf1(s, s.shape[2] * rates[i], s.shape[3] * rates[i]), self.storage)
            outs.append(s)
        x = torch.cat(outs, dim=1)
        x = self._conv_bn_relu(x, "final_stage_conv", "final_stage_bn", 3, self.segf * 4, self.dec_bn)
        return x

    def _linknet_decoder(self, x, skips):
        """segmentation_models 0.2.1 Linknet [DEP]: 1x1(C/4) -> up x2 -> 3x3(C/4) -> 1x1(out) -> Add(skip)."""
        for i in range(5):
            pre = "decoder_stage%d_" % i
            skip = skips[i] if i < len(skips) else None
            cin = x.shape[1]
            cout = skip.shape[1] if skip is not None else self.decoder_filters[i] if self.decoder_filters[i] else 16
            z = self._conv_bn_relu(x, pre + "conv1", pre + "bn1", 1, cin // 4, self.dec_bn)
            z = L.upsample_nearest(z, 2)
            z = self._conv_bn_relu(z, pre + "conv2", pre + "bn2", 3, cin // 4, self.dec_bn)
            z = self._conv_bn_relu(z, pre + "conv3", pre + "bn3", 1, cout, self.dec_bn)
            x = L.rb(z + skip, self.storage) if skip is not None else z
        return x

    def _pspnet_decoder(self, feat):
        """segmentation_models 0.2.1 PSPNet [DEP, recalled] (schema segmentation.raml:226-248): pyramid pooling over the
        feature map at 1/downsample_factor -- for level in (1, 2, 3, 6): AveragePooling2D(size/level) -> 1x1 conv
        (psp_conv_filters) + BN + ReLU -> bilinear resize back (TF1 legacy); Concatenate([features, levels...]); 1x1 conv (512)
        + BN + ReLU; then final_conv 3x3 and a bilinear x downsample_factor upsample of the logits (in forward())."""
        h, w = feat.shape[2], feat.shape[3]
        outs = [feat]
        for level in (1, 2, 3, 6):
            k = h // level
            z = L.rb(torch.nn.functional.avg_pool2d(feat, k, k), self.storage)
            z = self._conv_bn_relu(z, "psp_level%d_conv" % level, "psp_level%d_bn" % level, 1, self.psp_filters, True)
            outs.append(L.rb(L.resize_bilinear_tf1(z, h, w), self.storage))
        x = torch.cat(outs, dim=1)
        return self._conv_bn_relu(x, "psp_conv", "psp_bn", 1, 512, True)

    # -- DeepLabV3+ / MobileNetV2 (reference impl/deeplab/model.py, in-tree) -------------------
    MOBILENETV2_BLOCKS = (
        (16, 1, 1, 0, False, 1),
        (24, 2, 6, 1, False, 1), (24, 1, 6, 2, True, 1),
        (32, 2, 6, 3, False, 1), (32, 1, 6, 4, True, 1), (32, 1, 6, 5, True, 1),
        (64, 1, 6, 6, False, 1), (64, 1, 6, 7, True, 2), (64, 1, 6, 8, True, 2), (64, 1, 6, 9, True, 2),
        (96, 1, 6, 10, False, 2), (96, 1, 6, 11, True, 2), (96, 1, 6, 12, True, 2),
        (160, 1, 6, 13, False, 2), (160, 1, 6, 14, True, 4), (160, 1, 6, 15, True, 4),
        (320, 1, 6, 16, False, 4),
    )

    def _bn_act(self, x, name, eps, act, momentum=BN_MOMENTUM, tap=None):
        y = self._bn(x, name, eps, momentum=momentum)
        if act == "relu6":      # keras relu(max_value=6) (model.py:38-39)
            y = torch.clamp(y, 0.0, 6.0)
        elif act == "relu":
            y = torch.relu(y)
        y = L.rb(y, self.storage)
        if tap:
            self.taps[tap] = y
        return y

    def _mobilenetv2(self, x):
        """model.py:386-433 (feature extractor, output stride 8) with `_inverted_res_block` :236-275"""
        EPS, MOM = 1e-3, 0.999
        x = self._conv(x, "Conv", 3, 32, stride=2, padding="same")
        x = self._bn_act(x, "Conv_BN", EPS, "relu6", MOM, tap="Conv_Relu6")
        for filters, stride, expansion, bid, skip, rate in self.MOBILENETV2_BLOCKS:
            pre = "expanded_conv_%d_" % bid if bid else "expanded_conv_"
            inp, t = x, x
            if bid:
                t = self._conv(t, pre + "expand", 1, expansion * inp.shape[1])
                t = self._bn_act(t, pre + "expand_BN", EPS, "relu6", MOM)
            wd = self.P.depthwise(pre + "depthwise", 3, t.shape[1])
            t = L.rb(L.depthwise_conv2d(t, wd, stride, rate, self.storage), self.storage)
            t = self._bn_act(t, pre + "depthwise_BN", EPS, "relu6", MOM)
            t = self._conv(t, pre + "project", 1, filters)
            t = self._bn_act(t, pre + "project_BN", EPS, None, MOM)
            x = L.rb(inp + t, self.storage) if skip else t
            self.taps[pre + ("add" if skip else "project_BN")] = x
        return x

    # -- DeepLabV3+ / modified aligned Xception (model.py:110-224, 339-383, 457-500) -----------
    def _sepconv_bn(self, x, filters, prefix, stride=1, rate=1, depth_activation=False, eps=1e-3):
        """model.py:110-147"""
        if not depth_activation:
            x = L.rb(torch.relu(x), self.storage)
        wd = self.P.depthwise(prefix + "_depthwise", 3, x.shape[1])
        ke = 3 + 2 * (rate - 1)
        pad = None if stride == 1 else ((ke - 1) // 2, (ke - 1) - (ke - 1) // 2)
        x = L.rb(L.depthwise_conv2d(x, wd, stride, rate, self.storage, explicit_pad=pad), self.storage)
        x = self._bn_act(x, prefix + "_depthwise_BN", eps, "relu" if depth_activation else None)
        x = self._conv(x, prefix + "_pointwise", 1, filters)
        return self._bn_act(x, prefix + "_pointwise_BN", eps, "relu" if depth_activation else None)

    def _xception_block(self, inputs, depth_list, prefix, skip_type, stride, rate=1, depth_activation=False):
        """model.py:177-216; returns (outputs, skip = the tensor after the second SepConv)"""
        residual, skip = inputs, None
        for i in range(3):
            residual = self._sepconv_bn(residual, depth_list[i], prefix + "_separable_conv%d" % (i + 1),
                                        stride=stride if i == 2 else 1, rate=rate, depth_activation=depth_activation)
            if i == 1:
                skip = residual
        if skip_type == "conv":
            sc = self._conv(inputs, prefix + "_shortcut", 1, depth_list[-1], stride=stride)   # k = 1: _conv2d_same pads nothing
            sc = self._bn_act(sc, prefix + "_shortcut_BN", 1e-3, None)
            out = L.rb(residual + sc, self.storage)
        elif skip_type == "sum":
            out = L.rb(residual + inputs, self.storage)
        else:
            out = residual
        self.taps[prefix] = out
        return out, skip

    def _xception(self, x):
        """model.py:339-383: entry / middle (16 units) / exit flow; OS 16: strides 2-2-2 then rates (1, (1, 2)); OS 8: the third
        entry stride becomes 1 with rates (2, (2, 4))"""
        s3, mid, ex = (1, 2, (2, 4)) if self.OS == 8 else (2, 1, (1, 2))
        x = self._conv(x, "entry_flow_conv1_1", 3, 32, stride=2, padding="same")
        x = self._bn_act(x, "entry_flow_conv1_1_BN", 1e-3, "relu")
        x = self._conv(x, "entry_flow_conv1_2", 3, 64, stride=1, padding="same")
        x = self._bn_act(x, "entry_flow_conv1_2_BN", 1e-3, "relu")
        x, _ = self._xception_block(x, [128, 128, 128], "entry_flow_block1", "conv", 2)
        x, skip1 = self._xception_block(x, [256, 256, 256], "entry_flow_block2", "conv", 2)
        x, _ = self._xception_block(x, [728, 728, 728], "entry_flow_block3", "conv", s3)
        for i in range(16):
            x, _ = self._xception_block(x, [728, 728, 728], "middle_flow_unit_%d" % (i + 1), "sum", 1, rate=mid)
        x, _ = self._xception_block(x, [728, 1024, 1024], "exit_flow_block1", "conv", 1, rate=ex[0])
        x, _ = self._xception_block(x, [1536, 1536, 2048], "exit_flow_block2", "none", 1, rate=ex[1], depth_activation=True)
        return x, skip1

    def _deeplab_head(self, x, skip1=None, input_hw=None):
        """model.py:457-500: image pooling + 1x1 ASPP branches, concat_projection, Dropout(0.1), Conv2D(classes, 1x1,
        activation), BilinearUpsampling(align_corners=True) to the input size.  Returns the network OUTPUT (probabilities)."""
        EPS = 1e-5
        h, w = x.shape[2], x.shape[3]
        b4 = L.rb(x.mean(dim=(2, 3), keepdim=True), self.storage)        # AveragePooling2D over the whole map
        b4 = self._conv(b4, "image_pooling", 1, 256)
        b4 = self._bn_act(b4, "image_pooling_BN", EPS, "relu")
        b4 = L.resize_bilinear_tf1(b4, h, w, align_corners=True)         # from 1x1: a broadcast
        b0 = self._conv(x, "aspp0", 1, 256)
        b0 = self._bn_act(b0, "aspp0_BN", EPS, "relu")
        if skip1 is not None:   # xception: three atrous separable branches as well (model.py:472-481)
            rates = (12, 24, 36) if self.OS == 8 else (6, 12, 18)
            bs = [self._sepconv_bn(x, 256, "aspp%d" % (k + 1), rate=r, depth_activation=True, eps=EPS) for k, r in enumerate(rates)]
            y = torch.cat([b4, b0] + bs, dim=1)
        else:
            y = torch.cat([b4, b0], dim=1)
        y = self._conv(y, "concat_projection", 1, 256)
        y = self._bn_act(y, "concat_projection_BN", EPS, "relu")
        if self.training and self.dropout is not None and self.dropout[0] > 0:
            from .philox import dropout_keep_mask
            rate, seed, salt, step = self.dropout
            n, c = y.shape[0], y.shape[1]
            keep = dropout_keep_mask(n * h * w, c, rate, seed, salt, step).reshape(n, h, w, c)
            keep = torch.from_numpy(keep).permute(0, 3, 1, 2).to(y.dtype)
            y = L.rb(y * keep * float(np.float32(1.0) / (np.float32(1.0) - np.float32(rate))), self.storage)
        self.taps["concat_projection_relu"] = y
        if skip1 is not None:   # DeepLab v3+ decoder (model.py:488-500)
            H, W = input_hw
            y = L.rb(L.resize_bilinear_tf1(y, -(-H // 4), -(-W // 4), align_corners=True), self.storage)
            d = self._conv(skip1, "feature_projection0", 1, 48)
            d = self._bn_act(d, "feature_projection0_BN", EPS, "relu")
            y = torch.cat([y, d], dim=1)
            y = self._sepconv_bn(y, 256, "decoder_conv0", depth_activation=True, eps=EPS)
            y = self._sepconv_bn(y, 256, "decoder_conv1", depth_activation=True, eps=EPS)
            self.taps["decoder_conv1"] = y
        name = "logits_semantic" if self.classes == 21 else "custom_logits_semantic"
        wk, b = self.P.conv(name, 1, 1, y.shape[1], self.classes, True)
        return L.conv2d(y, wk, b, 1, "same", self.storage)

    # -- full graph ------------------------------------------------------------------------
    def forward(self, x_nhwc: torch.Tensor, emit_logits: bool = False) -> torch.Tensor:
        """x: float NHWC raw 0..255 (no preprocessing call anywhere in the reference).

        emit_logits=True strips the trailing Activation (what musket compile does for lovasz_loss)."""
        self.P._in_encoder = True
        x = x_nhwc.permute(0, 3, 1, 2).contiguous().to(getattr(self, "dtype", torch.float32))
        if self.arch == "DeepLabV3":
            H, W = x.shape[2], x.shape[3]
            if self.backbone == "xception":
                x, skip1 = self._xception(x)
                self.P._in_encoder = False
                z = self._deeplab_head(x, skip1, (H, W))
            else:
                x = self._mobilenetv2(x)
                self.P._in_encoder = False
                z = self._deeplab_head(x)
            self.taps["logits_small"] = z
            if emit_logits or self.activation in (None, "none", "linear"):
                p = z
            elif self.activation == "sigmoid":
                p = torch.sigmoid(z)
            else:
                p = torch.softmax(z, dim=1)
            # the activation sits INSIDE the 1x1 Conv2D; the upsampling acts on its output (model.py:499-500)
            return L.resize_bilinear_tf1(p, H, W, align_corners=True).permute(0, 2, 3, 1)
        if self.backbone == "vgg16":
            x, skips = self._vgg16(x)
        else:
            x, skips = self._resnet(x)
        self.P._in_encoder = False
        if self.arch == "Unet":
            x = self._unet_decoder(x, skips)
            w, b = self.P.conv("final_conv", 3, 3, x.shape[1], self.classes, True)
            logits = L.conv2d(x, w, b, 1, "same", self.storage)
        elif self.arch == "FPN":
            x = self._fpn_decoder(x, skips)
            w, b = self.P.conv("head_conv", 3, 3, x.shape[1], self.classes, True)
            logits = L.conv2d(x, w, b, 1, "same", self.storage)
            logits = L.resize_bilinear_tf1(logits, logits.shape[2] * 4, logits.shape[3] * 4)
        elif self.arch == "Linknet":
            x = self._linknet_decoder(x, skips)
            w, b = self.P.conv("final_conv", 3, 3, x.shape[1], self.classes, True)
            logits = L.conv2d(x, w, b, 1, "same", self.storage)
        elif self.arch == "PSPNet":
            x = self._pspnet_decoder(x)
            w, b = self.P.conv("final_conv", 3, 3, x.shape[1], self.classes, True)
            logits = L.conv2d(x, w, b, 1, "same", self.storage)
            logits = L.resize_bilinear_tf1(logits, logits.shape[2] * self.psp_factor, logits.shape[3] * self.psp_factor)
        else:
            raise ValueError("Unknown architecture")
        self.taps["logits"] = logits
        if emit_logits or self.activation in (None, "none", "linear"):
            y = logits
        elif self.activation == "sigmoid":
            y = torch.sigmoid(logits)
        elif self.activation == "softmax":
            y = torch.softmax(logits, dim=1)
        else:
            raise ValueError("unknown activation " + str(self.activation))
        return y.permute(0, 2, 3, 1)

    __call__ = forward

    def state_numpy(self) -> Dict[str, np.ndarray]:
        d = {k: v.detach().numpy().copy() for k, v in self.P.params.items()}
        d.update({k: v.numpy().copy() for k, v in self.P.buffers.items()})
        return d

    def load_numpy(self, d: Dict[str, np.ndarray]):
        with torch.no_grad():
            for k, v in self.P.params.items():
                v.copy_(torch.from_numpy(np.asarray(d[k])))
            for k, v in self.P.buffers.items():
                if k in d:
                    v.copy_(torch.from_numpy(np.asarray(d[k])))
