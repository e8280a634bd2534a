"""Model graph builders on the static-graph engine: what the reference builds at segmentation.py:96-155
(createNet1 -> segmentation_models.Unet(backbone_name=..., ...)) [DEP segmentation_models==0.2.1,
classification_models].  Layer / parameter names are the Keras names (weights exchangeable by name).

Layout decisions: every skip tensor is written by its producer straight into the channel slice
of the decoder's concat buffer, and every decoder stage output is written 2x-upsampled into the next
concat buffer by its BN-apply kernel, so UpSampling2D + Concatenate cost no kernel and no extra traffic.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import engine as E

ENC_BN_EPS = 2e-5   # classification_models get_bn_params()
DEC_BN_EPS = 1e-3   # keras BatchNormalization default
RESNET_REPS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3),
               "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
RESNET_BOTTLENECK = {"resnet18": False, "resnet34": False, "resnet50": True, "resnet101": True, "resnet152": True}
VGG16_BLOCKS = ((64, 2), (128, 2), (256, 3), (512, 3), (512, 3))  # keras.applications.VGG16 [DEP]
KNOWN_BACKBONES = sorted(RESNET_REPS) + ["vgg16"]
KNOWN_ARCHITECTURES = ["Unet", "FPN", "Linknet", "PSPNet", "DeepLabV3"]
DEEPLAB_BACKBONES = ["mobilenetv2", "xception"]   # impl/deeplab/model.py:324-326
# MobileNetV2 feature extractor of the reference's DeepLabV3+ (impl/deeplab/model.py:386-433): (filters, stride, expansion,
# block_id, skip_connection, atrous rate); strides after block 3 are replaced by rates (output stride 8)
MOBILENETV2_BLOCKS = (
    (16, 1, 1, 0, False, 1),
    (24, 2, 6, 1, False, 1), (24, 1, 6, 2, True, 1),
    (32, 2, 6, 3, False, 1), (32, 1, 6, 4, True, 1), (32, 1, 6, 5, True, 1),
    (64, 1, 6, 6, False, 1), (64, 1, 6, 7, True, 2), (64, 1, 6, 8, True, 2), (64, 1, 6, 9, True, 2),
    (96, 1, 6, 10, False, 2), (96, 1, 6, 11, True, 2), (96, 1, 6, 12, True, 2),
    (160, 1, 6, 13, False, 2), (160, 1, 6, 14, True, 4), (160, 1, 6, 15, True, 4),
    (320, 1, 6, 16, False, 4),
)


def _make_divisible(v, divisor=8, min_value=None):
    """impl/deeplab/model.py:225-233"""
    min_value = divisor if min_value is None else min_value
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


class SegNet(E.Net):
    """U-Net over a qubvel pre-activation ResNet; input uint8 NHWC image, target uint8 NHWC mask."""

    def __init__(self, backbone="resnet34", classes=1, input_shape=(512, 512, 3), batch=16,
                 decoder_filters=(256, 128, 64, 32, 16), device="cuda:0", seed=0,
                 enc_init="he_uniform", dec_init="glorot_uniform", loss=(1.0, 0.0, 0.0), architecture="Unet",
                 decoder_block_type="upsampling", pyramid_block_filters=256, segmentation_block_filters=128,
                 dropout=None, precision="bf16", decoder_use_batchnorm=True, downsample_factor=8, psp_conv_filters=512,
                 activation="sigmoid", OS=16):
        super().__init__(batch, device, seed, precision)
        self.dec_bn = bool(decoder_use_batchnorm)
        backbone = backbone.lower()
        if architecture == "DeepLabV3":
            self._build_deeplabv3(backbone, classes, input_shape, batch, activation, loss,
                                  0.1 if dropout is None else float(dropout), int(OS))
            return
        if architecture not in KNOWN_ARCHITECTURES:
            print("Unknown architecture:" + str(architecture))
            print("Known architectures:", KNOWN_ARCHITECTURES)
            raise ValueError("Unknown architecture")
        self.architecture = architecture
        linknet = architecture == "Linknet"
        fpn = architecture == "FPN"
        psp = architecture == "PSPNet"
        if psp and backbone.lower() == "vgg16":
            raise NotImplementedError("PSPNet is built over the ResNet encoders only")
        if psp and downsample_factor not in (4, 8, 16):
            raise ValueError("PSPNet: downsample_factor must be 4, 8 or 16")
        if fpn and dropout:
            raise NotImplementedError("FPN dropout (SpatialDropout2D) is not built; the schema default is None")
        if fpn and backbone == "vgg16":
            raise NotImplementedError("FPN is built over the ResNet encoders only")
        if decoder_block_type not in ("upsampling", "transpose"):
            raise ValueError("decoder_block_type must be 'upsampling' or 'transpose'")
        if not self.dec_bn and (linknet or fpn or psp or decoder_block_type == "transpose"):
            raise NotImplementedError("decoder_use_batchnorm: false is built for the Unet upsampling decoder only")
        transpose = decoder_block_type == "transpose" and not linknet   # schema segmentation.raml:162-165 (Unet only)
        if transpose and backbone == "vgg16":
            raise NotImplementedError("decoder_block_type: transpose is built over the ResNet encoders only")
        if backbone == "vgg16":
            if linknet:
                raise NotImplementedError("Linknet is built over the ResNet encoders only")
            self._build_vgg16_unet(classes, input_shape, batch, decoder_filters, dec_init, loss)
            return
        if backbone not in RESNET_REPS:
            print("Unknown backbone:" + backbone)
            print("Known backbones:", KNOWN_BACKBONES)
            raise ValueError("Unknown backbone")
        H, W, CI = input_shape
        if psp:
            if H % (6 * downsample_factor) or W % (6 * downsample_factor):
                # segmentation_models' own check: the pyramid levels (1, 2, 3, 6) must tile the feature map exactly
                raise ValueError("PSPNet: input height/width must be divisible by 6 * downsample_factor = %d" % (6 * downsample_factor))
        elif H % 32 or W % 32:
            raise ValueError("input height/width must be divisible by 32")
        if not 1 <= CI <= 4:
            raise NotImplementedError("input channels: 1..4 are built (uint8 augmentation / stem kernels); shape[2] = %d" % CI)
        if len(decoder_filters) != 5:
            raise ValueError("decoder_filters must have 5 entries")
        N = batch
        reps, bott = RESNET_REPS[backbone], RESNET_BOTTLENECK[backbone]
        exp = 4 if bott else 1
        df = list(decoder_filters)
        self.input_shape, self.classes, self.backbone = (H, W, CI), classes, backbone

        self.img = E.Buf(self, N, H, W, CI, E.U8, name="image")
        self.mask = E.Buf(self, N, H, W, classes, E.U8, name="mask")

        # decoder concat buffers [up | skip]; skip channel counts per stage (0.2.1 skip names)
        skip_c = [256 * exp, 128 * exp, 64 * exp, 64, 0]
        up_c = df[:] if transpose else [512 * exp] + df[:4]   # transpose blocks concat [ConvT output (f_i) | skip]
        cat: List[E.Buf] = []
        for i in range(0 if (linknet or fpn or psp) else 5):
            s = 32 >> i  # input of stage i is at H/32 * 2^i after upsampling -> H / (16 >> i) ... computed below
            hh, ww = H // (16 >> i) if i < 4 else H, W // (16 >> i) if i < 4 else W
            cat.append(E.Buf(self, N, hh, ww, up_c[i] + skip_c[i], name="cat%d" % i))
        skip_view = {  # keras layer name -> (stage index)
            "stage4_unit1_relu1": 0, "stage3_unit1_relu1": 1, "stage2_unit1_relu1": 2, "relu0": 3}
        skip_names = list(skip_view)
        if linknet or fpn or psp:  # Linknet / FPN add their skips instead of concatenating them: plain buffers, no concat layout
            skip_view, cat = {}, []
        psp_tap, psp_cat = None, None
        if psp:   # the feature map the pyramid pools is written straight into channel slice 0 of the PSP concat buffer
            psp_tap = {4: "stage2_unit1_relu1", 8: "stage3_unit1_relu1", 16: "stage4_unit1_relu1"}[downsample_factor]

        def skip_buf(name, n, h, w, c):
            if name in skip_view:
                i = skip_view[name]
                assert cat[i].h == h and cat[i].w == w and skip_c[i] == c, (name, cat[i].h, h, skip_c[i], c)
                return cat[i].slice(up_c[i], c, name=name)
            return E.Buf(self, n, h, w, c, name=name)

        # ---- encoder ---------------------------------------------------------------------------
        # bn_data output in space-to-depth layout [N, H/2, W/2, 4 sub-pixels x 8 channels] (see engine.StemConv)
        z = E.Buf(self, N, H // 2, W // 2, 64, name="conv0")
        if precision == "fp32":   # parity mode: the plain 7x7/2 convolution over the 8-channel padded bn_data output
            x0 = E.Buf(self, N, H, W, 8, name="bn_data")
            inorm = E.InputNorm(self, self.img, x0, "bn_data", ENC_BN_EPS)
            E.Conv(self, x0, z, "conv0", 7, stride=2, pad=3, init=enc_init, needs_dgrad=False, cin_real=CI, stem_beta=inorm.beta)
        else:
            x0 = E.Buf(self, N, H // 2, W // 2, 32, name="bn_data_s2d")
            inorm = E.InputNorm(self, self.img, x0, "bn_data", ENC_BN_EPS)
            E.StemConv(self, x0, z, "conv0", cin_real=CI, stem_beta=inorm.beta, init=enc_init)
        relu0 = skip_buf("relu0", N, H // 2, W // 2, 64)
        E.BNRelu(self, z, relu0, "bn0", ENC_BN_EPS)
        x = E.Buf(self, N, H // 4, W // 4, 64, name="pooling0")
        E.MaxPool(self, relu0, x, 3, 2, 1)
        h, w = H // 4, W // 4
        for stage, rep in enumerate(reps):
            f = 64 * 2 ** stage
            for block in range(rep):
                pre = "stage%d_unit%d_" % (stage + 1, block + 1)
                first = block == 0
                stride = 2 if (first and stage > 0) else 1
                ho, wo = h // stride, w // stride
                if psp and pre + "relu1" == psp_tap:
                    psp_cat = E.Buf(self, N, h, w, x.c + 4 * psp_conv_filters, name="psp_concat")
                    feat = psp_cat.slice(0, x.c, name=psp_tap)
                    E.BNRelu(self, x, feat, pre + "bn1", ENC_BN_EPS)
                    self.encoder_param_names = list(self.params.keys())
                    self._build_pspnet_decoder(feat, psp_cat, psp_conv_filters, downsample_factor, classes, dec_init, loss)
                    return   # the Keras Model ends the encoder here: later encoder layers are not part of the graph
                y = skip_buf(pre + "relu1", N, h, w, x.c)
                out = E.Buf(self, N, ho, wo, f * exp, name=pre + "add")
                if first:
                    E.BNRelu(self, x, y, pre + "bn1", ENC_BN_EPS)
                    sc = E.Buf(self, N, ho, wo, f * exp, name=pre + "sc")
                    E.Conv(self, y, sc, pre + "sc", 1, stride=stride, pad=0, init=enc_init)
                    sc.set_grad(out.grad())  # d(shortcut) == d(sum)
                    res = sc
                else:
                    E.BNRelu(self, x, y, pre + "bn1", ENC_BN_EPS, extra_grad=(lambda o=out: o.grad()))
                    res = x
                if bott:
                    z1 = E.Buf(self, N, h, w, f, name=pre + "conv1")
                    E.Conv(self, y, z1, pre + "conv1", 1, init=enc_init)
                    a2 = E.Buf(self, N, h, w, f, name=pre + "relu2")
                    E.BNRelu(self, z1, a2, pre + "bn2", ENC_BN_EPS)
                    z2 = E.Buf(self, N, ho, wo, f, name=pre + "conv2")
                    E.Conv(self, a2, z2, pre + "conv2", 3, stride=stride, pad=1, init=enc_init)
                    a3 = E.Buf(self, N, ho, wo, f, name=pre + "relu3")
                    E.BNRelu(self, z2, a3, pre + "bn3", ENC_BN_EPS)
                    E.Conv(self, a3, out, pre + "conv3", 1, residual=res, init=enc_init)
                else:
                    z1 = E.Buf(self, N, ho, wo, f, name=pre + "conv1")
                    E.Conv(self, y, z1, pre + "conv1", 3, stride=stride, pad=1, init=enc_init)
                    a2 = E.Buf(self, N, ho, wo, f, name=pre + "relu2")
                    E.BNRelu(self, z1, a2, pre + "bn2", ENC_BN_EPS)
                    E.Conv(self, a2, out, pre + "conv2", 3, pad=1, residual=res, init=enc_init)
                x, h, w = out, ho, wo
        self.encoder_param_names = list(self.params.keys())
        if linknet or transpose or fpn:
            top = E.Buf(self, N, h, w, x.c, name="relu1")
            E.BNRelu(self, x, top, "bn1", ENC_BN_EPS)
            self.encoder_param_names = list(self.params.keys())
            if fpn:
                self._build_fpn_decoder(top, [self.bufs[nm] for nm in skip_names[:3]], pyramid_block_filters,
                                        segmentation_block_filters, classes, dec_init, loss)
            elif linknet:
                self._build_linknet_decoder(top, [self.bufs[nm] for nm in skip_names], df, classes, dec_init, loss)
            else:
                self._build_transpose_decoder(top, cat, df, classes, dec_init, loss)
            return
        # bn1/relu1 written 2x-upsampled straight into the first concat buffer
        E.BNRelu(self, x, cat[0].slice(0, up_c[0], name="relu1_up"), "bn1", ENC_BN_EPS, up=2)
        self.encoder_param_names = list(self.params.keys())

        self._build_decoder(cat, df, classes, dec_init, loss)

    def _build_transpose_decoder(self, x, cat, df, classes, dec_init, loss):
        """Unet `decoder_block_type: transpose` (segmentation.raml:162-165) [DEP segmentation_models 0.2.1]:
        Conv2DTranspose(f, 4x4, strides 2, 'same') + BN + ReLU -> Concatenate(skip) -> 3x3 conv + BN + ReLU.  The
        transposed conv runs as a convolution of the zero-inserted input with the flipped kernel (engine.Conv up=2)."""
        N = self.batch
        for i, f in enumerate(df):
            pre = "decoder_stage%d_" % i
            z = E.Buf(self, N, 2 * x.h, 2 * x.w, f, name=pre + "transpose")
            E.Conv(self, x, z, pre + "transpose", 4, pad=2, up=2, transposed=True, init=dec_init)
            E.BNRelu(self, z, cat[i].slice(0, f, name=pre + "relu1"), pre + "bn1", DEC_BN_EPS)
            z2 = E.Buf(self, N, z.h, z.w, f, name=pre + "conv2")
            E.Conv(self, cat[i], z2, pre + "conv2", 3, pad=1, init=dec_init)
            a2 = E.Buf(self, N, z.h, z.w, f, name=pre + "relu2")
            E.BNRelu(self, z2, a2, pre + "bn2", DEC_BN_EPS)
            x = a2
        self.head = E.Head(self, x, classes, "final_conv", init=dec_init)
        self.loss = E.Loss(self, self.head, self.mask, *loss)
        self.finalize()

    def _build_fpn_decoder(self, top, skips, pyr, segf, classes, dec_init, loss):
        """segmentation_models 0.2.1 FPN [DEP] (reference segmentation.py:109-113; schema segmentation.raml:179-204; SURVEY.md
        8 a-5).  Top-down pyramid over relu1 (H/32) and the three deepest skips: 1x1 lateral conv (pyramid_block_filters,
        bias) + Add(UpSampling2D(2)(previous level)) -- the Add is the residual input of the lateral conv's epilogue; per
        level 2x [3x3 conv (segmentation_block_filters) + BN + ReLU], bilinear x8/x4/x2/x1 written straight into the channel
        slices of the concat buffer at H/4; 3x3 conv (4*segmentation_block_filters) + BN + ReLU; head conv 3x3 (classes,
        bias) and the x4 bilinear `last_upsample` of the logits (engine.UpHead)."""
        N = self.batch
        feats = [top] + list(skips)
        p = None
        pyramid = []
        for i, c in enumerate(feats):
            name = "pyramid_stage_%d_conv1x1" % i
            lat = E.Buf(self, N, c.h, c.w, pyr, name="pyramid_stage_%d" % i)
            if p is None:
                E.Conv(self, c, lat, name, 1, bias=True, init=dec_init)
            else:
                up = E.Buf(self, N, c.h, c.w, pyr, name="pyramid_stage_%d_up" % i)
                E.Upsample2x(self, p, up)
                E.Conv(self, c, lat, name, 1, bias=True, residual=up, init=dec_init)
                up.set_grad(lat.grad())  # d(Add)/d(upsampled) == d(lateral output)
            p = lat
            pyramid.append(p)
        hq, wq = pyramid[-1].h, pyramid[-1].w
        cat = E.Buf(self, N, hq, wq, 4 * segf, name="fpn_concat")
        for i, p in enumerate(pyramid):
            pre = "segm_stage_%d_" % i
            z1 = E.Buf(self, N, p.h, p.w, segf, name=pre + "conv1")
            E.Conv(self, p, z1, pre + "conv1", 3, pad=1, init=dec_init)
            a1 = E.Buf(self, N, p.h, p.w, segf, name=pre + "relu1")
            E.BNRelu(self, z1, a1, pre + "bn1", DEC_BN_EPS)
            z2 = E.Buf(self, N, p.h, p.w, segf, name=pre + "conv2")
            E.Conv(self, a1, z2, pre + "conv2", 3, pad=1, init=dec_init)
            dst = cat.slice(i * segf, segf, name=pre + "out")
            if p.h == hq:
                E.BNRelu(self, z2, dst, pre + "bn2", DEC_BN_EPS)
            else:
                a2 = E.Buf(self, N, p.h, p.w, segf, name=pre + "relu2")
                E.BNRelu(self, z2, a2, pre + "bn2", DEC_BN_EPS)
                E.Resize(self, a2, dst)
        zf = E.Buf(self, N, hq, wq, 4 * segf, name="final_stage_conv")
        E.Conv(self, cat, zf, "final_stage_conv", 3, pad=1, init=dec_init)
        af = E.Buf(self, N, hq, wq, 4 * segf, name="final_stage_relu")
        E.BNRelu(self, zf, af, "final_stage_bn", DEC_BN_EPS)
        self.head = E.UpHead(self, af, classes, "head_conv", up=4, init=dec_init)
        self.loss = E.Loss(self, self.head, self.mask, *loss)
        self.finalize()

    def _build_pspnet_decoder(self, feat, cat, filters, factor, classes, dec_init, loss):
        """segmentation_models 0.2.1 PSPNet [DEP, recalled] (reference segmentation.py:109-113; schema segmentation.raml:226-248):
        for level in (1, 2, 3, 6): AveragePooling2D(size / level) -> 1x1 conv (psp_conv_filters) + BN + ReLU -> bilinear resize
        (TF1 legacy) back to the feature size, written into its channel slice of the concat buffer whose slice 0 IS the feature
        map; 1x1 conv (512) + BN + ReLU; final_conv 3x3 (classes, bias) and the bilinear x downsample_factor upsample of the
        logits (engine.UpHead)."""
        N, h, w = self.batch, feat.h, feat.w
        for li, level in enumerate((1, 2, 3, 6)):
            pre = "psp_level%d_" % level
            pooled = E.Buf(self, N, level, level, feat.c, name=pre + "pool")
            E.AvgPool(self, feat, pooled, h // level)
            z = E.Buf(self, N, level, level, filters, name=pre + "conv")
            E.Conv(self, pooled, z, pre + "conv", 1, init=dec_init)
            a = E.Buf(self, N, level, level, filters, name=pre + "relu")
            E.BNRelu(self, z, a, pre + "bn", DEC_BN_EPS)
            E.Resize(self, a, cat.slice(feat.c + li * filters, filters, name=pre + "up"))
        z = E.Buf(self, N, h, w, 512, name="psp_conv")
        E.Conv(self, cat, z, "psp_conv", 1, init=dec_init)
        a = E.Buf(self, N, h, w, 512, name="psp_relu")
        E.BNRelu(self, z, a, "psp_bn", DEC_BN_EPS)
        self.head = E.UpHead(self, a, classes, "final_conv", up=factor, init=dec_init)
        self.loss = E.Loss(self, self.head, self.mask, *loss)
        self.finalize()

    def _build_linknet_decoder(self, x, skips, df, classes, dec_init, loss):
        """segmentation_models 0.2.1 Linknet decoder [DEP] (schema segmentation.raml:205-225): per stage
        1x1 conv (C/4) + BN + ReLU -> UpSampling2D(2) -> 3x3 conv (C/4) + BN + ReLU -> 1x1 conv (C_skip | 16) + BN + ReLU
        -> Add(skip); the nearest upsample is the `up=2` write mode of the first BatchNorm-apply."""
        N = self.batch
        for i in range(5):
            pre = "decoder_stage%d_" % i
            skip = skips[i] if i < len(skips) else None
            cin = x.c
            cout = skip.c if skip is not None else (df[i] if df[i] else 16)
            z1 = E.Buf(self, N, x.h, x.w, cin // 4, name=pre + "conv1")
            E.Conv(self, x, z1, pre + "conv1", 1, init=dec_init)
            a1 = E.Buf(self, N, 2 * x.h, 2 * x.w, cin // 4, name=pre + "relu1_up")
            E.BNRelu(self, z1, a1, pre + "bn1", DEC_BN_EPS, up=2)
            z2 = E.Buf(self, N, a1.h, a1.w, cin // 4, name=pre + "conv2")
            E.Conv(self, a1, z2, pre + "conv2", 3, pad=1, init=dec_init)
            a2 = E.Buf(self, N, a1.h, a1.w, cin // 4, name=pre + "relu2")
            E.BNRelu(self, z2, a2, pre + "bn2", DEC_BN_EPS)
            z3 = E.Buf(self, N, a1.h, a1.w, cout, name=pre + "conv3")
            E.Conv(self, a2, z3, pre + "conv3", 1, init=dec_init)
            a3 = E.Buf(self, N, a1.h, a1.w, cout, name=pre + "relu3")
            E.BNRelu(self, z3, a3, pre + "bn3", DEC_BN_EPS)
            if skip is not None:
                out = E.Buf(self, N, a1.h, a1.w, cout, name=pre + "add")
                E.Add(self, a3, skip, out)
                x = out
            else:
                x = a3
        self.head = E.Head(self, x, classes, "final_conv", init=dec_init)
        self.loss = E.Loss(self, self.head, self.mask, *loss)
        self.finalize()

    def _build_decoder(self, cat, df, classes, dec_init, loss):
        N = self.batch
        # ---- decoder -----------------------------------------------------------------------------
        for i, f in enumerate(df):
            pre = "decoder_stage%d_" % i
            cin = cat[i]
            if not self.dec_bn:
                # `use_batchnorm: false` (schema segmentation.raml:173-176 -> decoder_use_batchnorm): Conv2D with bias + ReLU,
                # no BatchNormalization; the post-ReLU output is copied 2x-upsampled into the next concat buffer
                a1 = E.Buf(self, N, cin.h, cin.w, f, name=pre + "relu1")
                E.Conv(self, cin, a1, pre + "conv1", 3, pad=1, bias=True, relu=True, init=dec_init)
                a2 = E.Buf(self, N, cin.h, cin.w, f, name=pre + "relu2")
                E.Conv(self, a1, a2, pre + "conv2", 3, pad=1, bias=True, relu=True, init=dec_init)
                if i < 4:
                    E.UpCopy(self, a2, cat[i + 1].slice(0, f, name=pre + "relu2_up"))
                else:
                    last = a2
                continue
            z1 = E.Buf(self, N, cin.h, cin.w, f, name=pre + "conv1")
            E.Conv(self, cin, z1, pre + "conv1", 3, pad=1, init=dec_init)
            a1 = E.Buf(self, N, cin.h, cin.w, f, name=pre + "relu1")
            E.BNRelu(self, z1, a1, pre + "bn1", DEC_BN_EPS)
            z2 = E.Buf(self, N, cin.h, cin.w, f, name=pre + "conv2")
            E.Conv(self, a1, z2, pre + "conv2", 3, pad=1, init=dec_init)
            if i < 4:
                E.BNRelu(self, z2, cat[i + 1].slice(0, f, name=pre + "relu2_up"), pre + "bn2", DEC_BN_EPS, up=2)
            else:
                last = E.Buf(self, N, cin.h, cin.w, f, name=pre + "relu2")
                E.BNRelu(self, z2, last, pre + "bn2", DEC_BN_EPS)
        self.head = E.Head(self, last, classes, "final_conv", init=dec_init)
        self.loss = E.Loss(self, self.head, self.mask, *loss)
        self.finalize()

    def _build_deeplabv3(self, backbone, classes, input_shape, batch, activation, loss, dropout, OS=16):
        """DeepLabV3+ over MobileNetV2, the model the reference ships in-tree and registers as architecture `DeepLabV3`
        (impl/deeplab/model.py:278-505, segmentation.py:31-33; all five example configs use it): Conv 3x3/2 + BN + ReLU6, 17
        inverted-residual blocks (1x1 expand -> 3x3 depthwise (stride / atrous rate) -> 1x1 project, BatchNorm eps 1e-3
        momentum 0.999, ReLU6, identity Add) at output stride 8, ASPP with the image-pooling branch and the 1x1 branch only
        (:462-481), concat_projection + BN + ReLU + Dropout(0.1), then Conv2D(classes, 1x1, activation) at 1/8 resolution and
        the align_corners bilinear resize of the PROBABILITIES to the input size (:494-500).  Raw 0..255 input like every model
        of the pipeline.  Layer names are the Keras names of that file (weights exchangeable by name)."""
        if backbone not in DEEPLAB_BACKBONES:
            print("Unknown backbone:" + backbone)
            print("Known backbones:", DEEPLAB_BACKBONES)
            raise ValueError("Unknown backbone")
        if backbone == "xception":
            return self._build_deeplabv3_xception(classes, input_shape, batch, activation, loss, dropout, OS)
        H, W, CI = input_shape
        if H % 8 or W % 8:
            raise ValueError("DeepLabV3: input height/width must be divisible by 8 (output stride of the feature extractor)")
        if not 1 <= CI <= 4:
            raise NotImplementedError("input channels: 1..4 are built (uint8 augmentation / stem kernels); shape[2] = %d" % CI)
        N = batch
        self.architecture, self.input_shape, self.classes, self.backbone = "DeepLabV3", (H, W, CI), classes, backbone
        self.img = E.Buf(self, N, H, W, CI, E.U8, name="image")
        self.mask = E.Buf(self, N, H, W, classes, E.U8, name="mask")
        BN_EPS, BN_MOM, RELU6, init = 1e-3, 0.999, 2, "glorot_uniform"
        x0 = E.Buf(self, N, H, W, 8, name="input_bf16")
        E.InputCast(self, self.img, x0)
        h, w = H // 2, W // 2
        z = E.Buf(self, N, h, w, _make_divisible(32), name="Conv")
        E.Conv(self, x0, z, "Conv", 3, stride=2, pad=0, init=init, needs_dgrad=False, cin_real=CI)   # TF 'same', even size: pad 0 / 1
        x = E.Buf(self, N, h, w, z.c, name="Conv_Relu6")
        E.BNRelu(self, z, x, "Conv_BN", BN_EPS, relu=RELU6, momentum=BN_MOM)
        for filters, stride, expansion, bid, skip, rate in MOBILENETV2_BLOCKS:
            pre = "expanded_conv_%d_" % bid if bid else "expanded_conv_"
            t = x
            if bid:
                e = E.Buf(self, N, h, w, expansion * x.c, name=pre + "expand")
                E.Conv(self, x, e, pre + "expand", 1, init=init)
                t = E.Buf(self, N, h, w, e.c, name=pre + "expand_relu")
                E.BNRelu(self, e, t, pre + "expand_BN", BN_EPS, relu=RELU6, momentum=BN_MOM)
            ho, wo = -(-h // stride), -(-w // stride)
            d = E.Buf(self, N, ho, wo, t.c, name=pre + "depthwise")
            E.DWConv(self, t, d, pre + "depthwise", 3, stride=stride, dilation=rate, init=init)
            dr = E.Buf(self, N, ho, wo, t.c, name=pre + "depthwise_relu")
            E.BNRelu(self, d, dr, pre + "depthwise_BN", BN_EPS, relu=RELU6, momentum=BN_MOM)
            pj = E.Buf(self, N, ho, wo, _make_divisible(filters), name=pre + "project")
            E.Conv(self, dr, pj, pre + "project", 1, init=init)
            pb = E.Buf(self, N, ho, wo, pj.c, name=pre + "project_BN")
            E.BNRelu(self, pj, pb, pre + "project_BN", BN_EPS, relu=False, momentum=BN_MOM)
            if skip:
                out = E.Buf(self, N, ho, wo, pj.c, name=pre + "add")
                E.Add(self, pb, x, out)
                x = out
            else:
                x = pb
            h, w = ho, wo
        self.encoder_param_names = list(self.params.keys())   # "# end of feature extractor" (model.py:455)
        # ---- ASPP: image pooling branch (whole-map mean -> 1x1 -> BN -> ReLU -> broadcast) | 1x1 branch --------------------
        ASPP_EPS = 1e-5
        cat = E.Buf(self, N, h, w, 512, name="aspp_concat")
        gp = E.Buf(self, N, 1, 1, x.c, name="image_pooling_avg")
        E.GlobalAvgPool(self, x, gp)
        ip = E.Buf(self, N, 1, 1, 256, name="image_pooling")
        E.Conv(self, gp, ip, "image_pooling", 1, init=init)
        ipr = E.Buf(self, N, 1, 1, 256, name="image_pooling_relu")
        E.BNRelu(self, ip, ipr, "image_pooling_BN", ASPP_EPS)
        E.Broadcast(self, ipr, cat.slice(0, 256, name="image_pooling_up"))
        a0 = E.Buf(self, N, h, w, 256, name="aspp0")
        E.Conv(self, x, a0, "aspp0", 1, init=init)
        E.BNRelu(self, a0, cat.slice(256, 256, name="aspp0_activation"), "aspp0_BN", ASPP_EPS)
        cp = E.Buf(self, N, h, w, 256, name="concat_projection")
        E.Conv(self, cat, cp, "concat_projection", 1, init=init)
        cpr = E.Buf(self, N, h, w, 256, name="concat_projection_relu")
        E.BNRelu(self, cp, cpr, "concat_projection_BN", ASPP_EPS)
        self.dropout = E.Dropout(self, cpr, dropout, salt=0xD0)
        self.head = E.ProbHead(self, cpr, classes, (H, W), activation or "none",
                               "logits_semantic" if classes == 21 else "custom_logits_semantic", init=init)
        self.loss = E.Loss(self, self.head, self.mask, *loss)
        self.finalize()

    def _sepconv_bn(self, x, filters, prefix, stride=1, rate=1, depth_activation=False, eps=1e-3, out=None):
        """SepConv_BN of impl/deeplab/model.py:110-147: [ReLU] -> depthwise 3x3 (stride > 1: explicit padding + 'valid') -> BN
        [-> ReLU] -> 1x1 -> BN [-> ReLU]; `out` = destination of the last BatchNorm (e.g. a concat slice)."""
        N = self.batch
        t = x
        if not depth_activation:
            t = E.Buf(self, N, x.h, x.w, x.c, name=prefix + "_relu")
            E.Relu(self, x, t)
        ho, wo = -(-x.h // stride), -(-x.w // stride)
        d = E.Buf(self, N, ho, wo, x.c, name=prefix + "_depthwise")
        E.DWConv(self, t, d, prefix + "_depthwise", 3, stride=stride, dilation=rate, pad=None if stride == 1 else rate)
        db = E.Buf(self, N, ho, wo, x.c, name=prefix + "_depthwise_BN")
        E.BNRelu(self, d, db, prefix + "_depthwise_BN", eps, relu=depth_activation)
        p = E.Buf(self, N, ho, wo, filters, name=prefix + "_pointwise")
        E.Conv(self, db, p, prefix + "_pointwise", 1, init="glorot_uniform")
        pb = out if out is not None else E.Buf(self, N, ho, wo, filters, name=prefix + "_pointwise_BN")
        E.BNRelu(self, p, pb, prefix + "_pointwise_BN", eps, relu=depth_activation)
        return pb

    def _xception_block(self, x, depth_list, prefix, skip_type, stride, rate=1, depth_activation=False):
        """_xception_block of impl/deeplab/model.py:177-216; returns (outputs, skip = tensor after the second SepConv)"""
        N = self.batch
        r, skip = x, None
        for i in range(3):
            r = self._sepconv_bn(r, depth_list[i], prefix + "_separable_conv%d" % (i + 1), stride=stride if i == 2 else 1, rate=rate,
                                 depth_activation=depth_activation)
            if i == 1:
                skip = r
        if skip_type == "none":
            return r, skip
        out = E.Buf(self, N, r.h, r.w, r.c, name=prefix)
        if skip_type == "conv":
            sc = E.Buf(self, N, r.h, r.w, r.c, name=prefix + "_shortcut")
            E.Conv(self, x, sc, prefix + "_shortcut", 1, stride=stride, pad=0, init="glorot_uniform")
            scb = E.Buf(self, N, r.h, r.w, r.c, name=prefix + "_shortcut_BN")
            E.BNRelu(self, sc, scb, prefix + "_shortcut_BN", 1e-3, relu=False)
            E.Add(self, r, scb, out)
        else:
            E.Add(self, r, x, out)
        return out, skip

    def _build_deeplabv3_xception(self, classes, input_shape, batch, activation, loss, dropout, OS):
        """DeepLabV3+ over the modified aligned Xception of the reference's in-tree model (impl/deeplab/model.py:339-383 entry /
        middle (16 units) / exit flow; ASPP with image pooling, 1x1 and three atrous separable branches :457-481; decoder with the
        1/4-resolution skip :488-498; probability head :494-500).  OS = 16 (schema default) or 8 (:340-349)."""
        if OS not in (8, 16):
            raise ValueError("DeepLabV3 / xception: OS must be 8 or 16")
        H, W, CI = input_shape
        if H % 16 or W % 16:
            raise ValueError("DeepLabV3 / xception: input height/width must be divisible by 16")
        if not 1 <= CI <= 4:
            raise NotImplementedError("input channels: 1..4 are built (uint8 augmentation / stem kernels); shape[2] = %d" % CI)
        N = batch
        self.architecture, self.input_shape, self.classes, self.backbone = "DeepLabV3", (H, W, CI), classes, "xception"
        self.img = E.Buf(self, N, H, W, CI, E.U8, name="image")
        self.mask = E.Buf(self, N, H, W, classes, E.U8, name="mask")
        s3, mid, ex = (1, 2, (2, 4)) if OS == 8 else (2, 1, (1, 2))
        rates = (12, 24, 36) if OS == 8 else (6, 12, 18)
        init = "glorot_uniform"
        x0 = E.Buf(self, N, H, W, 8, name="input_bf16")
        E.InputCast(self, self.img, x0)
        h, w = H // 2, W // 2
        z = E.Buf(self, N, h, w, 32, name="entry_flow_conv1_1")
        E.Conv(self, x0, z, "entry_flow_conv1_1", 3, stride=2, pad=0, init=init, needs_dgrad=False, cin_real=CI)   # TF 'same', even size
        a = E.Buf(self, N, h, w, 32, name="entry_flow_conv1_1_relu")
        E.BNRelu(self, z, a, "entry_flow_conv1_1_BN", 1e-3)
        z = E.Buf(self, N, h, w, 64, name="entry_flow_conv1_2")
        E.Conv(self, a, z, "entry_flow_conv1_2", 3, pad=1, init=init)
        x = E.Buf(self, N, h, w, 64, name="entry_flow_conv1_2_relu")
        E.BNRelu(self, z, x, "entry_flow_conv1_2_BN", 1e-3)
        x, _ = self._xception_block(x, [128, 128, 128], "entry_flow_block1", "conv", 2)
        x, skip1 = self._xception_block(x, [256, 256, 256], "entry_flow_block2", "conv", 2)
        x, _ = self._xception_block(x, [728, 728, 728], "entry_flow_block3", "conv", s3)
        for i in range(16):
            x, _ = self._xception_block(x, [728, 728, 728], "middle_flow_unit_%d" % (i + 1), "sum", 1, rate=mid)
        x, _ = self._xception_block(x, [728, 1024, 1024], "exit_flow_block1", "conv", 1, rate=ex[0])
        x, _ = self._xception_block(x, [1536, 1536, 2048], "exit_flow_block2", "none", 1, rate=ex[1], depth_activation=True)
        self.encoder_param_names = list(self.params.keys())
        EPS = 1e-5
        h, w = x.h, x.w
        cat = E.Buf(self, N, h, w, 5 * 256, name="aspp_concat")
        gp = E.Buf(self, N, 1, 1, x.c, name="image_pooling_avg")
        E.GlobalAvgPool(self, x, gp)
        ip = E.Buf(self, N, 1, 1, 256, name="image_pooling")
        E.Conv(self, gp, ip, "image_pooling", 1, init=init)
        ipr = E.Buf(self, N, 1, 1, 256, name="image_pooling_relu")
        E.BNRelu(self, ip, ipr, "image_pooling_BN", EPS)
        E.Broadcast(self, ipr, cat.slice(0, 256, name="image_pooling_up"))
        a0 = E.Buf(self, N, h, w, 256, name="aspp0")
        E.Conv(self, x, a0, "aspp0", 1, init=init)
        E.BNRelu(self, a0, cat.slice(256, 256, name="aspp0_activation"), "aspp0_BN", EPS)
        for k, r in enumerate(rates):
            self._sepconv_bn(x, 256, "aspp%d" % (k + 1), rate=r, depth_activation=True, eps=EPS,
                             out=cat.slice(512 + 256 * k, 256, name="aspp%d_out" % (k + 1)))
        cp = E.Buf(self, N, h, w, 256, name="concat_projection")
        E.Conv(self, cat, cp, "concat_projection", 1, init=init)
        cpr = E.Buf(self, N, h, w, 256, name="concat_projection_relu")
        E.BNRelu(self, cp, cpr, "concat_projection_BN", EPS)
        self.dropout = E.Dropout(self, cpr, dropout, salt=0xD0)
        # DeepLab v3+ decoder at 1/4 resolution
        h4, w4 = H // 4, W // 4
        cat2 = E.Buf(self, N, h4, w4, 256 + 48, name="decoder_concat")
        E.Resize(self, cpr, cat2.slice(0, 256, name="decoder_up"), align_corners=True)
        fp = E.Buf(self, N, h4, w4, 48, name="feature_projection0")
        E.Conv(self, skip1, fp, "feature_projection0", 1, init=init)
        E.BNRelu(self, fp, cat2.slice(256, 48, name="feature_projection0_relu"), "feature_projection0_BN", EPS)
        y = self._sepconv_bn(cat2, 256, "decoder_conv0", depth_activation=True, eps=EPS)
        y = self._sepconv_bn(y, 256, "decoder_conv1", depth_activation=True, eps=EPS)
        self.bufs.setdefault("decoder_conv1", y)
        self.head = E.ProbHead(self, y, classes, (H, W), activation or "none",
                               "logits_semantic" if classes == 21 else "custom_logits_semantic", init=init)
        self.loss = E.Loss(self, self.head, self.mask, *loss)
        self.finalize()

    def _build_vgg16_unet(self, classes, input_shape, batch, decoder_filters, dec_init, loss):
        """U-Net over keras.applications.VGG16 (BASELINE.json configs[0]): 3x3 'same' conv + bias + ReLU blocks
        (2,2,3,3,3), 2x2/2 max pooling, no BatchNorm, raw 0..255 input; skips block5_conv3 .. block1_conv2, decoder
        input block5_pool.  Reference: segmentation.py:109-117 -> segmentation_models.Unet."""
        H, W, CI = input_shape
        if H % 32 or W % 32:
            raise ValueError("input height/width must be divisible by 32")
        if len(decoder_filters) != 5:
            raise ValueError("decoder_filters must have 5 entries")
        N = batch
        df = list(decoder_filters)
        self.input_shape, self.classes, self.backbone = (H, W, CI), classes, "vgg16"
        self.img = E.Buf(self, N, H, W, CI, E.U8, name="image")
        self.mask = E.Buf(self, N, H, W, classes, E.U8, name="mask")
        skip_c = [512, 512, 256, 128, 64]
        up_c = [512] + df[:4]
        cat = []
        for i in range(5):
            hh, ww = (H // (16 >> i), W // (16 >> i)) if i < 4 else (H, W)
            cat.append(E.Buf(self, N, hh, ww, up_c[i] + skip_c[i], name="cat%d" % i))
        x = E.Buf(self, N, H, W, 8, name="input_bf16")
        E.InputCast(self, self.img, x)
        h, w = H, W
        for bi, (f, reps) in enumerate(VGG16_BLOCKS):
            for ci in range(reps):
                name = "block%d_conv%d" % (bi + 1, ci + 1)
                last_of_block = ci == reps - 1
                ci_cat = 4 - bi  # block1 -> cat4 ... block5 -> cat0
                y = cat[ci_cat].slice(up_c[ci_cat], f, name=name) if last_of_block else E.Buf(self, N, h, w, f, name=name)
                first = bi == 0 and ci == 0
                E.Conv(self, x, y, name, 3, pad=1, bias=True, relu=True, init="glorot_uniform", needs_dgrad=not first,
                       cin_real=CI if first else None)
                x = y
            p = E.Buf(self, N, h // 2, w // 2, f, name="block%d_pool" % (bi + 1))
            E.MaxPool(self, x, p, 2, 2, 0)
            x, h, w = p, h // 2, w // 2
        self.encoder_param_names = list(self.params.keys())
        E.UpCopy(self, x, cat[0].slice(0, up_c[0], name="block5_pool_up"))
        self._build_decoder(cat, df, classes, dec_init, loss)
