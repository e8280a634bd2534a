"""K1 augmentation parity: drawn parameters vs oracle/philox+augment (discrete draws exact, matrices to 1e-12),
pixels/indices BIT EXACT vs cv2.warpAffine (the backend imgaug calls) given the device-drawn matrices."""
import ctypes as C

import numpy as np
import pytest
import torch

from segmentation_training_pipeline_b200 import lib
from tests.util import stream

pytestmark = pytest.mark.gpu


def _spec():
    from oracle import augment as OA
    o = OA.AugSpec(fliplr=0.5, flipud=0.5, affine=True, scale=(0.8, 1.5), translate_x=(-0.2, 0.2),
                   translate_y=(-0.2, 0.2), rotate=(-16, 16), shear=(-16, 16), multiply=(0.8, 1.2), add=(-10, 10))
    c = lib.AugSpec(0.5, 0.5, 1, 0.8, 1.5, -0.2, 0.2, -0.2, 0.2, -16, 16, -16, 16, 1, 0.8, 1.2, 1, -10, 10, 0)
    return o, c


def _draw(stp, cuda, cspec, seed, step, n, pool, h, w):
    d_step = torch.tensor([step], dtype=torch.int64, device=cuda)
    buf = torch.zeros(n * C.sizeof(lib.AugSample), dtype=torch.uint8, device=cuda)
    stp.augment_draw(C.byref(cspec), seed, d_step.data_ptr(), n, pool, h, w, buf.data_ptr(), stream())
    host = buf.cpu().numpy().tobytes()
    return buf, [lib.AugSample.from_buffer_copy(host, i * C.sizeof(lib.AugSample)) for i in range(n)]


@pytest.mark.parametrize("h,w", [(128, 128), (96, 132), (512, 512)])
def test_draw_matches_oracle(stp, cuda, h, w):
    from oracle import augment as OA
    ospec, cspec = _spec()
    n, pool, seed = 16, 64, 1234
    for step in (0, 1, 7, 2 ** 33 + 5):
        _, samples = _draw(stp, cuda, cspec, seed, step, n, pool, h, w)
        for i, s in enumerate(samples):
            sid = (step * n + i) % pool
            p = OA.draw_params(ospec, seed, step, sid, h, w)
            assert s.src_index == sid
            assert bool(s.fliplr) == p.fliplr and bool(s.flipud) == p.flipud
            assert s.add == p.add
            assert np.float32(s.mul) == np.float32(p.mul)
            m = np.array(list(s.m)).reshape(2, 3)
            assert np.allclose(m, p.matrix, rtol=1e-12, atol=1e-10), (m, p.matrix)
            inv = OA.invert_affine(m)
            assert np.allclose(np.array(list(s.inv)), np.array(inv), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("h,w", [(128, 128), (96, 132), (512, 512)])
def test_apply_bit_exact_vs_cv2(stp, cuda, h, w):
    from oracle import augment as OA
    ospec, cspec = _spec()
    n, pool, seed = 8, 8, 99
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (pool, h, w, 3), dtype=np.uint8)
    masks = (rng.random((pool, h, w, 1)) > 0.6).astype(np.uint8)
    d_img, d_msk = torch.from_numpy(imgs).to(cuda), torch.from_numpy(masks).to(cuda)
    total_bad = 0
    for step in (0, 3):
        buf, samples = _draw(stp, cuda, cspec, seed, step, n, pool, h, w)
        out_i = torch.zeros((n, h, w, 3), dtype=torch.uint8, device=cuda)
        out_m = torch.zeros((n, h, w, 1), dtype=torch.uint8, device=cuda)
        stp.augment_apply(d_img.data_ptr(), d_msk.data_ptr(), buf.data_ptr(), out_i.data_ptr(), out_m.data_ptr(), n, h, w,
                          3, 1, 0, stream())
        gi, gm = out_i.cpu().numpy(), out_m.cpu().numpy()
        for i, s in enumerate(samples):
            p = OA.SampleParams(bool(s.fliplr), bool(s.flipud), np.array(list(s.m)).reshape(2, 3), True, float(s.mul),
                                True, int(s.add))
            ri, rm = OA.apply(imgs[s.src_index], masks[s.src_index], p, use_cv2=True)
            total_bad += int((ri != gi[i]).sum()) + int((rm != gm[i]).sum())
    assert total_bad == 0


def test_identity_and_flip_only(stp, cuda):
    from oracle import augment as OA
    h, w, n = 40, 36, 4  # w % 4 == 0 vector path; also the scalar path below
    rng = np.random.default_rng(1)
    for ww in (w, 37):
        imgs = rng.integers(0, 256, (n, h, ww, 3), dtype=np.uint8)
        masks = rng.integers(0, 2, (n, h, ww, 1), dtype=np.uint8)
        cspec = lib.AugSpec(0.5, 0.5, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0)
        ospec = OA.AugSpec(fliplr=0.5, flipud=0.5)
        buf, samples = _draw(stp, cuda, cspec, 5, 2, n, n, h, ww)
        d_img, d_msk = torch.from_numpy(imgs).to(cuda), torch.from_numpy(masks).to(cuda)
        out_i = torch.zeros((n, h, ww, 3), dtype=torch.uint8, device=cuda)
        out_m = torch.zeros((n, h, ww, 1), dtype=torch.uint8, device=cuda)
        stp.augment_apply(d_img.data_ptr(), d_msk.data_ptr(), buf.data_ptr(), out_i.data_ptr(), out_m.data_ptr(), n, h,
                          ww, 3, 1, 0, stream())
        for i, s in enumerate(samples):
            p = OA.draw_params(ospec, 5, 2, s.src_index, h, ww)
            ri, rm = OA.apply(imgs[s.src_index], masks[s.src_index], p)
            assert np.array_equal(ri, out_i[i].cpu().numpy()) and np.array_equal(rm, out_m[i].cpu().numpy())


@pytest.mark.parametrize("pre", [0, 3, 1, 2])
@pytest.mark.parametrize("order", [(0, 1, 2), (2, 1, 0), (1, 2, 0)])
@pytest.mark.parametrize("affine", [0, 1])
def test_rotate90_invert_and_colour_order(stp, cuda, order, affine, pre):
    """musket's Rotate90 (np.rot90 by a uniform k), imgaug Invert(p) and the colour stage in YAML order (saturating uint8
    ops do not commute): draws equal to the oracle's, pixels / mask indices bit exact.  pre = which flips are listed BEFORE
    Rotate90 in the YAML block (bit 0 Fliplr, bit 1 Flipud; the reference's examples list both first): the oracle applies the
    literal Sequential order, the kernel its rot90 -> flips gather with the flags swapped on odd quarter turns."""
    from oracle import augment as OA
    h = w = 96
    ospec = OA.AugSpec(fliplr=0.5, flipud=0.5, affine=bool(affine), scale=(0.8, 1.3), translate_x=(-0.1, 0.1), translate_y=(-0.1, 0.1),
                       rotate=(-20, 20), shear=(-8, 8), multiply=(0.5, 1.9), add=(-60, 60), rot90=True, invert=0.5, color_order=order,
                       flip_before_rot90=pre)
    cspec = lib.AugSpec(0.5, 0.5, affine, 0.8, 1.3, -0.1, 0.1, -0.1, 0.1, -20, 20, -8, 8, 1, 0.5, 1.9, 1, -60, 60, 0, 1, 0.5)
    for i, o in enumerate(order):
        cspec.color_order[i] = o
    cspec.flip_before_rot90 = pre
    n, pool, seed = 16, 16, 7
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (pool, h, w, 3), dtype=np.uint8)
    masks = (rng.random((pool, h, w, 1)) > 0.6).astype(np.uint8)
    d_img, d_msk = torch.from_numpy(imgs).to(cuda), torch.from_numpy(masks).to(cuda)
    ks, invs, bad = set(), set(), 0
    for step in (0, 5):
        buf, samples = _draw(stp, cuda, cspec, seed, step, n, pool, h, w)
        out_i = torch.zeros((n, h, w, 3), dtype=torch.uint8, device=cuda)
        out_m = torch.zeros((n, h, w, 1), dtype=torch.uint8, device=cuda)
        stp.augment_apply(d_img.data_ptr(), d_msk.data_ptr(), buf.data_ptr(), out_i.data_ptr(), out_m.data_ptr(), n, h, w,
                          3, 1, 0, stream())
        gi, gm = out_i.cpu().numpy(), out_m.cpu().numpy()
        for i, s in enumerate(samples):
            p = OA.draw_params(ospec, seed, step, s.src_index, h, w)
            assert (s.flags2 & 3) == p.rot90_k and bool(s.flags2 & 4) == p.invert
            assert tuple((s.flags2 >> (4 + 2 * k)) & 3 for k in range(3)) == tuple(order)
            ks.add(p.rot90_k)
            invs.add(p.invert)
            p.matrix = np.array(list(s.m)).reshape(2, 3)   # pixels are pinned given the device-drawn matrix
            ri, rm = OA.apply(imgs[s.src_index], masks[s.src_index], p, use_cv2=True)
            bad += int((ri != gi[i]).sum()) + int((rm != gm[i]).sum())
    assert bad == 0
    assert ks == {0, 1, 2, 3} and invs == {False, True}


PIX_BLOCKS = [
    {"AddElementwise": {"range": [-20, 20], "per_channel": 0.5}},
    {"Multiply": [0.5, 1.5], "MultiplyElementwise": {"range": [0.5, 1.5], "per_channel": 1.0}, "Add": [-30, 30]},
    {"Dropout": {"p": [0.0, 0.3], "per_channel": 0.5}, "Invert": 0.5},
    {"Grayscale": {"alpha": [0.0, 1.0]}, "AdditiveGaussianNoise": {"scale": [0.0, 25.0], "per_channel": 0.5}},
    {"Fliplr": 0.5, "Sequential": [{"Add": [-5, 5]}, {"OneOf": [{"Dropout": 0.2}, {"Grayscale": 1.0}, {"AddElementwise": [-40, 40]}]}]},
]


@pytest.mark.parametrize("block", PIX_BLOCKS, ids=[str(i) for i in range(len(PIX_BLOCKS))])
def test_pixelwise_augmenters_match_oracle(cuda, block):
    """AddElementwise / MultiplyElementwise / Dropout / AdditiveGaussianNoise / Grayscale, Sequential and OneOf, mixed with
    Multiply / Add / Invert in YAML order, through Trainer.run_augment: every pixel against oracle.augment.apply_pixel_ops
    (same Philox counters, same fp32 operation order).  Bit exact except AdditiveGaussianNoise, whose logf / cosf differ from
    numpy's in the last ulp: there at most 1e-4 of the values may differ, by 1."""
    from oracle import augment as OA
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.segmentation import parse_augmentation
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size, pool, seed = 4, 64, 8, 31
    net = SegNet("resnet18", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0)
    cfg = parse_augmentation(block, seed=seed)
    assert cfg.pix_ops
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (pool, size, size, 3), dtype=np.uint8)
    masks = (rng.random((pool, size, size, 1)) > 0.5).astype(np.uint8)
    tr = Trainer(net, augment=cfg)
    tr.set_pool(torch.from_numpy(imgs), torch.from_numpy(masks))
    ospec = OA.AugSpec(fliplr=cfg.fliplr, flipud=cfg.flipud, multiply=cfg.multiply, add=cfg.add, invert=cfg.invert)
    bad = tot = changed = 0
    picks = set()
    for step in (0, 3):
        net.d_step.fill_(step)
        tr.run_augment()
        torch.cuda.synchronize()
        gi = net.img.storage.view(n, size, size, 3).cpu().numpy()
        gm = net.mask.storage.view(n, size, size, 1).cpu().numpy()
        for i in range(n):
            sid = (step * n + i) % pool
            p = OA.draw_params(ospec, seed, step, sid, size, size)
            geo = OA.SampleParams(p.fliplr, p.flipud, p.matrix, False, 1.0, False, 0, 0, False, (0, 1, 2), 0)
            base, bm = OA.apply(imgs[sid], masks[sid], geo)                    # geometric part only (flips)
            ref = OA.apply_pixel_ops(base, cfg.pix_ops, seed, step, sid, p)
            d = np.abs(ref.astype(int) - gi[i].astype(int))
            assert d.max() <= 1, (step, i, d.max())
            bad += int((d != 0).sum())
            tot += d.size
            changed += int((gi[i] != base).sum())
            assert np.array_equal(gm[i], bm)
            picks.add(bytes(gi[i][:2, :2].tobytes()))
    has_gauss = any(o[0] == 6 for o in cfg.pix_ops)
    assert bad <= (1e-4 * tot if has_gauss else 0), (bad, tot)
    assert changed > 0.05 * tot          # the block really changes pixels


NB_BLOCKS = [
    {"GaussianBlur": {"sigma": [0.0, 3.0]}},
    {"Fliplr": 0.5, "AverageBlur": {"k": [2, 9]}},
    {"MedianBlur": {"k": [3, 7]}},
    {"Sharpen": {"alpha": [0.0, 1.0], "lightness": [0.75, 1.5]}},
    {"Emboss": {"alpha": [0.0, 1.0], "strength": [0.0, 2.0]}, "EdgeDetect": {"alpha": [0.0, 0.7]}},
    {"DirectedEdgeDetect": {"alpha": [0.2, 1.0], "direction": [0.0, 1.0]}},
    {"Fliplr": 0.5, "Multiply": [0.8, 1.2], "GaussianBlur": {"sigma": [0.5, 5.5]}, "Add": [-10, 10],
     "OneOf": {"AverageBlur": {"k": [2, 7]}, "MedianBlur": {"k": [3, 5]}, "Sharpen": {"alpha": [0.2, 1.0], "lightness": [0.75, 1.5]}},
     "Dropout": {"p": [0.0, 0.05]}, "Emboss": {"alpha": [0.1, 1.0], "strength": [0.5, 1.5]}},
]


@pytest.mark.parametrize("block", NB_BLOCKS, ids=[str(i) for i in range(len(NB_BLOCKS))])
@pytest.mark.parametrize("channels", [3, 1])
def test_neighbourhood_augmenters_match_oracle(cuda, block, channels):
    """GaussianBlur / AverageBlur / MedianBlur / Sharpen / Emboss / EdgeDetect, alone, inside OneOf and interleaved with pixel-wise
    augmenters in YAML order, through Trainer.run_augment: EVERY pixel equals the oracle twin (oracle.augment.apply_neighbourhood_op,
    whose arithmetic is pinned against the real cv2 in tests/test_cpu_oracle.py); masks untouched."""
    from oracle import augment as OA
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.segmentation import parse_augmentation
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size, pool, seed = 4, 64, 8, 77
    net = SegNet("resnet18", classes=1, input_shape=(size, size, channels), batch=n, device="cuda:0", seed=0)
    cfg = parse_augmentation(block, seed=seed)
    assert cfg.colour_seq
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (pool, size, size, channels), dtype=np.uint8)
    imgs[:, 20:40, 10:30] //= 4                                  # some structure for the edge / median filters
    masks = (rng.random((pool, size, size, 1)) > 0.5).astype(np.uint8)
    tr = Trainer(net, augment=cfg)
    tr.set_pool(torch.from_numpy(imgs), torch.from_numpy(masks))
    ospec = OA.AugSpec(fliplr=cfg.fliplr, flipud=cfg.flipud, multiply=cfg.multiply, add=cfg.add, invert=cfg.invert)
    changed = tot = 0
    for step in (0, 5):
        net.d_step.fill_(step)
        tr.run_augment()
        torch.cuda.synchronize()
        gi = net.img.storage.view(n, size, size, channels).cpu().numpy()
        gm = net.mask.storage.view(n, size, size, 1).cpu().numpy()
        for i in range(n):
            sid = (step * n + i) % pool
            p = OA.draw_params(ospec, seed, step, sid, size, size)
            geo = OA.SampleParams(p.fliplr, p.flipud, p.matrix, False, 1.0, False, 0, 0, False, (0, 1, 2), 0)
            ref, bm = OA.apply(imgs[sid], masks[sid], geo)
            base = ref
            for typ, k, payload in cfg.colour_runs():
                if typ == "pix":
                    ref = OA.apply_pixel_ops(ref, payload, seed, step, sid, p, k_base=k)
                else:
                    kind, a, b, c, d, gid, gsz, gm_ = payload
                    ref = OA.apply_neighbourhood_op(ref, (kind, a, b, c, d, k, gid, gsz, gm_), seed, step, sid)
            assert np.array_equal(gi[i], ref), (step, i, int((gi[i] != ref).sum()), int(np.abs(gi[i].astype(int) - ref.astype(int)).max()))
            assert np.array_equal(gm[i], bm)
            changed += int((gi[i] != base).sum())
            tot += base.size
    assert changed > 0.05 * tot


def test_neighbourhood_augmenter_in_the_captured_step(cuda):
    """a blur inside the whole-step CUDA graph: the step runs, the loss is finite and differs from the un-blurred run"""
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.segmentation import parse_augmentation
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size = 4, 64
    rng = np.random.default_rng(1)
    imgs = torch.from_numpy(rng.integers(0, 256, (8, size, size, 3), dtype=np.uint8))
    masks = torch.from_numpy((rng.random((8, size, size, 1)) > 0.5).astype(np.uint8))
    losses = []
    for block in ({"Fliplr": 0.5}, {"Fliplr": 0.5, "GaussianBlur": {"sigma": [1.0, 2.0]}, "Sharpen": {"alpha": 0.5}}):
        net = SegNet("resnet18", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0)
        tr = Trainer(net, optimizer="Adam", lr=1e-3, augment=parse_augmentation(block, seed=3))
        tr.set_pool(imgs, masks)
        tr.capture()
        for _ in range(3):
            tr.step()
        losses.append(tr.loss_value())
    assert np.isfinite(losses).all() and abs(losses[0] - losses[1]) > 1e-6, losses
