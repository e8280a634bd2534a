"""stp_resize_u8 (csrc/resize_u8.cu): cv2.resize arithmetic on the device.  BIT EXACT against oracle/resize.py (OpenCV's scalar
path restated) for cubic and nearest, incl. zero-padded / cropped virtual windows and batches of differently sized sources;
against the real cv2 (IPP off) identical except at rounding ties of cv2's SIMD path (<= 0.05 % of pixels, |d| = 1)."""
import ctypes as C

import numpy as np
import pytest
import torch

from segmentation_training_pipeline_b200 import lib
from tests.util import stream

pytestmark = pytest.mark.gpu


def _run(stp, cuda, images, windows, H, W, mode):
    """images: list of uint8 [h, w, C]; windows: list of (vy0, vx0, vh, vw) or None"""
    c = images[0].shape[2]
    offs, blobs, off = [], [], 0
    for im in images:
        offs.append(off)
        blobs.append(np.ascontiguousarray(im).reshape(-1))
        off += (im.size + 15) // 16 * 16
    arena = np.zeros(off, np.uint8)
    for o, b in zip(offs, blobs):
        arena[o:o + b.size] = b
    items = (lib.ResizeItem * len(images))()
    for i, (im, wdw) in enumerate(zip(images, windows)):
        vy0, vx0, vh, vw = wdw if wdw is not None else (0, 0, im.shape[0], im.shape[1])
        items[i] = lib.ResizeItem(offs[i], im.shape[0], im.shape[1], vy0, vx0, vh, vw)
    d_arena = torch.from_numpy(arena).to(cuda)
    d_items = torch.from_numpy(np.frombuffer(bytes(items), dtype=np.uint8).copy()).to(cuda)
    out = torch.zeros((len(images), H, W, c), dtype=torch.uint8, device=cuda)
    stp.resize_u8(d_arena.data_ptr(), d_items.data_ptr(), len(images), c, out.data_ptr(), H, W, mode, stream())
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("c", [1, 3, 4])
def test_resize_matches_oracle_and_cv2(stp, cuda, c):
    import cv2
    from oracle import resize as OR
    rng = np.random.default_rng(c)
    H, W = 96, 128
    sizes = [(37, 53), (300, 200), (96, 128), (64, 64), (100, 333), (513, 257)]
    imgs = [rng.integers(0, 256, (h, w, c), dtype=np.uint8) for h, w in sizes]
    got = _run(stp, cuda, imgs, [None] * len(imgs), H, W, lib.RESIZE_CUBIC)
    gotn = _run(stp, cuda, imgs, [None] * len(imgs), H, W, lib.RESIZE_NEAREST)
    prev = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        tot = bad = 0
        for i, im in enumerate(imgs):
            assert np.array_equal(got[i], OR.resize_cubic_u8(im, H, W)), sizes[i]          # bit exact vs the restated scalar path
            assert np.array_equal(gotn[i], OR.resize_nearest_u8(im, H, W)), sizes[i]
            ref = cv2.resize(im, (W, H), interpolation=cv2.INTER_CUBIC).reshape(H, W, c)
            d = np.abs(got[i].astype(int) - ref.astype(int))
            assert d.max() <= 1
            tot += d.size
            bad += int((d != 0).sum())
            assert np.array_equal(gotn[i], cv2.resize(im, (W, H), interpolation=cv2.INTER_NEAREST).reshape(H, W, c))
        assert bad <= 5e-4 * tot, (bad, tot)
    finally:
        cv2.ipp.setUseIPP(prev)


def test_resize_virtual_windows_crop_and_pad(stp, cuda):
    """the crop / pad family: a window of the stored image (crop), a window larger than it (constant zero padding), both"""
    from oracle import resize as OR
    rng = np.random.default_rng(9)
    H, W = 64, 80
    im = rng.integers(1, 256, (50, 70, 3), dtype=np.uint8)
    wins = [(5, 7, 30, 40), (-6, -9, 64, 90), (10, -4, 64, 80), (0, 0, 50, 70), (-3, 20, 20, 60)]
    got = _run(stp, cuda, [im] * len(wins), wins, H, W, lib.RESIZE_CUBIC)
    gotn = _run(stp, cuda, [im] * len(wins), wins, H, W, lib.RESIZE_NEAREST)
    for i, wdw in enumerate(wins):
        v = OR.window(im, *wdw)
        assert np.array_equal(got[i], OR.resize_cubic_u8(v, H, W)), wdw
        assert np.array_equal(gotn[i], OR.resize_nearest_u8(v, H, W)), wdw


CP_BLOCKS = [
    [(lib.CP_CROP_TO_FIXED, 0, 48, 40, 0, 0)],
    [(lib.CP_PAD_TO_FIXED, 0, 100, 90, 0, 0), (lib.CP_CROP_TO_FIXED, 0, 64, 64, 0, 0)],      # pure index work: no resampling
    [(lib.CP_PAD, 0, 3, 5, 7, 2)],
    [(lib.CP_CROP_AND_PAD, 1, -0.2, 0.25, 0, 0)],
    [(lib.CP_CROP_TO_FIXED, 0, 50, 64, 0, 0), (lib.CP_CROP_AND_PAD, 0, 0.1, -0.05, 0.0, 0.2)],
]


@pytest.mark.parametrize("ops", CP_BLOCKS)
def test_crop_pad_augmenters_match_oracle(stp, cuda, ops):
    """Pad / PadToFixedSize / CropToFixedSize / CropAndPad: windows drawn on the device (Philox calls 6-7) equal the oracle's,
    pixels (cubic) and mask indices (nearest) bit exact against oracle/resize.py."""
    from oracle import augment as OA
    rng = np.random.default_rng(len(ops) * 7 + int(ops[0][0]))
    n, pool, H, W, seed = 8, 12, 64, 64, 4242
    imgs = rng.integers(0, 256, (pool, H, W, 3), dtype=np.uint8)
    masks = (rng.random((pool, H, W, 1)) > 0.6).astype(np.uint8)
    d_img, d_msk = torch.from_numpy(imgs).to(cuda), torch.from_numpy(masks).to(cuda)
    spec = lib.CropPadSpec()
    spec.n_ops = len(ops)
    for i, o in enumerate(ops):
        spec.ops[i] = lib.CropPadOp(*[int(o[0]), int(o[1])] + [float(v) for v in o[2:]])
    isz = C.sizeof(lib.ResizeItem)
    for step in (0, 3, 2 ** 33 + 1):
        d_step = torch.tensor([step], dtype=torch.int64, device=cuda)
        items = torch.zeros(2 * n * isz, dtype=torch.uint8, device=cuda)
        stp.croppad_draw(C.byref(spec), seed, d_step.data_ptr(), n, pool, H, W, 3, 1, items.data_ptr(), items.data_ptr() + n * isz, stream())
        oi = torch.zeros((n, H, W, 3), dtype=torch.uint8, device=cuda)
        om = torch.zeros((n, H, W, 1), dtype=torch.uint8, device=cuda)
        stp.resize_u8(d_img.data_ptr(), items.data_ptr(), n, 3, oi.data_ptr(), H, W, lib.RESIZE_CUBIC, stream())
        stp.resize_u8(d_msk.data_ptr(), items.data_ptr() + n * isz, n, 1, om.data_ptr(), H, W, lib.RESIZE_NEAREST, stream())
        torch.cuda.synchronize()
        host = items.cpu().numpy().tobytes()
        gi, gm = oi.cpu().numpy(), om.cpu().numpy()
        distinct = set()
        for i in range(n):
            it = lib.ResizeItem.from_buffer_copy(host, i * isz)
            sid = (step * n + i) % pool
            win = OA.crop_pad_window(ops, seed, step, sid, H, W)
            assert (it.vy0, it.vx0, it.vh, it.vw) == win and it.src_off == sid * H * W * 3 and (it.sh, it.sw) == (H, W), (i, win)
            ri, rm = OA.apply_crop_pad(imgs[sid], masks[sid], win, H, W)
            assert np.array_equal(gi[i], ri) and np.array_equal(gm[i], rm), (step, i, win)
            distinct.add(win)
        if any(o[0] in (lib.CP_PAD_TO_FIXED, lib.CP_CROP_TO_FIXED) or o[1] for o in ops):
            assert len(distinct) > 1    # positions / percentages really are random per sample


def test_crop_pad_stage_in_the_training_step(cuda):
    """YAML block `CropToFixedSize, Fliplr` through Trainer.run_augment (what the captured step launches): crop / pad stage ->
    staging batch -> fused flip kernel, against the oracle chain."""
    from oracle import augment as OA
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.segmentation import parse_augmentation
    from segmentation_training_pipeline_b200.trainer import Trainer
    n, size, pool = 4, 64, 6
    net = SegNet("resnet18", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0)
    cfg = parse_augmentation({"CropToFixedSize": {"width": 40, "height": 48}, "Fliplr": 1.0}, seed=77)
    assert cfg.crop_pad == ((3, 0, 40, 48, 0, 0),) and cfg.fliplr == 1.0
    rng = np.random.default_rng(2)
    imgs = rng.integers(0, 256, (pool, size, size, 3), dtype=np.uint8)
    masks = (rng.random((pool, size, size, 1)) > 0.5).astype(np.uint8)
    tr = Trainer(net, augment=cfg)
    tr.set_pool(torch.from_numpy(imgs), torch.from_numpy(masks))
    net.d_step.fill_(5)
    tr.run_augment()
    torch.cuda.synchronize()
    gi = net.img.storage.view(n, size, size, 3).cpu().numpy()
    gm = net.mask.storage.view(n, size, size, 1).cpu().numpy()
    for i in range(n):
        sid = (5 * n + i) % pool
        win = OA.crop_pad_window(cfg.crop_pad, 77, 5, sid, size, size)
        ri, rm = OA.apply_crop_pad(imgs[sid], masks[sid], win, size, size)
        assert np.array_equal(gi[i], ri[:, ::-1]) and np.array_equal(gm[i], rm[:, ::-1]), i


def test_on_device_ingest_of_unresized_samples(cuda, tmp_path):
    """`device_resize: true` -- HostLoader hands out UNRESIZED decoded samples (loader.RawBatch), the
    trainer copies them and stp_resize_u8 writes the network-shape batch into the step's input pool: bit exact against the
    restated cv2 arithmetic, and the host-fed step runs on it."""
    from oracle import resize as OR
    from segmentation_training_pipeline_b200.impl.datasets import PredictionItem
    from segmentation_training_pipeline_b200.loader import HostLoader
    from segmentation_training_pipeline_b200.models import SegNet
    from segmentation_training_pipeline_b200.trainer import Trainer

    rng = np.random.default_rng(3)
    sizes = [(50, 70), (64, 64), (120, 90), (33, 200), (64, 64), (80, 80)]

    class DS:
        def __len__(self):
            return len(sizes)

        def __getitem__(self, i):
            h, w = sizes[i]
            r = np.random.default_rng(100 + i)
            return PredictionItem("s%d" % i, r.integers(0, 256, (h, w, 3), dtype=np.uint8), (r.random((h, w)) > 0.5).astype(np.uint8))

    ds, n, size = DS(), 4, 64
    net = SegNet("resnet18", classes=1, input_shape=(size, size, 3), batch=n, device="cuda:0", seed=0, loss=(1.0, 1.0, 0.0))
    tr = Trainer(net, optimizer="SGD", lr=1e-3)
    tr.enable_host_feed()
    loader = HostLoader(ds, (size, size, 3), 1, n, workers=2, device_resize=True)
    batches = [[0, 1, 2, 3], [4, 5, 0, 2], [3, 1, 5, 4]]
    seen = 0
    for k, (rb, none) in enumerate(loader.iterate(batches)):
        assert none is None and rb.n == n
        tr.step_from_host_pipelined(rb, None)
        torch.cuda.synchronize()
        gi, gm = tr.pool_img.cpu().numpy(), tr.pool_mask.cpu().numpy()
        for j, idx in enumerate(batches[k]):
            it = ds[idx]
            assert np.array_equal(gi[j], OR.resize_cubic_u8(it.x, size, size)), (k, j)
            assert np.array_equal(gm[j][:, :, 0], OR.resize_nearest_u8(it.y, size, size)), (k, j)
        seen += 1
    m = tr.flush_host_pipeline()
    loader.close()
    assert seen == 3 and m is not None and np.isfinite(m["loss"])
    assert tr.last_h2d_bytes > 0
