"""Golden vectors for the Kaggle RLE helpers, produced by IMPORTING the reference's own module
(/root/reference/segmentation_pipeline/impl/rle.py -- pure numpy apart from skimage.morphology.label, which is stubbed because
scikit-image is not installed here and multi_rle_encode is therefore not part of the fixture).

    python tests/golden/make_rle_golden.py     # needs /root/reference; writes tests/golden/rle_golden.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/segmentation_pipeline/impl/rle.py"


def load_reference():
    sk = types.ModuleType("skimage")
    mo = types.ModuleType("skimage.morphology")
    mo.label = lambda a: (_ for _ in ()).throw(RuntimeError("skimage.morphology.label is stubbed"))
    sk.morphology = mo
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.morphology", mo)
    spec = importlib.util.spec_from_file_location("ref_rle", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ref = load_reference()
    rng = np.random.default_rng(20261017)
    out = {}
    shapes = [(1, 1), (4, 3), (3, 4), (7, 7), (16, 16), (5, 12), (12, 5), (32, 48), (64, 64)]
    for k, shape in enumerate(shapes):
        a = (rng.random(shape) > 0.55).astype(np.uint8)
        if k == 3:
            a[:] = 0
        if k == 4:
            a[:] = 1
        enc = ref.rle_encode(a)
        out["mask_%d" % k] = a
        out["enc_%d" % k] = np.array(enc)
        out["shape_%d" % k] = np.array(shape)
        if enc:   # the reference's decode fails on an empty string (np.asarray of an empty list of str -> float): not a vector
            out["dec_%d" % k] = ref.rle_decode(enc, shape)          # NOTE: (w, h)-shaped for non-square inputs (reference quirk)
    # masks_as_image / masks_as_images on a square shape (the reference raises a broadcast error on non-square ones)
    a1, a2 = (rng.random((9, 9)) > 0.7).astype(np.uint8), (rng.random((9, 9)) > 0.7).astype(np.uint8)
    l = [ref.rle_encode(a1), ref.rle_encode(a2), float("nan")]
    out["mai_in_0"], out["mai_in_1"] = np.array(l[0]), np.array(l[1])
    out["mai_out"] = ref.masks_as_image(l, (9, 9))
    out["mais_out"] = np.stack(ref.masks_as_images(l, (9, 9)))
    out["n"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rle_golden.npz"), **out)
    print("wrote rle_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
