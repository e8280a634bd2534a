"""The reference's README example (README.md:116-126, 493-534), unchanged API:

    python examples/train.py images/ masks/                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train.py images/ masks/   # cfg.gpus = 8
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmentation_pipeline import segmentation                              # noqa: E402
from segmentation_pipeline.impl.datasets import SimplePNGMaskDataSet       # noqa: E402
from segmentation_pipeline.impl.rle import rle_encode                       # noqa: E402


def main():
    images, masks = sys.argv[1], sys.argv[2]
    cfg = segmentation.parse(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config.yaml"))
    ds = SimplePNGMaskDataSet(images, masks)
    cfg.fit(ds)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    print(cfg.info())
    predictions, names = [], []

    def on_predict(file_name, img, data):
        data["pred"].append(rle_encode(img.arr[:, :, 0] > 0.5))           # img.arr: float probabilities at the image's size
        data["images"].append(file_name[:file_name.index(".")])

    # ensemble of all folds of the last stage, flip test-time augmentation
    cfg.predict_in_directory(images, list(range(cfg.folds_count)), len(cfg.stages) - 1, on_predict,
                             {"pred": predictions, "images": names}, ttflips=True)
    print("%d masks encoded" % len(predictions))


if __name__ == "__main__":
    main()
