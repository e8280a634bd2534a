"""Convert a Keras weights file (.h5, `model.save_weights`) into the .npz of Keras-named arrays that `encoder_weights:` /
`initial_weights:` accept here (this image has no h5py; run this where h5py is installed).

    python scripts/keras_h5_to_npz.py deeplabv3_mobilenetv2_tf_dim_ordering_tf_kernels.h5

writes `<same stem>.npz` with keys `<layer>/<weight>` (`Conv/kernel`, `Conv_BN/gamma`, `expanded_conv_depthwise/depthwise_kernel`,
`.../moving_mean`, ...), the names the reference's `model.load_weights(path, by_name=True)` matches on (impl/deeplab/model.py:505-513).
"""
import sys

import numpy as np


def convert(path):
    import h5py
    out = {}
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        for layer in g.attrs["layer_names"]:
            layer = layer.decode() if isinstance(layer, bytes) else layer
            lg = g[layer]
            for wn in lg.attrs["weight_names"]:
                wn = wn.decode() if isinstance(wn, bytes) else wn
                key = wn.split(":")[0]                       # 'Conv/kernel:0' -> 'Conv/kernel'
                if not key.startswith(layer + "/"):          # nested scopes: keep '<layer>/<last component>'
                    key = layer + "/" + key.rsplit("/", 1)[-1]
                out[key] = np.asarray(lg[wn], dtype=np.float32)
    dst = path.rsplit(".", 1)[0] + ".npz"
    np.savez(dst, **out)
    return dst, len(out)


if __name__ == "__main__":
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    print("%s (%d arrays)" % convert(sys.argv[1]))
