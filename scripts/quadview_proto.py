"""CPU prototype of the "pixel-merged view" planned for the narrow decoder layers -- no GPU, no product code.

A 16-channel NHWC tensor [N, H, W, C] read as [N, H, W/G, G*C] (G = 4: 128-byte rows) turns a 3x3 stride-1 'same' convolution into a
3x3 convolution over pixel GROUPS whose [G*Cin -> G*Cout] weight per (filter row r, group offset d in {-1, 0, +1}) is block sparse:
block (f, e) -- input pixel f of the neighbouring group, output pixel e of this group -- holds the original tap s = G*d + f - e + 1 when
0 <= s <= 2 and is zero otherwise.  `expand_weights` builds those weights, `mma_schedule` lists, per K = 16 step of the merged input
row, the contiguous range of output-pixel blocks it contributes to (the N extent of the tcgen05.mma that step needs), and
`check` verifies on random data that the merged-view convolution with the expanded weights equals the plain convolution.

    python scripts/quadview_proto.py            # prints the schedule for 16 -> 16, G = 4 and the verification error
"""
import numpy as np
import torch
import torch.nn.functional as F


def expand_weights(w_krsc: np.ndarray, G: int) -> np.ndarray:
    """[Cout][3][3][Cin] -> [G*Cout][3][3 (d = -1, 0, +1)][G*Cin]; output channel e*Cout + co, input channel f*Cin + ci"""
    cout, R, S, cin = w_krsc.shape
    assert (R, S) == (3, 3)
    out = np.zeros((G * cout, 3, 3, G * cin), w_krsc.dtype)
    for d in (-1, 0, 1):
        for f in range(G):
            for e in range(G):
                s = G * d + f - e + 1          # input pixel G*(j+d) + f feeds output pixel G*j + e through tap s = (in - out) + 1
                if 0 <= s <= 2:
                    out[e * cout:(e + 1) * cout, :, d + 1, f * cin:(f + 1) * cin] = w_krsc[:, :, s, :]
    return out


def mma_schedule(cin: int, cout: int, G: int):
    """per group offset d: list of (k16 step of the merged row, first output block e_lo, number of blocks) -- one MMA each with
    N = blocks * cout columns written at TMEM column e_lo * cout; steps whose pixel f reaches no output pixel are skipped"""
    sched = {}
    for d in (-1, 0, 1):
        steps = []
        for ks in range(G * cin // 16):
            f = (ks * 16) // cin
            es = [e for e in range(G) if 0 <= G * d + f - e + 1 <= 2]
            if es:
                assert es == list(range(es[0], es[-1] + 1))
                steps.append((ks, es[0], len(es)))
        sched[d] = steps
    return sched


def check(n=2, h=9, w=16, cin=16, cout=16, G=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, cin, generator=g, dtype=torch.float64)
    wt = torch.randn(cout, 3, 3, cin, generator=g, dtype=torch.float64)
    want = F.conv2d(x.permute(0, 3, 1, 2), wt.permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    xq = x.reshape(n, h, w // G, G * cin)                                  # the merged VIEW: no data movement
    wq = torch.from_numpy(expand_weights(wt.numpy(), G))
    got = F.conv2d(xq.permute(0, 3, 1, 2), wq.permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1).reshape(n, h, w, cout)
    return float((got - want).abs().max())


if __name__ == "__main__":
    for d, steps in mma_schedule(16, 16, 4).items():
        print("group offset %+d:" % d, ", ".join("k16 step %d -> blocks [%d, %d) N = %d" % (ks, lo, lo + nb, nb * 16) for ks, lo, nb in steps))
    total = sum(len(s) for s in mma_schedule(16, 16, 4).values())
    print("MMAs per filter row and 128 pixel groups (512 pixels): %d (the per-pixel scheme issues 3 x 4 = 12 of N = 16)" % total)
    print("max |merged-view conv - plain conv| =", check())
    print("32 -> 16, G = 4:", {d: s for d, s in mma_schedule(32, 16, 4).items()})
