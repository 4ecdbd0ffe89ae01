"""Prototype (numpy) of the front-end's certain-background bound: counts candidates per frame and checks that no true
foreground pixel of either threshold image is ever classified as certain background."""
import sys
import numpy as np, cv2
sys.path.insert(0, '.')
from ysmr_b200 import synth

K = cv2.getGaussianKernel(11, 0, cv2.CV_32F).ravel().astype(np.float64)
# decimated minorant: one value valid for outputs p and p+1 -> weights min(k[j], k[j+1]) on the 10 common taps
H10 = np.floor(256 * np.minimum(K[:-1], K[1:])).astype(np.int64)      # taps at offsets -4..+5 relative to the even px

def bound(blurred, white_on_dark, t_star, TW=128, TH=64):
    H, W = blurred.shape
    b = blurred.astype(np.int64) if white_on_dark else 255 - blurred.astype(np.int64)
    pad = np.pad(b, 5, mode='edge')
    cand = np.zeros((H, W), bool)
    n_tiles = 0
    for ty in range(0, H, TH):
        for tx in range(0, W, TW):
            y1, x1 = min(H, ty + TH), min(W, tx + TW)
            tile = pad[ty:y1 + 10, tx:x1 + 10]                      # rows ty-5 .. y1+4
            base = tile.min(); rng = tile.max() - base
            s = 5
            while (H10.sum() * rng) >> s > 255: s += 1
            xpp = tile - base
            th, tw = tile.shape
            # row pass for even output columns x (relative col index c = x - tx, c even): taps c+5-4 .. c+5+5 in tile coords
            nxp = (x1 - tx + 1) // 2
            R = np.zeros((th, nxp), np.int64)
            for j in range(10):
                cols = np.arange(nxp) * 2 + 1 + j
                cols = np.minimum(cols, tw - 1)                       # (odd widths: harmless clamp)
                R += H10[j] * xpp[:, cols]
            rq = R >> s
            nyp = (y1 - ty + 1) // 2
            L2 = np.zeros((nyp, nxp), np.int64)
            for j in range(10):
                rows = np.minimum(np.arange(nyp) * 2 + 1 + j, th - 1)
                L2 += H10[j] * rq[rows, :]
            # certainly background iff  b - t* - 0.49 <= base + L2 * 2^s / 65536
            Lq = base + (L2 << s) / 65536.0
            Lfull = np.repeat(np.repeat(Lq, 2, 0), 2, 1)[:y1 - ty, :x1 - tx]
            bb = b[ty:y1, tx:x1]
            cand[ty:y1, tx:x1] = (bb - t_star - 0.49) > Lfull
            n_tiles += 1
    return cand

def run(name, n=3, adt=2.0, wod=True):
    cfg = synth.CONFIGS[name]
    import dataclasses
    cfg = dataclasses.replace(cfg, n_frames=n)
    sc = synth.make_scene(cfg)
    fr = synth.render_frames(sc, 0, n)
    off = 5
    for f in fr:
        bl = cv2.GaussianBlur(f, (3, 3), 0)
        so = off if wod else -off
        tt = cv2.THRESH_BINARY if wod else cv2.THRESH_BINARY_INV
        m1 = cv2.adaptiveThreshold(bl, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, tt, 11, -so)
        m2 = cv2.adaptiveThreshold(bl, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, tt, 11, -(so + adt))
        # white: d > 5, d > 7  -> t* = 5 ; dark: d <= -5 (mask), d <= -3 (marker): on complemented bytes d' = -d >= 3 -> d' > 2 -> t* = 2
        t_star = 5 if wod else 2
        c = bound(bl, wod, t_star)
        fg = (m1 > 0) | (m2 > 0)
        missed = (fg & ~c).sum()
        print(f'{name}: candidates {c.sum()} ({100 * c.mean():.3f} %), foreground {fg.sum()}, missed {missed}')
        assert missed == 0

run('cfg2'); run('cfg3'); run('cfg4', wod=False)
