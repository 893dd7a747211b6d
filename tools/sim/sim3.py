import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfnet_b200 import synth
from collections import Counter
from tools.sim.sim2 import geom

def run_dense(xb, yb, G, TX, TY, WY, pitch_fn, W=6, rowrot=True, colrot=False):
    """dense layout [row][x], LDS.32, warp = TX x WY points (TX*WY = 32), tile TX x TY; pitch in words"""
    hist = Counter()
    for e in range(xb.shape[0]):
        for ty in range(0, G, TY):
            for tx in range(0, G, TX):
                X = xb[e, ty:ty+TY, tx:tx+TX]; Y = yb[e, ty:ty+TY, tx:tx+TX]
                X0 = X.min() & ~3; Y0 = Y.min()
                for w0 in range(0, X.shape[0], WY):
                    u = X[w0:w0+WY].ravel() - X0; oy = Y[w0:w0+WY].ravel() - Y0
                    n = X.shape[1]
                    shear = (Y[w0:w0+WY, n // 2:].mean() - Y[w0:w0+WY, :n // 2].mean())
                    pitch = pitch_fn(shear)
                    for j in range(W):
                        row = oy + ((j - oy) % W) if rowrot else oy + j
                        for i in range(W):
                            col = u + ((i - u) % W) if colrot else u + i
                            a = np.unique(row * pitch + col)
                            bc = np.bincount(a % 32, minlength=32)
                            hist[int(bc.max())] += 1
    tot = sum(hist.values())
    return round(sum(k * v for k, v in hist.items()) / tot, 3), {k: round(v / tot, 3) for k, v in sorted(hist.items())}

if __name__ == "__main__":
    gen = torch.Generator().manual_seed(0); cgen = torch.Generator().manual_seed(0)
    Hn = [synth.random_homography(cgen) for _ in range(4)]
    Hs = Hn + [np.linalg.inv(h) for h in Hn]
    hs, G = 224, 128
    flow = synth.homography_flow(Hs, G, hs, gen, "cpu")
    xb, yb = geom(flow, hs, 2)
    for (TX, WY) in ((16, 2), (32, 1), (8, 4)):
        for P in (64, 65, 67, 69, 71, 73, 75, 77, 79):
            print(TX, WY, "pitch", P, "rowrot", run_dense(xb, yb, G, TX, 8, WY, lambda s: P), "both", run_dense(xb, yb, G, TX, 8, WY, lambda s: P, colrot=True))
        print(TX, WY, "pitch 64 no rot", run_dense(xb, yb, G, TX, 8, WY, lambda s: 64, rowrot=False))
