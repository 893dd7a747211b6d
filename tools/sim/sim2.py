import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfnet_b200 import synth
from collections import Counter

def geom(flow, hs, R):
    sx = ((flow[:, 0].double() + 1) * hs - 1) / 2
    sy = ((flow[:, 1].double() + 1) * hs - 1) / 2
    return (torch.floor(sx).long() - R).numpy(), (torch.floor(sy).long() - R).numpy()

def run(xb, yb, G, TY, pitch_fn, NPL=3, W=6, rowrot=True, verbose=False):
    hist = Counter(); pairs = Counter()
    for e in range(xb.shape[0]):
        for ty in range(0, G, TY):
            for tx in range(0, G, 32):
                X = xb[e, ty:ty+TY, tx:tx+32]; Y = yb[e, ty:ty+TY, tx:tx+32]
                X0 = X.min() & ~3; Y0 = Y.min()
                for w in range(X.shape[0]):
                    u = X[w] - X0; oy = Y[w] - Y0; uq, um = u // NPL, u % NPL
                    n = len(u)
                    shear = (oy[n // 2:].mean() - oy[:n // 2].mean())
                    pitch = pitch_fn(shear)
                    for j in range(W):
                        row = oy + ((j - oy) % W) if rowrot else oy + j
                        for q in range(NPL):
                            idx = uq + (q < um)
                            for s in range(2):
                                a = np.unique(row * pitch + q * 32 + idx + s)
                                bc = np.bincount(a % 32, minlength=32)
                                hist[int(bc.max())] += 1
    tot = sum(hist.values())
    return sum(k * v for k, v in hist.items()) / tot, {k: round(v / tot, 3) for k, v in sorted(hist.items())}

if __name__ == "__main__":
    gen = torch.Generator().manual_seed(0); cgen = torch.Generator().manual_seed(0)
    Hn = [synth.random_homography(cgen) for _ in range(4)]
    Hs = Hn + [np.linalg.inv(h) for h in Hn]
    for (hs, G) in ((224, 128), (280, 160)):
        flow = synth.homography_flow(Hs, G, hs, gen, "cpu")
        xb, yb = geom(flow, hs, 2)
        print(hs, G)
        for (pa, pb) in ((97, 101), (97, 127), (99, 125), (97, 97), (101, 101), (97, 103), (113, 111)):
            print(f"  rowrot pitch +{pa} / -{pb}:", run(xb, yb, G, 8, lambda sh: pa if sh >= 0 else pb))
