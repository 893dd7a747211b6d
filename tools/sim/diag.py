import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfnet_b200 import synth
from tools.sim.bank_sim import geom
from collections import Counter
gen = torch.Generator().manual_seed(0); cgen = torch.Generator().manual_seed(0)
Hn = [synth.random_homography(cgen) for _ in range(4)]
Hs = Hn + [np.linalg.inv(h) for h in Hn]
flow = synth.homography_flow(Hs, 128, 224, gen, "cpu")
xb, yb = geom(flow, 224, 2)
pitch = 100; NPL = 3; W = 6; TY = 8
pairs = Counter(); hist = Counter(); spans = []
for e in range(len(Hs)):
    for ty in range(0, 128, TY):
        for tx in range(0, 128, 32):
            X = xb[e, ty:ty+TY, tx:tx+32]; Y = yb[e, ty:ty+TY, tx:tx+32]
            X0 = X.min() & ~3; Y0 = Y.min()
            for w in range(TY):
                u = X[w] - X0; oy = Y[w] - Y0; uq, um = u // NPL, u % NPL
                spans.append(uq.max() - uq.min())
                for j in range(W):
                    row = oy + ((j - oy) % W)
                    for q in range(NPL):
                        idx = uq + (q < um)
                        a = row * pitch + q * 32 + idx
                        ua = np.unique(a)
                        bc = np.bincount(ua % 32, minlength=32)
                        hist[bc.max()] += 1
                        if bc.max() > 1:
                            bk = bc.argmax()
                            cells = sorted({(int(i), int(r)) for i, r in zip(idx, row) if (r * pitch + q*32 + i) % 32 == bk})
                            pairs[(cells[1][0] - cells[0][0], cells[1][1] - cells[0][1])] += 1
print(hist); print(pairs.most_common(12)); print("idx span mean/max", np.mean(spans), np.max(spans), np.percentile(spans, [50, 90, 99]))
