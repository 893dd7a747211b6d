import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfnet_b200 import synth
from tools.sim.sim2 import geom
from tools.sim.sim3 import run_dense
gen = torch.Generator().manual_seed(1); cgen = torch.Generator().manual_seed(1)
for jit in (0.15, 0.3):
    Hn = [synth.random_homography(cgen, jitter=jit) for _ in range(6)]
    Hs = Hn + [np.linalg.inv(h) for h in Hn]
    for (hs, G) in ((224, 128), (280, 160), (112, 64), (128, 128)):
        flow = synth.homography_flow(Hs, G, hs, gen, "cpu")
        xb, yb = geom(flow, hs, 2)
        # box extents per 16x16 tile
        ext = []
        for e in range(len(Hs)):
            for ty in range(0, G, 16):
                for tx in range(0, G, 16):
                    X = xb[e, ty:ty+16, tx:tx+16]; Y = yb[e, ty:ty+16, tx:tx+16]
                    ext.append((X.max() + 6 - (X.min() & ~3), Y.max() + 6 - Y.min()))
        ext = np.array(ext)
        print("jitter", jit, hs, G, "box ext x p50/p99/max", np.percentile(ext[:, 0], [50, 99, 100]), "y", np.percentile(ext[:, 1], [50, 99, 100]))
        for P in (56, 72):
            print("   pitch", P, run_dense(xb, yb, G, 16, 16, 2, lambda s: P, colrot=True))
