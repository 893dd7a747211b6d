import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfnet_b200 import synth
from tools.sim.sim2 import geom
from tools.sim.sim3 import run_dense
gen = torch.Generator().manual_seed(0); cgen = torch.Generator().manual_seed(0)
Hn = [synth.random_homography(cgen) for _ in range(4)]
Hs = Hn + [np.linalg.inv(h) for h in Hn]
hs, G = 224, 128
flow = synth.homography_flow(Hs, G, hs, gen, "cpu")
xb, yb = geom(flow, hs, 2)
for (TX, WY) in ((16, 2), (8, 4), (32, 1)):
    for P in (56, 60, 64, 68, 72, 76, 80, 84, 88, 96, 100):
        print(TX, WY, "pitch", P, "col only", run_dense(xb, yb, G, TX, 8, WY, lambda s: P, rowrot=False, colrot=True)[0],
              "both", run_dense(xb, yb, G, TX, 8, WY, lambda s: P, colrot=True)[0])
