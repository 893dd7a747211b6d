import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
from gfnet_b200.ops import ALGO_PT, local_correlation_v2_counters
b = 64
gen = torch.Generator(device="cuda").manual_seed(0); cgen = torch.Generator().manual_seed(0)
Hs = [synth.random_homography(cgen) for _ in range(b)]
c, hs, G, r = 16, 224, 128, 2
f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, "cuda")
out = torch.empty((b, 25, G, G), device="cuda")
for t in [int(v) for v in sys.argv[1].split(",")]:
    for d in (4, 7):
        for _ in range(2):
            gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=ALGO_PT | ((t | (d << 8)) << 4), out=out)
        local_correlation_v2_counters(reset=True)
        gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=ALGO_PT | ((t | (d << 8)) << 4), out=out)
        cnt = local_correlation_v2_counters(reset=True)
        n = 296
        print(t, d, "per-CTA kclk: wait_empty %.1f wait_gready %.1f total %.1f" % (cnt[4] / n / 1e3, cnt[5] / n / 1e3, cnt[7] / n / 1e3), flush=True)
