"""One local_correlation shape under ncu: python tools/profile_lc.py --scale 2 --algo 146 [--pass2]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=2)
ap.add_argument("--algo", default="146")
ap.add_argument("--b", type=int, default=64)
ap.add_argument("--pass2", action="store_true")
args = ap.parse_args()
shapes = synth.pyramid_config(448, upsample_res=560) if args.pass2 else synth.pyramid_config(448)
(s, c, hs, g, r) = [x for x in shapes if x[0] == args.scale][0]
gen = torch.Generator(device="cuda").manual_seed(0)
cgen = torch.Generator().manual_seed(0)
Hs = [synth.random_homography(cgen) for _ in range(args.b)]
f0, f1, flow = synth.scale_inputs(Hs, c, hs, g, gen, "cuda")
out = torch.empty((args.b, (2 * r + 1) ** 2, g, g), device="cuda")
algos = [int(a) for a in args.algo.split(",")]
for a in algos:
    gf.local_correlation((args.b, c, hs, hs), f0, f1, r, g, flow=flow, algo=a, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for a in algos:
    gf.local_correlation((args.b, c, hs, hs), f0, f1, r, g, flow=flow, algo=a, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
