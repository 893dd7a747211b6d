"""kde under ncu: python tools/profile_kde.py [--pairs 32] [--algo 2]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=32)
ap.add_argument("--algo", type=int, default=2)
args = ap.parse_args()
gen = torch.Generator(device="cuda").manual_seed(0)
cgen = torch.Generator().manual_seed(0)
x = torch.stack([synth.make_matches(synth.random_homography(cgen), 20000, gen, "cuda") for _ in range(args.pairs)])
gf.kde(x, 0.1, half=False, algo=args.algo)
torch.cuda.synchronize()
torch.cuda.profiler.start()
gf.kde(x, 0.1, half=False, algo=args.algo)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
