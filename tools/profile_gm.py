"""coarse_match under ncu: python tools/profile_gm.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
gen = torch.Generator(device="cuda").manual_seed(0)
cgen = torch.Generator().manual_seed(0)
Hs = [synth.random_homography(cgen) for _ in range(64)]
f0, f1, _ = synth.scale_inputs(Hs, 64, 32, 32, gen, "cuda")
gf.coarse_match(f0, f1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
gf.coarse_match(f0, f1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
