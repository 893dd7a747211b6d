"""Debug aid: one small call of a v2 local-correlation kernel. python tools/dbg_v2.py pt|tc2 [scale]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
import oracle
which = sys.argv[1]
shape = {"pt": (1, 16, 224, 128, 2), "tc2": (2, 64, 32, 32, 7), "tc2s4": (1, 32, 112, 64, 4), "tc2s8": (2, 64, 56, 32, 6)}[which]
b, c, hs, G, r = shape
gen = torch.Generator(device="cuda").manual_seed(5)
cgen = torch.Generator().manual_seed(13)
Hs = [synth.random_homography(cgen) for _ in range(b)]
f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, "cuda")
algo = 4 if which == "pt" else 5
out = gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=algo)
torch.cuda.synchronize()
ref = oracle.local_correlation_port((b, c, hs, hs), f0.cpu(), f1.cpu(), r, G, flow=flow.cpu())
d = (out.cpu() - ref).abs()
print(which, "max abs diff", float(d.max()), "ref max", float(ref.abs().max()), "bad frac", float((d > 1e-4 * ref.abs() + 4e-5 * ref.abs().max()).float().mean()))
