"""Ablation timings of the tc2 kernel (debug bits: 2 = no output stores, 4 = no staging stores): python tools/dbg_tc2_ablate.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
b = 64
gen = torch.Generator(device="cuda").manual_seed(0)
cgen = torch.Generator().manual_seed(0)
Hs = [synth.random_homography(cgen) for _ in range(b)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for (s, c, hs, g, r) in synth.pyramid_config(448) + synth.pyramid_config(448, upsample_res=560):
    if c < 32:
        continue
    f0, f1, flow = synth.scale_inputs(Hs, c, hs, g, gen, "cuda")
    out = torch.empty((b, (2 * r + 1) ** 2, g, g), device="cuda")
    res = []
    for dbg in (0, 128, 94, 222):
        algo = 5 | ((256 * dbg) << 4)
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gf.local_correlation((b, c, hs, hs), f0, f1, r, g, flow=flow, algo=algo, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(sorted(ts)[2])
    print(f"scale {s} hs {hs} G {g}: full {res[0]:.0f} us, full with test_wait polling {res[1]:.0f}, bare skeleton {res[2]:.0f}, bare skeleton with test_wait polling {res[3]:.0f}")
