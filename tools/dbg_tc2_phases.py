"""Per-phase clocks of the tc2 epilogue (debug counters): python tools/dbg_tc2_phases.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
from gfnet_b200.ops import local_correlation_v2_counters
b = 64
gen = torch.Generator(device="cuda").manual_seed(0)
cgen = torch.Generator().manual_seed(0)
Hs = [synth.random_homography(cgen) for _ in range(b)]
for (s, c, hs, g, r) in synth.pyramid_config(448) + synth.pyramid_config(448, upsample_res=560):
    if c < 32:
        continue
    f0, f1, flow = synth.scale_inputs(Hs, c, hs, g, gen, "cuda")
    out = torch.empty((b, (2 * r + 1) ** 2, g, g), device="cuda")
    for dbg in (1, 3, 5, 7):
        algo = 5 | ((256 * dbg) << 4)
        gf.local_correlation((b, c, hs, hs), f0, f1, r, g, flow=flow, algo=algo, out=out)
        local_correlation_v2_counters(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gf.local_correlation((b, c, hs, hs), f0, f1, r, g, flow=flow, algo=algo, out=out)
        e1.record()
        cnt = local_correlation_v2_counters(reset=True)
        n = max(cnt[7], 1)
        print(f"scale {s} hs {hs} G {g} dbg {dbg}: {e0.elapsed_time(e1)*1e3:.0f} us; clocks per chunk: wait {cnt[4]/n:.0f} pull {cnt[5]/n:.0f} rows {cnt[6]/n:.0f}; gather pts {cnt[1]} tiles {cnt[2]}")
