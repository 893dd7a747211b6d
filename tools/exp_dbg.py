"""Ablation timing of lc_rot_kernel through the DEBUG entry (gfb_debug_local_corr_pt_f32; results are wrong with a switch
set): debug 1 = no window reads / FMAs, 2 = no TMA loads, 3 = both (barrier skeleton + epilogue only).
    python tools/exp_dbg.py 0,32 0,1,2,3"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfnet_b200 import synth
from gfnet_b200._lib import check, lib, ptr, stream_ptr
from tools.bench_kernels import timeit

tunes = [int(v) for v in sys.argv[1].split(",")]
dbgs = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3").split(",")]
b = 64
gen = torch.Generator(device="cuda").manual_seed(0)
cgen = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
Hs = [synth.random_homography(cgen) for _ in range(b)]
c, hs, G, r = 16, 224, 128, 2
f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, "cuda")
out = torch.empty((b, 25, G, G), device="cuda")
for t in tunes:
    for d in dbgs:
        fn = lambda: check(lib.gfb_debug_local_corr_pt_f32(ptr(f0), ptr(f1), ptr(flow), ptr(out), b, c, hs, hs, 0, G, r, 25, 0, t, d,
                                                          stream_ptr(f0.device)), "debug pt")
        med, best = timeit(fn, iters=10, flush=flush)
        print(json.dumps(dict(tune=t, dbg=d, ms=round(med, 4), best=round(best, 4))), flush=True)
