"""Ablation of lc_rot_kernel through the debug entry point (results are wrong with any bit set): bit 0 = no window reads / FMAs,
bit 1 = no TMA loads.  Usage: python tools/exp_rot_ablate.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfnet_b200 import synth
from gfnet_b200._lib import lib, ptr, stream_ptr, check
from tools.bench_kernels import timeit

dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(0)
cgen = torch.Generator().manual_seed(0)
b = 64
Hs = [synth.random_homography(cgen) for _ in range(b)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for (c, hs, G, r) in [(16, 224, 128, 2), (16, 280, 160, 2)]:
    f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, dev)
    out = torch.empty((b, 25, G, G), device=dev)
    for tune in (0, 32):
        for dbg in (0, 1, 2, 3):
            fn = lambda: check(lib.gfb_debug_local_corr_pt_f32(ptr(f0), ptr(f1), ptr(flow), ptr(out), b, c, hs, hs, 0, G, r, 25, 0, tune, dbg,
                                                               stream_ptr(f0.device)), "dbg")
            med, best = timeit(fn, iters=8, flush=flush)
            print((c, hs, G, r), "tune", tune, "debug", dbg, "ms %.3f" % med, flush=True)
