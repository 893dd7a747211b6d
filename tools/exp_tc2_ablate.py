"""Ablation of lc_tc2_kernel through the debug entry point (results are wrong with any bit set): bit 1 no output stores, bit 2 no
staging stores, bit 3 no B loads, bit 4 no MMAs, bit 6 pulls and hand-offs only.  Usage: python tools/exp_tc2_ablate.py"""
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
for (c, hs, G, r) in [(64, 56, 32, 6), (64, 70, 40, 6), (64, 32, 32, 7)]:
    f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, dev)
    kk = (2 * r + 1) ** 2
    out = torch.empty((b, kk, G, G), device=dev)
    nws = int(lib.gfb_local_corr_tc2_workspace_bytes(b, c, hs, hs, G, r, 0))
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    check(lib.gfb_local_corr_tc2_prepare_f32(ptr(f0), ptr(f1), b, c, hs, hs, 0, G, r, ptr(ws), nws, stream_ptr(f0.device)), "prep")
    t_run, _ = timeit(lambda: check(lib.gfb_local_corr_tc2_run_f32(ptr(f0), ptr(f1), ptr(flow), ptr(out), b, c, hs, hs, 0, G, r, kk, 0, ptr(ws), nws,
                                                                   stream_ptr(f0.device)), "run"), iters=8, flush=flush)
    print((c, hs, G, r), "plan + main on a prepared workspace: %.3f ms" % t_run, flush=True)
    for dbg, what in [(0, "full (pre-pass + plan + main)"), (2, "no output stores"), (2 | 4, "+ no staging stores"), (2 | 4 | 8, "+ no B loads"),
                      (2 | 4 | 8 | 16, "+ no MMAs"), (2 | 4 | 8 | 16 | 64, "+ pulls and hand-offs only"), (16, "no MMAs only"), (8, "no B loads only")]:
        fn = lambda: check(lib.gfb_debug_local_corr_tc2_f32(ptr(f0), ptr(f1), ptr(flow), ptr(out), b, c, hs, hs, 0, G, r, kk, 0, 0, dbg, ptr(ws), nws,
                                                            stream_ptr(f0.device)), "dbg")
        med, _ = timeit(fn, iters=8, flush=flush)
        print("   debug %3d %-32s %.3f ms" % (dbg, what, med), flush=True)
