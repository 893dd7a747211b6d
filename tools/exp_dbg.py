"""timing with parts switched off (debug bits in tune >> 8): python tools/exp_dbg.py 32,36 0,1,2,3"""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
from gfnet_b200.ops import ALGO_PT
from tools.bench_kernels import timeit
tunes = [int(v) for v in sys.argv[1].split(",")]
dbgs = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3").split(",")]
jit = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
b = 64
gen = torch.Generator(device="cuda").manual_seed(0)
cgen = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
Hs = [synth.random_homography(cgen, jitter=jit) for _ in range(b)]
for (c, hs, G, r) in [(16, 224, 128, 2)]:
    f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, "cuda")
    out = torch.empty((b, 25, G, G), device="cuda")
    for t in tunes:
        for d in dbgs:
            med, best = timeit(lambda: gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=ALGO_PT | ((t | (d << 8)) << 4), out=out), iters=10, flush=flush)
            print(json.dumps(dict(tune=t, dbg=d, ms=round(med, 4), best=round(best, 4))), flush=True)
