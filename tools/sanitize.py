"""One small launch of every hot kernel (for compute-sanitizer): the rotated-order point kernel, the tcgen05 kernel (+ pre-pass),
the mma.sync kernel, the refiner assemble kernel, the tcgen05 global match, the symmetric kde, top-k, the homography solver,
the refiner convolution tail.
Usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize.py [names...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
from gfnet_b200.ops import ALGO_MMA, ALGO_PT, ALGO_TC2

dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(0)
cgen = torch.Generator().manual_seed(0)
Hs = [synth.random_homography(cgen) for _ in range(2)]
want = set(sys.argv[1:])


def run(name, fn):
    if want and name not in want:
        return
    fn()
    torch.cuda.synchronize()
    print("ran", name, flush=True)


def lc(c, hs, G, r, algo):
    f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, dev)
    return lambda: gf.local_correlation((2, c, hs, hs), f0, f1, r, G, flow=flow, algo=algo)


run("lc_rot", lc(16, 112, 64, 2, ALGO_PT))
run("lc_tc2", lc(64, 56, 32, 6, ALGO_TC2))
run("lc_tc2_c32", lc(32, 56, 32, 4, ALGO_TC2))
run("lc_mma", lc(32, 56, 32, 4, ALGO_MMA))
run("lc_mma_c64", lc(64, 32, 32, 7, ALGO_MMA))
x = torch.randn((2, 32, 56, 56), generator=gen, device=dev)
y = torch.randn((2, 32, 56, 56), generator=gen, device=dev)
fl = synth.homography_flow(Hs, 32, 56, gen, dev)
w, b = torch.randn((32, 2), generator=gen, device=dev), torch.randn((32,), generator=gen, device=dev)
run("refiner_input", lambda: gf.refiner_input(32, x, y, fl, w, b, 4))
f0 = torch.randn((2, 64, 32, 32), generator=gen, device=dev)
f1 = torch.randn((2, 64, 32, 32), generator=gen, device=dev)
run("gm_tc", lambda: gf.coarse_match(f0, f1))
pts = torch.stack([synth.make_matches(Hs[i], 4096, gen, dev) for i in range(2)])
run("kde4_sym", lambda: gf.kde(pts, 0.1, half=False))
key = torch.rand((2, 20000), generator=gen, device=dev)
run("topk", lambda: gf.topk_desc(key, 5000))
run("homography", lambda: gf.estimate_homography(pts, 448, 448, 448, 448))
# refiner convolution tail: pack, depth-wise, 1x1 GEMM (tcgen05 at C = 177, mma.sync at C = 24 / 73), out_conv, glue
from gfnet_b200 import refiner as RF


def tail(c, G, algo=0):
    torch.manual_seed(c)
    blocks = [torch.nn.Sequential(torch.nn.Conv2d(c, c, 5, 1, 2, groups=c), torch.nn.BatchNorm2d(c), torch.nn.ReLU(inplace=True),
                                  torch.nn.Conv2d(c, c, 1, 1, 0)).cuda().eval() for _ in range(2)]
    rb = RF.RefinerBlocks(blocks, torch.nn.Conv2d(c, 3, 1, 1, 0).cuda().eval())
    d = torch.randn((2, c, G, G), generator=gen, device=dev)
    return lambda: rb(d, algo=algo)


run("rb_tail_c177", tail(177, 24))
run("rb_tail_c73", tail(73, 40))
run("rb_tail_c24", tail(24, 48))
run("rb_tail_c24_tc", tail(24, 48, algo=2))
dl = torch.randn((2, 3, 32, 32), generator=gen, device=dev)
fw, ce = torch.rand((2, 2, 32, 32), generator=gen, device=dev), torch.zeros((2, 1, 32, 32), device=dev)
pre = torch.zeros_like(fw) + 1e-7
run("flow_update", lambda: RF.flow_update(dl, fw, ce, pre, 8, 448, 448))
run("upsample", lambda: RF.upsample_bilinear(fw, 64))
print("done")
