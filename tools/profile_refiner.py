"""One refiner-tail call of a pipeline shape inside cudaProfilerStart/Stop (for ncu --profile-from-start off), or a chunk sweep.

    ncu --set full -k regex:rb_dw --launch-count 1 ... python tools/profile_refiner.py --shape p1_s1 --b 16
    python tools/profile_refiner.py --sweep 1
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfnet_b200 import refiner as RF  # noqa: E402
from tools.bench_refiner import SHAPES, make, timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="p1_s1")
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--sweep", type=int, default=0)
    a = ap.parse_args()
    shapes = {n: (c, G) for n, c, G in SHAPES}
    if a.sweep:
        for name in ("p1_s16", "p1_s4", "p1_s2", "p1_s1", "p2_s1"):
            c, G = shapes[name]
            blocks, oc = make(c)
            rb = RF.RefinerBlocks(blocks, oc)
            d = torch.randn(a.b, c, G, G, device="cuda")
            auto = int(RF.lib.gfb_refiner_blocks_chunk(a.b, c, G))
            row = {"shape": name, "auto_chunk": auto}
            for ch in sorted({auto, min(a.b, 2 * auto), min(a.b, 4 * auto), a.b}):
                row[f"chunk{ch}_ms"] = timed(lambda: rb(d, chunk=ch), 3)
            print(json.dumps(row), flush=True)
        return
    c, G = shapes[a.shape]
    blocks, oc = make(c)
    rb = RF.RefinerBlocks(blocks, oc)
    d = torch.randn(a.b, c, G, G, device="cuda")
    rb(d)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    rb(d)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
