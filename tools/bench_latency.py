"""Single-pair latency of the refinement loop (the reference's own operating point, test.py: one pair per call): decoder.refine
eager (kernels launched from Python) against decoder.GraphedRefine (one CUDA-graph replay), 448 pass + 560 upsample pass,
random pyramids and random-init refiners; outputs must be identical.

    python tools/bench_latency.py [--pairs 1]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfnet_b200 import decoder  # noqa: E402

CH = {"16": 64, "8": 64, "4": 32, "2": 16, "1": 8}
RAD = {"16": 7, "8": 6, "4": 4, "2": 2, "1": 0}


class Refiner(torch.nn.Module):                      # the structure and widths of model/network.py:76-155, 444-531
    def __init__(self, s):
        super().__init__()
        c, kk = CH[s], (2 * RAD[s] + 1) ** 2 if s != "1" else 0
        dim = 2 * c + CH[s] + kk
        mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(dim, dim, 5, 1, 2, groups=dim), torch.nn.BatchNorm2d(dim),
                                         torch.nn.ReLU(inplace=True), torch.nn.Conv2d(dim, dim, 1, 1, 0))
        self.block1, self.hidden_blocks = mk(), torch.nn.Sequential(*[mk() for _ in range(8)])
        self.out_conv, self.disp_emb = torch.nn.Conv2d(dim, 3, 1, 1, 0), torch.nn.Conv2d(2, CH[s], 1, 1, 0)
        self.local_corr_radius, self.corr_in_other, self.amp, self.amp_dtype = RAD[s], s != "1", True, torch.float16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/r2_latency.json")
    a = ap.parse_args()
    dev, b = "cuda", 2 * a.pairs
    torch.manual_seed(0)
    refiners = torch.nn.ModuleDict({s: Refiner(s) for s in CH}).to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(3)

    def pyramid(r, scales):
        size = {"16": r // 14, "8": r // 8, "4": r // 4, "2": r // 2, "1": r}
        mk = lambda: {s: torch.randn((b, CH[s], size[s], size[s]), generator=gen, device=dev) for s in scales}
        return mk(), mk()

    f0a, f1a = pyramid(448, ("16", "8", "4", "2", "1"))
    f0b, f1b = pyramid(560, ("8", "4", "2", "1"))
    ga, gb = [32, 32, 64, 128, 256], [40, 80, 160, 320]

    def eager():
        c = decoder.refine(f0a, f1a, refiners, ga, [2] * 5, 448, 448)
        return decoder.refine(f0b, f1b, refiners, gb, [2] * 4, 560, 560, scale_factor=1.25, upsample=True, pre_corresps=c["1"][2])

    with torch.inference_mode():
        ref = eager()
        torch.cuda.synchronize()
        g1 = decoder.GraphedRefine(f0a, f1a, refiners, ga, [2] * 5, 448, 448)
        c1 = g1(f0a, f1a)
        g2 = decoder.GraphedRefine(f0b, f1b, refiners, gb, [2] * 4, 560, 560, scale_factor=1.25, upsample=True, pre_corresps=c1["1"][2])

        def graphed():
            c = g1(f0a, f1a)
            return g2(f0b, f1b, c["1"][2])

        out = graphed()
        torch.cuda.synchronize()
        same = all(torch.equal(ref[s][it][k], out[s][it][k]) for s in ref for it in ref[s] for k in ("flow", "certainty"))

        def wall(fn, n):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n * 1e3
        row = {"pairs": a.pairs, "eager_ms": wall(eager, 10), "graph_ms": wall(graphed, 10), "identical": bool(same)}
    row["speedup"] = row["eager_ms"] / row["graph_ms"]
    print(json.dumps(row))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(row, open(a.out, "w"))


if __name__ == "__main__":
    main()
