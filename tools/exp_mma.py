"""The mma.sync local-correlation kernel: parity against the oracle at small batch (regular + adversarial flows), timing of
every CTA shape at op batch 64 on the seven shapes of a bench step, next to the default kernels.
Usage: python tools/exp_mma.py [--b 64] [--shapes 0,36,20,18] [--no-parity]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
import oracle
from gfnet_b200 import synth
from gfnet_b200.ops import ALGO_MMA, local_correlation_mma_counters
from tools.bench_kernels import timeit

PYR = [(64, 32, 32, 7), (64, 56, 32, 6), (32, 112, 64, 4), (16, 224, 128, 2), (64, 70, 40, 6), (32, 140, 80, 4), (16, 280, 160, 2)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="0,36,20,18")          # 4|2<<4 = 36, 4|1<<4 = 20, 2|1<<4 = 18
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--out", default="gpurun_out/exp_mma.json")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-default", action="store_true")
    args = ap.parse_args()
    shapes = [int(v) for v in args.shapes.split(",")]
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(0)
    cgen = torch.Generator().manual_seed(0)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = []
    if not args.no_parity:
        cases = [(2, c, hs, G, r) for (c, hs, G, r) in PYR] + [(1, 16, 96, 96, 2), (3, 32, 50, 20, 4), (1, 64, 36, 36, 3),
                                                              (1, 128, 28, 28, 5), (2, 32, 30, 44, 2), (1, 16, 20, 12, 1)]
        for (b, c, hs, G, r) in cases:
            for adv in (False, True):
                Hs = [synth.random_homography(cgen) for _ in range(b)]
                f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, dev, adversarial=adv)
                ref = oracle.local_correlation_port((b, c, hs, hs), f0.cpu(), f1.cpu(), r, G, flow=flow.cpu())
                for sh in shapes:
                    if sh == 36 and c > 32:
                        continue
                    local_correlation_mma_counters(reset=True)
                    try:
                        out = gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=ALGO_MMA | (sh << 4))
                    except NotImplementedError:
                        continue
                    cnt = local_correlation_mma_counters(reset=True)
                    err = float((out.cpu() - ref).abs().max() / ref.abs().max())
                    row = dict(kind="parity", shape=[b, c, hs, G, r], adversarial=adv, cta=sh, rel_err=err, slow_points=cnt[0])
                    rows.append(row)
                    print(json.dumps(row), flush=True)
    b = args.b
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    Hs = [synth.random_homography(cgen) for _ in range(b)]
    for (c, hs, G, r) in PYR:
        f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, dev)
        out = torch.empty((b, (2 * r + 1) ** 2, G, G), device=dev)
        nbytes = gf.local_correlation_bytes(b, c, hs, hs, G, r)
        variants = [("mma", sh) for sh in shapes if not (sh == 36 and c > 32)] + ([] if args.no_default else [("default", 0)])
        for kind, sh in variants:
            algo = (ALGO_MMA | (sh << 4)) if kind == "mma" else 0
            fn = lambda: gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=algo, out=out)
            med, best = timeit(fn, iters=10, flush=flush)
            local_correlation_mma_counters(reset=True)
            fn()
            cnt = local_correlation_mma_counters(reset=True)
            row = dict(kind="time", shape=[b, c, hs, G, r], kernel=kind, cta=sh, ms=med, ms_best=best, GBps=nbytes / med / 1e6,
                       frac=nbytes / med / 1e6 / peak, slow_points=cnt[0])
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
