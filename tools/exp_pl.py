"""A/B of the C = 16 local-correlation kernels: parity against the oracle at small batch, timing at op batch 64.
Usage: python tools/exp_pl.py [--tunes 0,16,17,18,19,20] [--b 64]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
import oracle
from gfnet_b200 import synth
from gfnet_b200.ops import ALGO_PT, local_correlation_v2_counters
from tools.bench_kernels import timeit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tunes", default="0,16,17,18,19,20")
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--out", default="gpurun_out/exp_pl.json")
    ap.add_argument("--jitter", type=float, default=0.15)
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    tunes = [int(v) for v in args.tunes.split(",")]
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(0)
    cgen = torch.Generator().manual_seed(0)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = []
    # parity
    for (b, c, hs, G, r) in [] if args.no_parity else [(1, 16, 224, 128, 2), (1, 16, 280, 160, 2), (3, 16, 224, 128, 2), (1, 16, 96, 96, 2), (2, 16, 112, 64, 2),
                             (1, 16, 70, 40, 2)]:
        for adv in (False, True):
            Hs = [synth.random_homography(cgen) for _ in range(b)]
            f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, dev, adversarial=adv)
            ref = oracle.local_correlation_port((b, c, hs, hs), f0.cpu(), f1.cpu(), r, G, flow=flow.cpu())
            for t in tunes:
                local_correlation_v2_counters(reset=True)
                out = gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=ALGO_PT | (t << 4))
                cnt = local_correlation_v2_counters(reset=True)
                err = float((out.cpu() - ref).abs().max() / ref.abs().max())
                row = dict(kind="parity", shape=[b, c, hs, G, r], adversarial=adv, tune=t, rel_err=err, slow_points=cnt[0])
                rows.append(row)
                print(json.dumps(row), flush=True)
                assert err < 4e-5, row
    # timing
    b = args.b
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    Hs = [synth.random_homography(cgen, jitter=args.jitter) for _ in range(b)]
    for (c, hs, G, r) in [(16, 224, 128, 2), (16, 280, 160, 2)]:
        f0, f1, flow = synth.scale_inputs(Hs, c, hs, G, gen, dev)
        out = torch.empty((b, (2 * r + 1) ** 2, G, G), device=dev)
        nbytes = gf.local_correlation_bytes(b, c, hs, hs, G, r)
        for t in tunes:
            fn = lambda: gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, algo=ALGO_PT | (t << 4), out=out)
            med, best = timeit(fn, iters=10, flush=flush)
            local_correlation_v2_counters(reset=True)
            fn()
            cnt = local_correlation_v2_counters(reset=True)
            row = dict(kind="time", shape=[b, c, hs, G, r], tune=t, ms=med, ms_best=best, GBps=nbytes / med / 1e6,
                       frac=nbytes / med / 1e6 / peak, slow_points=cnt[0])
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
