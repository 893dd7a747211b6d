"""Per-kernel timings at the BASELINE shapes (CUDA events, inputs larger than L2 or L2 flushed).
Usage: python tools/bench_kernels.py [--b 64] [--out gpurun_out/kbench.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
from gfnet_b200 import synth
from gfnet_b200.pipeline import HotPath

PEAK_HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']


def flush_l2(buf):
    buf.zero_()


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush_l2(flush)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--out", default="gpurun_out/kbench.json")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--lc-only", action="store_true")
    ap.add_argument("--no-generic", action="store_true")
    args = ap.parse_args()
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(0)
    cgen = torch.Generator().manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    res = {"device": torch.cuda.get_device_name(0), "b": args.b, "local_corr": [], "other": {}}
    b = args.b
    Hs = [synth.random_homography(cgen) for _ in range(b)]
    shapes = synth.pyramid_config(448) + synth.pyramid_config(448, upsample_res=560)
    for (s, c, hs, g, r) in shapes:
        f0, f1, flow = synth.scale_inputs(Hs, c, hs, g, gen, dev)
        out = torch.empty((b, (2 * r + 1) ** 2, g, g), device=dev)
        nbytes = gf.local_correlation_bytes(b, c, hs, hs, g, r)
        for algo in [int(v) for v in args.variants.split(",")] + ([] if args.no_generic else [1]):
            try:
                med, best = timeit(lambda: gf.local_correlation((b, c, hs, hs), f0, f1, r, g, flow=flow, algo=algo, out=out),
                                   iters=5 if algo == 1 else 10, flush=flush)
            except NotImplementedError:
                continue
            from gfnet_b200.ops import local_correlation_v2_counters
            local_correlation_v2_counters(reset=True)
            gf.local_correlation((b, c, hs, hs), f0, f1, r, g, flow=flow, algo=algo, out=out)
            cnt = list(local_correlation_v2_counters(reset=True))
            row = dict(scale=s, c=c, hs=hs, G=g, r=r, algo=algo, ms=med, ms_best=best, GBps=nbytes / med / 1e6, counters=cnt,
                       frac=nbytes / med / 1e6 / PEAK_HBM, fma_T=(b * c * (2 * r + 2) ** 2 * g * g) / med / 1e9)
            res["local_corr"].append(row)
            print(json.dumps(row), flush=True)
        del f0, f1, flow, out
    if args.lc_only:
        json.dump(res, open(args.out, 'w'), indent=1)
        return
    # global match
    f0, f1, _ = synth.scale_inputs(Hs, 64, 32, 32, gen, dev)
    for name, kw in (("tc_3xtf32", dict(precision=0)), ("tc_tf32", dict(precision=1)), ("simt", dict(algo=1))):
        med, best = timeit(lambda: gf.coarse_match(f0, f1, **kw), flush=flush)
        fl = gf.global_match_flops(b, 64, 1024, 1024)
        res["other"]["global_match_" + name] = dict(ms=med, ms_best=best, TFLOPs=fl / med / 1e9)
    # kde
    Bp = max(1, b // 2)
    x = torch.stack([synth.make_matches(Hs[i], 20000, gen, dev) for i in range(Bp)])
    from gfnet_b200.ops import KDE_FULL, KDE_SYMMETRIC
    med, best = timeit(lambda: gf.kde(x, 0.1, half=False, algo=KDE_FULL), iters=5, flush=flush)
    res["other"]["kde_M20000_full"] = dict(ms=med, ms_best=best, pairs=Bp, Gevals_per_s=Bp * 4e8 / med / 1e6,
                                           mufu_frac=Bp * 4e8 / med / 1e6 / 4654.0)
    med, best = timeit(lambda: gf.kde(x, 0.1, half=False, algo=KDE_SYMMETRIC), iters=5, flush=flush)
    res["other"]["kde_M20000_symmetric"] = dict(ms=med, ms_best=best, pairs=Bp, Gentries_per_s=Bp * 4e8 / med / 1e6,
                                                mufu_frac_of_evaluated_half=Bp * 2e8 / med / 1e6 / 4654.0)
    x1 = x[:1].contiguous()
    med, best = timeit(lambda: gf.kde(x1, 0.1, half=False), iters=5)
    res["other"]["kde_M20000_single_pair"] = dict(ms=med, ms_best=best)
    # selection
    G = 320
    cert = torch.rand((Bp, G * 2 * G), generator=gen, device=dev) ** 4
    q = torch.empty_like(cert).exponential_(1, generator=gen)
    key = cert / q
    med, best = timeit(lambda: gf.topk_desc(key, 20000), flush=flush)
    res["other"]["topk_204800_to_20000"] = dict(ms=med, ms_best=best, pairs=Bp)
    key2 = key[:, :20000].contiguous()
    med, best = timeit(lambda: gf.topk_desc(key2, 5000))
    res["other"]["topk_20000_to_5000"] = dict(ms=med, ms_best=best, pairs=Bp)
    med, best = timeit(lambda: torch.topk(key, 20000, dim=1))
    res["other"]["torch_topk_204800_to_20000"] = dict(ms=med, ms_best=best, pairs=Bp)
    # homography
    m = torch.stack([synth.make_matches(Hs[i], 5000, gen, dev, sigma=0.001, outlier_frac=0.1) for i in range(Bp)])
    med, best = timeit(lambda: gf.estimate_homography(m, 448, 448, 448, 448))
    res["other"]["homography_cv2_faithful"] = dict(ms=med, ms_best=best, pairs=Bp)
    med, best = timeit(lambda: gf.estimate_homography(m, 448, 448, 448, 448, n_hyp=512))
    res["other"]["homography_ransac512"] = dict(ms=med, ms_best=best, pairs=Bp)
    med, best = timeit(lambda: gf.estimate_homography(m, 448, 448, 448, 448, n_hyp=0))
    res["other"]["homography_dlt_only"] = dict(ms=med, ms_best=best, pairs=Bp)
    del x, cert, q, key, m
    # whole path
    for itr in (1, 2):
        batch = synth.PairBatch(Bp, num_itr=itr, device=dev)
        hp = HotPath()
        med, best = timeit(lambda: hp.run(batch, generator=gen), iters=5, warmup=2)
        res["other"][f"hot_path_B{Bp}_itr{itr}"] = dict(ms=med, ms_best=best, pairs_per_s=Bp / med * 1e3)
        del batch, hp
    for k, v in res["other"].items():
        print(k, json.dumps(v), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
