"""BASELINE.json config 5: local_correlation microbench sweep -- r in {2,3,4,6,7,8} x C in {64,128,256,512} x hs in
{28,32,56,84} x G in {hs/2, hs/1.75, hs} -- against the reference's CPU function.

    python tools/sweep.py [--out gpurun_out/r2_kbench_sweep.json] [--cpu-seconds 150]

Per shape: op batch b so that inputs + outputs exceed the 126 MB L2 (and L2 flushed between iterations), median of 5
CUDA-event timings, algorithmic GB/s and fraction of the measured HBM peak; for as many shapes as the CPU budget allows
(smallest first) the reference's own function (oracle.reference, else the oracle port) is timed at b = 2 on the host
cores and its output compared with ours (max |diff| / max |ref|).  Also times kde and the homography solver over sizes.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfnet_b200 as gf
import oracle
from gfnet_b200 import synth
from oracle import reference as R
from tools.bench_kernels import timeit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r2_kbench_sweep.json")
    ap.add_argument("--cpu-seconds", type=float, default=150.0)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = "cuda"
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    gen = torch.Generator(device=dev).manual_seed(0)
    cgen = torch.Generator().manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ref_lc = R.load_reference().local_correlation if R.available() else oracle.local_correlation_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    radii, chans, sizes = (2, 3, 4, 6, 7, 8), (64, 128, 256, 512), (28, 32, 56, 84)
    if args.quick:
        radii, chans, sizes = (2, 7), (64, 256), (32, 56)
    shapes = []
    for hs in sizes:
        for G in sorted({hs // 2, int(hs / 1.75), hs}):
            for c in chans:
                for r in radii:
                    shapes.append((c, hs, G, r))
    rows = []
    Hs_all = [synth.random_homography(cgen) for _ in range(4096)]
    for (c, hs, G, r) in shapes:
        kk = (2 * r + 1) ** 2
        per = 4 * (c * G * G + c * hs * hs + 2 * G * G + kk * G * G)
        b = int(min(max(2, -(-200e6 // per)), (2 ** 31 - 1) // (kk * G * G), 4096))
        f0, f1, flow = synth.scale_inputs(Hs_all[:b], c, hs, G, gen, dev)
        out = torch.empty((b, kk, G, G), device=dev)
        med, best = timeit(lambda: gf.local_correlation((b, c, hs, hs), f0, f1, r, G, flow=flow, out=out), iters=5, warmup=2, flush=flush)
        nbytes = gf.local_correlation_bytes(b, c, hs, hs, G, r)
        rows.append(dict(c=c, hs=hs, G=G, r=r, b=b, ms=med, GBps=nbytes / med / 1e6, frac_hbm=nbytes / med / 1e6 / peak,
                         temp_mb_ref=c * G * G * kk * 4 / 1e6, cpu_ms_b2=None, rel_err_vs_cpu=None))
        print(json.dumps(rows[-1]), flush=True)
        del f0, f1, flow, out
    # reference CPU function at b = 2, cheapest shapes first, within the time budget
    t_start = time.perf_counter()
    order = sorted(range(len(rows)), key=lambda i: rows[i]["temp_mb_ref"] * (1 + rows[i]["hs"] / 84))
    ncpu = 0
    for i in order:
        if time.perf_counter() - t_start > args.cpu_seconds:
            break
        rw = rows[i]
        c, hs, G, r = rw["c"], rw["hs"], rw["G"], rw["r"]
        if rw["temp_mb_ref"] * 2 > 3000:
            continue
        f0, f1, flow = synth.scale_inputs(Hs_all[:2], c, hs, G, gen, dev)
        ours = gf.local_correlation((2, c, hs, hs), f0, f1, r, G, flow=flow).cpu()
        a0, a1, fl = f0.cpu(), f1.cpu(), flow.cpu()
        with torch.no_grad():
            ref_lc((2, c, hs, hs), a0, a1, r, G, flow=fl)
            t0 = time.perf_counter()
            ref = ref_lc((2, c, hs, hs), a0, a1, r, G, flow=fl)
            rw["cpu_ms_b2"] = (time.perf_counter() - t0) * 1e3
        rw["rel_err_vs_cpu"] = float((ours - ref).abs().max() / ref.abs().max())
        rw["gpu_over_cpu_per_element"] = (rw["cpu_ms_b2"] / 2) / (rw["ms"] / rw["b"])
        ncpu += 1
    res = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak, "cpu_cores": cores,
           "cpu_function": "reference utils.local_correlation.local_correlation" if R.available() else "oracle port",
           "cpu_shapes_timed": ncpu, "local_correlation": rows, "other": {}}
    worst = max((rw["rel_err_vs_cpu"] or 0.0) for rw in rows)
    print("shapes", len(rows), "cpu-timed", ncpu, "worst rel err vs reference CPU", worst, flush=True)
    # kde and homography over sizes
    for M in (5000, 20000, 40000):
        x = torch.stack([synth.make_matches(Hs_all[i], M, gen, dev) for i in range(8)])
        med, _ = timeit(lambda: gf.kde(x, 0.1, half=False), iters=5)
        t0 = time.perf_counter(); oracle.kde_port(x[0].cpu(), 0.1, half=False, down=8); tc = time.perf_counter() - t0
        res["other"][f"kde_M{M}"] = dict(pairs=8, ms=med, Gevals_per_s=8 * M * M / med / 1e6, cpu_ms_one_pair_down8=tc * 1e3)
    for (sigma, outl) in ((0.002, 0.0), (0.002, 0.2), (0.002, 0.5)):
        m = torch.stack([synth.make_matches(Hs_all[i], 5000, gen, dev, sigma=sigma, outlier_frac=outl) for i in range(32)])
        med, _ = timeit(lambda: gf.estimate_homography(m, 448, 448, 448, 448), iters=5)
        import numpy as np
        pa, pb = oracle.convert_coordinates(m[0, :, :2].cpu().numpy(), m[0, :, 2:].cpu().numpy(), 448, 448, 448, 448)
        t0 = time.perf_counter(); oracle.find_homography_cv2(pa, pb); tc = time.perf_counter() - t0
        res["other"][f"homography_outliers{int(outl * 100)}"] = dict(pairs=32, ms=med, cpu_cv2_ms_one_pair=tc * 1e3)
    for k, v in res["other"].items():
        print(k, json.dumps(v), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
