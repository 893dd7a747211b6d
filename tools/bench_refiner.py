"""Kernel bench of the refiner convolution tail at the pipeline's shapes: ms per call, useful TFLOP/s,
fraction of the measured bf16 peak, the HBM floor of the fused formulation, torch's own modules (cuDNN / cuBLAS under fp16
autocast = what the reference runs on a GPU) on the same device and the same modules on the box's host cores (fp32 = the
reference's CPU path, one batch element) beside it.

    python tools/bench_refiner.py [--b 64] [--out gpurun_out/refiner_bench.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfnet_b200 import refiner as RF  # noqa: E402

# (scale, C = 2c + dd + (2r+1)^2, G) of model/network.py:79-153 with vis_ir.json; 448 pass then 560 pass
SHAPES = [("p1_s16", 417, 32), ("p1_s8", 361, 32), ("p1_s4", 177, 64), ("p1_s2", 73, 128), ("p1_s1", 24, 256),
          ("p2_s8", 361, 40), ("p2_s4", 177, 80), ("p2_s2", 73, 160), ("p2_s1", 24, 320)]


def make(c):
    torch.manual_seed(c)
    blocks = [torch.nn.Sequential(torch.nn.Conv2d(c, c, 5, 1, 2, groups=c), torch.nn.BatchNorm2d(c), torch.nn.ReLU(inplace=True),
                                  torch.nn.Conv2d(c, c, 1, 1, 0)).cuda().eval() for _ in range(9)]
    return blocks, torch.nn.Conv2d(c, 3, 1, 1, 0).cuda().eval()


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/refiner_bench.json")
    ap.add_argument("--torch", type=int, default=1)
    ap.add_argument("--cpu", type=int, default=1)
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    rows, tot = [], {"ms": 0.0, "flops": 0.0, "torch_ms": 0.0}
    for name, c, G in SHAPES:
        blocks, oc = make(c)
        rb = RF.RefinerBlocks(blocks, oc)
        d = torch.randn(a.b, c, G, G, device="cuda")
        ms = timed(lambda: rb(d), a.iters)
        flops = RF.refiner_blocks_flops(a.b, c, G)
        cp = RF.pad16(c)
        h = RF.pack_nhwc_f16(d[:8])
        wf, shift, b2, w2 = rb.folded[0]
        ms_dw = timed(lambda: RF.dw5_bn_relu(h, wf, shift, G), 10)
        ms_pw = timed(lambda: RF.pointwise(h, w2, b2), 10)
        row = {"shape": name, "C": c, "Cp": cp, "G": G, "b": a.b, "ms": ms, "tflops": flops / ms * 1e-9,
               "frac_bf16_peak": flops / ms * 1e-9 / peaks["bf16_tflops_sustained"],
               "hbm_floor_ms": RF.refiner_blocks_bytes(a.b, c, G) / peaks["hbm_gbs"] * 1e-6,
               "launches": rb.launches(a.b, G), "chunk": int(RF.lib.gfb_refiner_blocks_chunk(a.b, c, G)),
               "dw_ms_b8": ms_dw, "dw_GBps_l2": 2 * 8 * G * G * cp * 2 / ms_dw * 1e-6,
               "pw_ms_b8": ms_pw, "pw_tflops": 2 * 8 * G * G * cp * cp / ms_pw * 1e-9}
        if a.cpu:      # the reference's CPU path (fp32: its autocast is off on the CPU, utils/utils.py:306-320), one batch element
            import copy
            import time
            cseq, coc = copy.deepcopy(torch.nn.Sequential(*blocks)).cpu(), copy.deepcopy(oc).cpu()
            dc = d[:1].cpu()
            with torch.no_grad():
                coc(cseq(dc.clone()))
                t0 = time.perf_counter()
                coc(cseq(dc.clone()))
                row["cpu_fp32_ms_b1"] = (time.perf_counter() - t0) * 1e3
            row["cpu_threads"] = torch.get_num_threads()
            row["gpu_over_cpu_per_element"] = row["cpu_fp32_ms_b1"] / (ms / a.b)
        if a.torch:
            seq = torch.nn.Sequential(*blocks)

            def ref():
                with torch.no_grad():
                    with torch.autocast("cuda", dtype=torch.float16):
                        hh = seq(d.clone())
                    return oc(hh.float())
            row["torch_autocast_ms"] = timed(ref, 3)
            tot["torch_ms"] += row["torch_autocast_ms"]
        tot["ms"] += ms
        tot["flops"] += flops
        rows.append(row)
        print(json.dumps(row), flush=True)
        del rb, d, blocks
        torch.cuda.empty_cache()
    summary = {"what": "refiner convolution tail, one call per shape (a pipeline step runs each num_itr = 2 times)", "rows": rows,
               "sum_ms": tot["ms"], "sum_tflops": tot["flops"] / tot["ms"] * 1e-9, "sum_torch_autocast_ms": tot["torch_ms"],
               "peaks": {"bf16_tflops_sustained": peaks["bf16_tflops_sustained"], "hbm_gbs": peaks["hbm_gbs"]}}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(summary, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "rows"}))


if __name__ == "__main__":
    main()
