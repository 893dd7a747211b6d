"""Experiment: the depth-wise 5x5 stage as banded MMAs on channel-planar fp16 (rb_dwm_kernel) against the
shipped NHWC FFMA2 kernel (rb_dw_kernel) on the refiner shapes, op batch 64: time and error against float64.

    python tools/exp_dw_planar.py
"""
import ctypes
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfnet_b200 import refiner as RF                      # noqa: E402
from gfnet_b200._lib import lib, check, ptr, stream_ptr   # noqa: E402
from tools.bench_refiner import SHAPES, timed             # noqa: E402

b = 64
rows = []
for name, c, G in SHAPES:
    cp = RF.pad16(c)
    g = torch.Generator(device="cuda").manual_seed(c)
    w = torch.randn((c, 25), generator=g, device="cuda") * 0.2
    wf = torch.zeros((25, cp), device="cuda"); wf[:, :c] = w.t()
    shift = torch.zeros(cp, device="cuda"); shift[:c] = torch.randn(c, generator=g, device="cuda") * 0.1
    d = torch.randn((b, c, G, G), generator=g, device="cuda")
    h = RF.pack_nhwc_f16(d)                                                   # [b, P, cp]
    planar = d.half().contiguous()                                            # [b, c, G, G]
    out_p = torch.empty_like(planar)

    def run_planar():
        check(lib.gfb_debug_refiner_dw5_planar_f16(ptr(planar), ptr(wf), ptr(shift), ptr(out_p), b, c, cp, G, stream_ptr("cuda")), "dw planar")
    run_planar()
    o_nhwc = RF.dw5_bn_relu(h, wf, shift, G)[..., :c].reshape(b, G, G, c).permute(0, 3, 1, 2)
    n_chk = min(b, 4)
    ref = F.relu(F.conv2d(planar[:n_chk].double(), w.double().reshape(c, 1, 5, 5), shift[:c].double(), 1, 2, 1, c))
    mx = float(ref.abs().max())
    e_planar = float((out_p[:n_chk].double() - ref).abs().max()) / mx
    e_nhwc = float((o_nhwc[:n_chk].double() - ref).abs().max()) / mx
    ms_planar = timed(run_planar, 10)
    ms_nhwc = timed(lambda: RF.dw5_bn_relu(h, wf, shift, G), 10)
    row = dict(shape=name, C=c, G=G, b=b, ms_planar_mma=ms_planar, ms_nhwc_ffma2=ms_nhwc, speedup=ms_nhwc / ms_planar,
               err_planar=e_planar, err_nhwc=e_nhwc, planar_GBps=2 * b * c * G * G * 2 / ms_planar * 1e-6)
    rows.append(row)
    print(json.dumps(row), flush=True)
    del d, h, planar, out_p, o_nhwc
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/r2_exp_dw_planar.json", "w"), indent=1)
