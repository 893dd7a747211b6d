"""CPU hot path for one batch of pairs, in the reference's call order.
TEST INFRASTRUCTURE / CPU BASELINE ONLY (tests, bench.py ``cpu_baseline`` and ``--impl reference``).

``impl="reference"`` calls the REFERENCE'S OWN functions (oracle.reference: ``local_correlation``,
``GFNet.corr_volume`` / ``pos_embed`` / ``sample`` -> ``kde``) plus ``cv2.findHomography`` exactly as estimation.py:66-72
does; the two stretches of the reference that cannot be imported (the tail of ``GFNet.match``, model/network.py:358-384,
which is inline in a method that needs the backbone, and estimation.py, which imports kornia) are the oracle ports.
``impl="port"`` uses the oracle ports throughout (same torch / cv2 operators in the same order).
With ``impl="reference"`` the batch may live on a CUDA device: the same reference functions then run as the reference runs
them on a GPU (torch CUDA kernels, fp16 kde at ``down=1`` -- model/network.py:405-408), cv2 on the host as in estimation.py.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from .coarse_match import corr_volume_port, pos_embed_port
from .local_correlation import local_correlation_port
from .sampling import match_postprocess_port, sample_port
from .estimation import convert_coordinates, find_homography_cv2, corner_error


def cpu_hot_path(batch, num_samples=5000, seed=0, timings=None, impl="port", noise=None, kde_down=8, return_matches=False):
    """``batch``: a gfnet_b200.synth.PairBatch living on the CPU.  Returns (H list, err list[, matches list]).

    ``noise = (q1 [B,n], q2 [B,n1])``: the two Exp(1) draws of ``torch.multinomial`` (port only; the reference's
    ``GFNet.sample`` draws them itself from the global generator, seeded with ``seed``).
    """
    t = {} if timings is None else timings
    ref = None
    if impl == "reference":
        from . import reference as R
        ref = R.load_reference()
        self_ = R.SampleSelf()
        torch.manual_seed(seed)

    def tick(name, t0):
        t[name] = t.get(name, 0.0) + time.perf_counter() - t0

    t0 = time.perf_counter()
    if ref is not None:
        flow = ref.GFNet.pos_embed(self_, ref.GFNet.corr_volume(self_, batch.coarse_f0, batch.coarse_f1))   # network.py:251-252
    else:
        flow = pos_embed_port(corr_volume_port(batch.coarse_f0, batch.coarse_f1))
    tick("coarse_match", t0)
    t0 = time.perf_counter()
    lc = ref.local_correlation if ref is not None else local_correlation_port
    t_asm = 0.0
    for scales in batch.passes:
        for sc in scales:
            b, c, hs, G, r = sc["f1"].shape[0], sc["c"], sc["hs"], sc["G"], sc["r"]
            for fl in sc["flows"]:
                f0 = sc["f0"]
                if sc.get("x") is not None:          # refiner input assembly, the torch calls of ConvRefiner.forward :537-555
                    ta = time.perf_counter()
                    x_hat = F.grid_sample(sc["f1"], fl.permute(0, 2, 3, 1).contiguous(), align_corners=False, mode="bilinear")
                    tt = torch.linspace(-1 + 1 / G, 1 - 1 / G, G, device=fl.device)
                    gy, gx = torch.meshgrid((tt, tt), indexing="ij")
                    coords = torch.stack((gx, gy))[None].expand(b, 2, G, G)
                    f0 = F.grid_sample(sc["x"], coords.permute(0, 2, 3, 1), align_corners=False, mode="bilinear")
                    emb = F.conv2d(40 / 32 * sc["scale_factor"] * (fl - coords), sc["disp_w"][:, :, None, None], sc["disp_b"])
                    t_asm += time.perf_counter() - ta
                corr = lc((b, c, hs, hs), f0, sc["f1"], r, G, flow=fl)                                       # network.py:553
                if sc.get("x") is not None:
                    ta = time.perf_counter()
                    torch.cat((f0, x_hat, emb, corr), dim=1)                                                 # :555
                    t_asm += time.perf_counter() - ta
    t["refiner_assembly"] = t.get("refiner_assembly", 0.0) + t_asm
    t["local_correlation"] = t.get("local_correlation", 0.0) + time.perf_counter() - t0 - t_asm
    t0 = time.perf_counter()
    warp, cert = match_postprocess_port(batch.final_flow, batch.cert_logits, symmetric=True)                 # :358-384
    tick("match_postprocess", t0)
    Hs, errs, ms = [], [], []
    gen = torch.Generator().manual_seed(seed)
    res = batch.res
    for i in range(batch.B):
        t0 = time.perf_counter()
        n = cert[i].numel()
        if ref is not None:
            m, c = ref.GFNet.sample(self_, warp[i], cert[i], num_samples)                                    # :385-414 (CPU: down=8)
        else:
            if noise is not None:
                q1, q2 = noise[0][i], noise[1][i]
            else:
                q1 = torch.empty(n).exponential_(1, generator=gen)
                q2 = torch.empty(min(4 * num_samples, n)).exponential_(1, generator=gen)
            m, c, _, _, _ = sample_port(warp[i], cert[i], num_samples, q1, q2, half=False, down=kde_down)
        tick("sample_kde", t0)
        t0 = time.perf_counter()
        mn = m.detach().float().cpu().numpy()
        pa, pb = convert_coordinates(mn[:, :2], mn[:, 2:], res, res, res, res)                               # estimation.py:62-64
        H, _, _ = find_homography_cv2(pa, pb)                                                                # :66-77
        errs.append(corner_error(H, batch.H_gt[i].cpu().numpy(), res, res))                                        # :79-92
        Hs.append(H)
        ms.append(mn)
        tick("homography", t0)
    return (Hs, errs, ms) if return_matches else (Hs, errs)
