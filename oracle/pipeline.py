"""CPU hot path for one batch of pairs, built from the oracle ports in the reference's call order
(SURVEY.md 3.1).  TEST INFRASTRUCTURE / CPU BASELINE ONLY (bench.py ``cpu_baseline`` and
``--impl reference``): this is what the reference's own functions cost on host cores."""
import time

import numpy as np
import torch

from .coarse_match import corr_volume_port, pos_embed_port
from .local_correlation import local_correlation_port
from .sampling import match_postprocess_port, sample_port
from .estimation import convert_coordinates, find_homography_cv2, corner_error


def cpu_hot_path(batch, num_samples=5000, seed=0, timings=None):
    """``batch``: a gfnet_b200.synth.PairBatch living on the CPU.  Returns (H list, err list)."""
    t = {} if timings is None else timings

    def tick(name, t0):
        t[name] = t.get(name, 0.0) + time.perf_counter() - t0

    t0 = time.perf_counter()
    flow = pos_embed_port(corr_volume_port(batch.coarse_f0, batch.coarse_f1))       # network.py:251-252
    tick("coarse_match", t0)
    t0 = time.perf_counter()
    for scales in batch.passes:
        for sc in scales:
            b, c, hs, G, r = sc["f1"].shape[0], sc["c"], sc["hs"], sc["G"], sc["r"]
            for fl in sc["flows"]:
                local_correlation_port((b, c, hs, hs), sc["f0"], sc["f1"], r, G, flow=fl)   # network.py:553
    tick("local_correlation", t0)
    t0 = time.perf_counter()
    warp, cert = match_postprocess_port(batch.final_flow, batch.cert_logits, symmetric=True)  # :358-384
    tick("match_postprocess", t0)
    Hs, errs = [], []
    gen = torch.Generator().manual_seed(seed)
    res = batch.res
    for i in range(batch.B):
        t0 = time.perf_counter()
        n = cert[i].numel()
        q1 = torch.empty(n).exponential_(1, generator=gen)
        q2 = torch.empty(min(4 * num_samples, n)).exponential_(1, generator=gen)
        m, c, _, _, _ = sample_port(warp[i], cert[i], num_samples, q1, q2, half=False, down=8)  # :385-414 (CPU: down=8)
        tick("sample_kde", t0)
        t0 = time.perf_counter()
        mn = m.numpy()
        pa, pb = convert_coordinates(mn[:, :2], mn[:, 2:], res, res, res, res)          # estimation.py:62-64
        H, _, _ = find_homography_cv2(pa, pb)                                            # :66-77
        errs.append(corner_error(H, batch.H_gt[i].numpy(), res, res))                   # :79-92
        Hs.append(H)
        tick("homography", t0)
    return Hs, errs
