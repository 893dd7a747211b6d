"""Match post-process and balanced sampling (reference: model/network.py:358-414), batched per pair.

The reference flattens the batch inside ``GFNet.sample`` (:395-398) so it only works for one pair
at a time; here every step carries a leading pair dimension.  The Exp(1) draw of
``torch.multinomial`` stays in torch (caller's generator); everything around it is our kernels.
"""
import torch

from ._lib import check, lib, ptr, require_cuda_f32, stream_ptr
from .ops import kde


def match_postprocess(flow, cert_logits, attenuation=None, symmetric=True):
    """Tail of ``GFNet.match``; reference: model/network.py:358-384.

    ``flow [b,2,G,G]``, ``cert_logits [b,1,G,G]`` (b = 2*B when symmetric: A->B first, then B->A).
    Returns ``warp [B,G,2G,4]``, ``certainty [B,G,2G]`` (or ``[b,G,G,4]``, ``[b,G,G]``).
    """
    fl = require_cuda_f32("flow", flow)
    lg = require_cuda_f32("cert_logits", cert_logits)
    at = require_cuda_f32("attenuation", attenuation) if attenuation is not None else None
    b, _, G, _ = (int(v) for v in fl.shape)
    if lg.numel() != b * G * G or (at is not None and at.numel() != b * G * G):
        raise ValueError("cert_logits / attenuation must be [b,1,G,G]")
    Bp, Wd = (b // 2, 2 * G) if symmetric else (b, G)
    warp = torch.empty((Bp, G, Wd, 4), device=fl.device, dtype=torch.float32)
    cert = torch.empty((Bp, G, Wd), device=fl.device, dtype=torch.float32)
    with torch.cuda.device(fl.device):
        check(lib.gfb_match_postprocess_f32(ptr(fl), ptr(lg), ptr(at), ptr(warp), ptr(cert), b, G, int(bool(symmetric)),
                                            stream_ptr(fl.device)), "match_postprocess")
    return warp, cert


def topk_desc(keys, k):
    """Indices of the k largest per row, descending, ties -> lower index (int64, like torch.topk)."""
    ks = require_cuda_f32("keys", keys)
    B, n = (int(v) for v in ks.shape)
    idx = torch.empty((B, k), device=ks.device, dtype=torch.int64)
    nbytes = lib.gfb_topk_workspace_bytes(B, n, k)
    ws = torch.empty(max(nbytes, 16), device=ks.device, dtype=torch.uint8)
    with torch.cuda.device(ks.device):
        check(lib.gfb_topk_desc_f32(ptr(ks), ptr(idx), B, n, int(k), ptr(ws), nbytes, stream_ptr(ks.device)), "topk")
    return idx


def multinomial_from_noise(p, q, n):
    """``torch.multinomial(p, n, replacement=False)`` given its Exp(1) draw q: topk(p / q, n)."""
    p2 = require_cuda_f32("p", p).reshape(1, -1) if p.dim() == 1 else require_cuda_f32("p", p)
    return topk_desc(p2 / q.reshape(p2.shape), n).reshape((n,) if p.dim() == 1 else (p2.shape[0], n))


def sample_batched(warp, certainty, num=5000, sample_thresh=0.05, balanced=True, generator=None,
                   noise=None, kde_std=0.1, kde_down=1, return_aux=False):
    """``GFNet.sample`` ("threshold_balanced") for a batch of pairs; reference: model/network.py:385-414.

    ``warp [B,...,4]``, ``certainty [B,...]`` -> ``matches [B,num,4]``, ``certainty [B,num]``.
    ``noise = (q1 [B,n], q2 [B,n1])`` supplies the two Exp(1) draws explicitly (parity tests);
    otherwise they are drawn with ``generator`` on the device.  ``kde_down=1`` is what the reference
    does on CUDA (:405), 8 what it does on CPU.
    """
    B = int(warp.shape[0])
    w = require_cuda_f32("warp", warp).reshape(B, -1, 4)
    c = require_cuda_f32("certainty", certainty).reshape(B, -1)
    n = int(c.shape[1])
    dev = w.device
    st = stream_ptr(dev)
    n1 = min((4 if balanced else 1) * num, n)
    q1 = noise[0] if noise is not None else torch.empty((B, n), device=dev).exponential_(1, generator=generator)
    q1 = require_cuda_f32("q1", q1).reshape(B, n)
    key = torch.empty((B, n), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib.gfb_sample_keys_f32(ptr(c), ptr(q1), ptr(key), B * n, float(sample_thresh), st), "sample_keys")
        idx1 = topk_desc(key, n1)
        gm = torch.empty((B, n1, 4), device=dev, dtype=torch.float32)
        gc = torch.empty((B, n1), device=dev, dtype=torch.float32)
        check(lib.gfb_gather_matches_f32(ptr(w), ptr(c), ptr(idx1), ptr(gm), ptr(gc), B, n, n1, float(sample_thresh), st), "gather")
        if not balanced:
            return (gm, gc, idx1, None, None) if return_aux else (gm, gc)
        rho = kde(gm, std=kde_std, half=False, down=kde_down)
        n2 = min(num, n1)
        q2 = noise[1] if noise is not None else torch.empty((B, n1), device=dev).exponential_(1, generator=generator)
        q2 = require_cuda_f32("q2", q2).reshape(B, n1)
        key2 = torch.empty((B, n1), device=dev, dtype=torch.float32)
        check(lib.gfb_balance_keys_f32(ptr(rho), ptr(q2), ptr(key2), B * n1, 10.0, st), "balance_keys")
        idx2 = topk_desc(key2, n2)
        om = torch.empty((B, n2, 4), device=dev, dtype=torch.float32)
        oc = torch.empty((B, n2), device=dev, dtype=torch.float32)
        # certainty was already thresholded by the first gather: thresh = +inf leaves it unchanged
        check(lib.gfb_gather_matches_f32(ptr(gm), ptr(gc), ptr(idx2), ptr(om), ptr(oc), B, n1, n2, float("inf"), st), "gather")
    return (om, oc, idx1, idx2, rho) if return_aux else (om, oc)


def sample(matches, certainty, num=5000, sample_mode="threshold_balanced", sample_thresh=0.05, generator=None):
    """Single-pair ``GFNet.sample(matches, certainty, num)``; reference: model/network.py:385-414."""
    if "threshold" not in sample_mode:
        sample_thresh = float("inf")
    m, c = sample_batched(matches.reshape(1, -1, 4), certainty.reshape(1, -1), num, sample_thresh,
                          balanced="balanced" in sample_mode, generator=generator)
    return m[0], c[0]
