"""Homography estimation + metrics on the device (reference: estimation.py:12-45, 60-92).

``find_homography`` keeps cv2.findHomography's call shape (pos_a, pos_b, method, confidence,
ransacReprojThreshold) -> (H float64 [3,3], mask uint8 [N,1]); ``estimate_homography`` is the batched
form of the tail of ``demo_estimation`` (matches -> pixels -> H -> 4-corner error).
"""
import numpy as np
import torch

from ._lib import check, lib, ptr, require_cuda_f32, stream_ptr

RANSAC = 8          # cv2.RANSAC
LEAST_SQUARES = 0   # cv2 method 0


def convert_coordinates(im_A_coords, im_A_to_im_B, wq, hq, wsup, hsup):
    """normalised -> pixel, ``(w-1)(x+1)/2``; reference: estimation.py:26-45 (torch or numpy input)."""
    stack = torch.stack if isinstance(im_A_coords, torch.Tensor) else np.stack
    a = stack(((wq - 1) * (im_A_coords[..., 0] + 1) / 2, (hq - 1) * (im_A_coords[..., 1] + 1) / 2), -1)
    b = stack(((wsup - 1) * (im_A_to_im_B[..., 0] + 1) / 2, (hsup - 1) * (im_A_to_im_B[..., 1] + 1) / 2), -1)
    return a, b


def estimate_homography(matches, wq, hq, wsup, hsup, weights=None, n_hyp=None, thresh=3.0, gn_iters=10,
                        seed=0, return_mask=False, pixel_input=False, confidence=0.99999, max_iters=2000, return_iters=False):
    """matches ``[B,N,4]`` (normalised) -> ``H [B,3,3]`` float64, ``status [B]``, ``n_inliers [B]``.

    reference: estimation.py:60-77 (convert_coordinates + cv2.findHomography RANSAC thr 3, confidence 0.99999 + the
    diag(0,0,1) fallback).  Default (``n_hyp=None``): OpenCV's own RANSAC loop restated on the device (its RNG, subset
    checks, adaptive iteration count; ``gfb_homography_cv_f32``) -- the same H and mask as cv2.findHomography.
    ``n_hyp=K > 0``: K hash-drawn hypotheses scored exhaustively (round-1 solver, accepts ``weights``); ``n_hyp=0`` skips
    RANSAC: weighted DLT + refinement on all points.
    """
    m = require_cuda_f32("matches", matches)
    if m.dim() != 3 or m.shape[2] != 4:
        raise ValueError("matches must be [B,N,4]")
    B, N = int(m.shape[0]), int(m.shape[1])
    w = require_cuda_f32("weights", weights).reshape(B, N) if weights is not None else None
    dev = m.device
    H = torch.empty((B, 3, 3), device=dev, dtype=torch.float64)
    status = torch.empty((B,), device=dev, dtype=torch.int32)
    ninl = torch.empty((B,), device=dev, dtype=torch.int32)
    mask = torch.empty((B, N), device=dev, dtype=torch.uint8) if return_mask else None
    if pixel_input:
        wq = 0.0
    if n_hyp is None:
        if w is not None:
            raise ValueError("the cv2-faithful solver takes no weights (cv2.findHomography has none); pass n_hyp=K or 0")
        nbytes = lib.gfb_homography_workspace_bytes(B, N, 1)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        iters = torch.empty((B,), device=dev, dtype=torch.int32) if return_iters else None
        with torch.cuda.device(dev):
            rc = lib.gfb_homography_cv_f32(ptr(m), B, N, float(wq), float(hq), float(wsup), float(hsup), float(thresh),
                                           int(max_iters), float(confidence), int(gn_iters), ptr(H), ptr(status), ptr(ninl),
                                           ptr(mask), ptr(iters), ptr(ws), nbytes, stream_ptr(dev))
        check(rc, "homography (cv2-faithful)")
        out = (H, status, ninl, mask) if return_mask else (H, status, ninl)
        return out + (iters,) if return_iters else out
    nbytes = lib.gfb_homography_workspace_bytes(B, N, n_hyp)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        rc = lib.gfb_homography_f32(ptr(m), ptr(w), B, N, float(wq), float(hq), float(wsup), float(hsup), int(n_hyp),
                                    float(thresh), int(gn_iters), int(seed) & 0xFFFFFFFF, ptr(H), ptr(status), ptr(ninl),
                                    ptr(mask), ptr(ws), nbytes, stream_ptr(dev))
    check(rc, "homography")
    return (H, status, ninl, mask) if return_mask else (H, status, ninl)


def find_homography(pos_a, pos_b, method=RANSAC, ransacReprojThreshold=3.0, confidence=0.99999, maxIters=2000,
                    n_hyp=None, seed=0):
    """cv2.findHomography-shaped call for one pair; reference call site: estimation.py:66-72.

    ``pos_a, pos_b [N,2]`` pixel coordinates (CUDA) -> ``(H [3,3] float64 numpy or None, mask [N,1] uint8)``.
    ``method=RANSAC`` runs OpenCV's own loop (``confidence``, ``maxIters`` as in cv2) unless ``n_hyp`` asks for the
    fixed-budget solver; ``method=0`` is the least-squares fit on all points.
    """
    pa = require_cuda_f32("pos_a", pos_a).reshape(-1, 2)
    pb = require_cuda_f32("pos_b", pos_b).reshape(-1, 2)
    m = torch.cat((pa, pb), dim=1)[None].contiguous()
    H, status, _, mask = estimate_homography(m, 0, 0, 0, 0, n_hyp=n_hyp if method == RANSAC else 0,
                                             thresh=ransacReprojThreshold, seed=seed, return_mask=True, pixel_input=True,
                                             confidence=confidence, max_iters=maxIters)
    if int(status[0].item()) == 0:
        return None, mask[0].reshape(-1, 1).cpu().numpy()
    return H[0].cpu().numpy(), mask[0].reshape(-1, 1).cpu().numpy()


def corner_error(H_pred, H_gt, w, h, clip=70.0):
    """Mean 4-corner transfer error, clipped at 70 px; reference: estimation.py:79-92.  ``[B,3,3]`` float64."""
    hp = H_pred.to(torch.float64).contiguous()
    hg = H_gt.to(device=hp.device, dtype=torch.float64).contiguous()
    B = int(hp.shape[0])
    err = torch.empty((B,), device=hp.device, dtype=torch.float32)
    with torch.cuda.device(hp.device):
        check(lib.gfb_corner_error_f64(ptr(hp), ptr(hg), ptr(err), B, float(w), float(h), float(clip), stream_ptr(hp.device)),
              "corner_error")
    return err


def auc(errors, thresholds):
    """AUC of the recall-vs-error curve; reference: estimation.py:12-24 (host-side, a few floats)."""
    e = np.sort(np.asarray(errors, dtype=np.float64))
    n = len(e)
    e = np.concatenate(([0.0], e))
    rec = np.concatenate(([0.0], (np.arange(n) + 1) / n))
    out = []
    for t in thresholds:
        last = int(np.searchsorted(e, t))
        x = np.concatenate((e[:last], [t]))
        y = np.concatenate((rec[:last], [rec[last - 1]]))
        out.append(float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) * 0.5) / t))
    return out
