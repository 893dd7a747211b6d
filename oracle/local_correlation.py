"""Oracle for ``local_correlation`` (reference: utils/local_correlation.py:4-72).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _window_offsets(r, h, w, num_grid, grid_based, device, dtype):
    """(2r+1)^2 x 2 table of normalised (dx, dy) offsets, k = iy*(2r+1)+ix.

    reference: utils/local_correlation.py:33-52 -- linspace(-2r/n, 2r/n, 2r+1) per axis,
    n = (h, w) of ``featuremap_size`` or ``num_grid`` when ``grid_based_correlation``.
    """
    ny, nx = (num_grid, num_grid) if grid_based else (h, w)
    oy = torch.linspace(-2 * r / ny, 2 * r / ny, 2 * r + 1, device=device, dtype=dtype)
    ox = torch.linspace(-2 * r / nx, 2 * r / nx, 2 * r + 1, device=device, dtype=dtype)
    dy, dx = torch.meshgrid(oy, ox, indexing="ij")
    return torch.stack((dx, dy), dim=-1).reshape(-1, 2)


def _identity_coords(b, h, w, device, dtype):
    """reference: utils/local_correlation.py:21-30 (flow is None => aligned lattice h x w)."""
    ys = torch.linspace(-1 + 1 / h, 1 - 1 / h, h, device=device, dtype=dtype)
    xs = torch.linspace(-1 + 1 / w, 1 - 1 / w, w, device=device, dtype=dtype)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((gx, gy), dim=-1)[None].expand(b, h, w, 2)


def local_correlation_port(featuremap_size, feature0, feature1, local_radius, num_grid,
                           padding_mode="zeros", flow=None, im_A_coords=None,
                           sample_mode="bilinear", grid_based_correlation=False, num_level=1):
    """Torch-CPU port with the reference's operator sequence and cost structure.

    reference: utils/local_correlation.py:4-72.  Per batch element it materialises the
    ``[c, G, G, K]`` sampled window tensor with ``F.grid_sample`` (:56-58) and reduces over
    channels (:60); with ``num_level > 1`` it repeats on 2x average-pooled ``feature1``
    (:61-71).  ``im_A_coords`` is accepted and ignored, as in the reference.
    """
    r = int(local_radius)
    kk = (2 * r + 1) ** 2
    b, c, h, w = featuremap_size
    dev, dt = feature0.device, feature0.dtype
    out = torch.empty((b, kk * num_level, num_grid, num_grid), device=dev, dtype=dt)
    coords = _identity_coords(b, h, w, dev, torch.float32) if flow is None else flow.permute(0, 2, 3, 1)
    offs = _window_offsets(r, h, w, num_grid, grid_based_correlation, dev, torch.float32)
    inv = 1.0 / (c ** 0.5)
    f1 = feature1
    for level in range(num_level):
        for i in range(b):
            with torch.no_grad():
                pts = (coords[i, :, :, None, :] + offs[None, None]).reshape(1, num_grid, num_grid * kk, 2)
                win = F.grid_sample(f1[i:i + 1], pts, mode=sample_mode, padding_mode=padding_mode,
                                    align_corners=False).reshape(c, num_grid, num_grid, kk)
            out[i, kk * level:kk * (level + 1)] = ((feature0[i, ..., None] * inv) * win).sum(0).permute(2, 0, 1)
        if level + 1 < num_level:
            f1 = F.avg_pool2d(f1, kernel_size=2, stride=2)
    return out


def _taps_bilinear(px, py, hs, ws, padding_mode):
    """Four (x, y, weight) taps of ``grid_sample(mode='bilinear', align_corners=False)``.

    reference semantic: ATen grid_sampler_2d as used at utils/local_correlation.py:56-58 --
    unnormalise ``ix = ((x+1)*ws-1)/2``; zeros padding drops out-of-range taps, border padding
    clamps the coordinate to [0, size-1] first.
    """
    ix = ((px + 1.0) * ws - 1.0) / 2.0
    iy = ((py + 1.0) * hs - 1.0) / 2.0
    if padding_mode == "border":
        ix = np.clip(ix, 0.0, ws - 1.0)
        iy = np.clip(iy, 0.0, hs - 1.0)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    fx = ix - x0
    fy = iy - y0
    taps = []
    for dy, wy in ((0, 1.0 - fy), (1, fy)):
        for dx, wx in ((0, 1.0 - fx), (1, fx)):
            taps.append((x0.astype(np.int64) + dx, y0.astype(np.int64) + dy, wx * wy))
    return taps


def local_correlation_def(featuremap_size, feature0, feature1, local_radius, num_grid,
                          padding_mode="zeros", flow=None, im_A_coords=None,
                          sample_mode="bilinear", grid_based_correlation=False, num_level=1):
    """Definition in numpy: explicit taps, float64 accumulation.  Small sizes only.

    corr[b, k, gy, gx] = (1/sqrt(c)) * sum_c f0[b,c,gy,gx] * sample(f1[b,c], coords + offset_k)
    with the coordinate arithmetic (``coords + offset``, unnormalise) carried in float32 exactly
    as the reference does (utils/local_correlation.py:55-58), k = iy*(2r+1)+ix.
    """
    r = int(local_radius)
    kk = (2 * r + 1) ** 2
    b, c, h, w = featuremap_size
    f0 = feature0.detach().cpu().double().numpy()
    out = np.zeros((b, kk * num_level, num_grid, num_grid), dtype=np.float64)
    coords = (_identity_coords(b, h, w, "cpu", torch.float32) if flow is None
              else flow.detach().cpu().float().permute(0, 2, 3, 1))
    offs = _window_offsets(r, h, w, num_grid, grid_based_correlation, "cpu", torch.float32)
    pts = (coords[:, :, :, None, :] + offs[None, None, None]).numpy()        # [b,G,G,K,2] fp32
    f1t = feature1.detach().cpu().float()
    for level in range(num_level):
        f1 = f1t.double().numpy()
        hs, ws = f1.shape[-2:]
        px = pts[..., 0].astype(np.float32)
        py = pts[..., 1].astype(np.float32)
        if sample_mode == "bilinear":
            taps = _taps_bilinear(px, py, np.float32(hs), np.float32(ws), padding_mode)
        elif sample_mode == "nearest":
            ix = ((px + np.float32(1)) * np.float32(ws) - np.float32(1)) / np.float32(2)
            iy = ((py + np.float32(1)) * np.float32(hs) - np.float32(1)) / np.float32(2)
            if padding_mode == "border":
                ix = np.clip(ix, 0, ws - 1)
                iy = np.clip(iy, 0, hs - 1)
            taps = [(np.rint(ix).astype(np.int64), np.rint(iy).astype(np.int64), np.ones_like(ix))]
        else:
            raise NotImplementedError(sample_mode)
        for i in range(b):
            acc = np.zeros((c, num_grid, num_grid, kk), dtype=np.float64)
            for tx, ty, tw in taps:
                ok = (tx[i] >= 0) & (tx[i] < ws) & (ty[i] >= 0) & (ty[i] < hs)
                cx = np.clip(tx[i], 0, ws - 1)
                cy = np.clip(ty[i], 0, hs - 1)
                acc += f1[i][:, cy, cx] * (tw[i].astype(np.float64) * ok)[None]
            out[i, kk * level:kk * (level + 1)] = np.einsum("cyx,cyxk->kyx", f0[i], acc) / math.sqrt(c)
        if level + 1 < num_level:
            f1t = F.avg_pool2d(f1t, kernel_size=2, stride=2)
    return torch.from_numpy(out)
