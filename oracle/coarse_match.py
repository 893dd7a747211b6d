"""Oracle for the coarse global match (reference: model/network.py:415-440).

TEST INFRASTRUCTURE ONLY.
"""
import math

import numpy as np
import torch


def corr_volume_port(feat0, feat1):
    """vol[b, y1, x1, y0, x0] = <f0[b,:,y0,x0], f1[b,:,y1,x1]> / sqrt(C).

    reference: model/network.py:415-428 (``einsum('bci,bcj->bji')`` then reshape, ``/sqrt(C)``).
    """
    b, c, h0, w0 = feat0.shape
    _, _, h1, w1 = feat1.shape
    a = feat0.reshape(b, c, h0 * w0)
    t = feat1.reshape(b, c, h1 * w1)
    vol = torch.bmm(t.transpose(1, 2), a) / math.sqrt(c)
    return vol.reshape(b, h1, w1, h0, w0)


def coarse_grid(h1, w1, dtype=torch.float32):
    """grid[j = y*W1 + x] = (-1 + (2x+1)/W1, -1 + (2y+1)/H1).

    reference: model/network.py:432-437 -- ``meshgrid(linspace(..W1), linspace(..H1), 'xy')``.
    """
    gx = torch.linspace(-1 + 1 / w1, 1 - 1 / w1, w1)
    gy = torch.linspace(-1 + 1 / h1, 1 - 1 / h1, h1)
    xx, yy = torch.meshgrid(gx, gy, indexing="xy")
    return torch.stack((xx, yy), dim=-1).to(dtype).reshape(h1 * w1, 2)


def pos_embed_port(corr_volume):
    """flow[b, :, i] = sum_j softmax_j(vol[b, j, i]) * grid[j].

    reference: model/network.py:430-440 (softmax over dim=1 = positions of image B).
    """
    b, h1, w1, h0, w0 = corr_volume.shape
    grid = coarse_grid(h1, w1, corr_volume.dtype).to(corr_volume.device)
    p = torch.softmax(corr_volume.reshape(b, h1 * w1, h0, w0), dim=1)
    return torch.einsum("bjhw,jd->bdhw", p, grid).contiguous()


def coarse_match_def(feat0, feat1):
    """Float64 definition of ``pos_embed(corr_volume(f0, f1))`` (model/network.py:415-440)."""
    b, c, h0, w0 = feat0.shape
    _, _, h1, w1 = feat1.shape
    a = feat0.detach().cpu().double().numpy().reshape(b, c, h0 * w0)
    t = feat1.detach().cpu().double().numpy().reshape(b, c, h1 * w1)
    vol = np.einsum("bci,bcj->bji", a, t) / math.sqrt(c)
    vol -= vol.max(axis=1, keepdims=True)
    p = np.exp(vol)
    p /= p.sum(axis=1, keepdims=True)
    grid = coarse_grid(h1, w1, torch.float32).double().numpy()
    flow = np.einsum("bji,jd->bdi", p, grid).reshape(b, 2, h0, w0)
    return torch.from_numpy(flow)
