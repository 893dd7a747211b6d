"""Oracle for the match post-process and balanced sampling (reference: model/network.py:358-414).

TEST INFRASTRUCTURE ONLY.
"""
import torch

from .kde import kde_port


def lattice(g, dtype=torch.float32):
    """grid[gy, gx] = (-1 + (2gx+1)/G, -1 + (2gy+1)/G); reference: model/network.py:362-367."""
    t = torch.linspace(-1 + 1 / g, 1 - 1 / g, g)
    xx, yy = torch.meshgrid(t, t, indexing="xy")
    return torch.stack((xx, yy), dim=-1).to(dtype)


def match_postprocess_port(flow, cert_logits, attenuation=None, symmetric=True):
    """Tail of ``GFNet.match``: finest flow/certainty -> ``warp [B,G,(2)G,4]``, ``certainty [B,G,(2)G]``.

    reference: model/network.py:358-384.  ``flow [b,2,G,G]`` (b = 2B when symmetric, A->B first then
    B->A, :213-222), ``cert_logits [b,1,G,G]``; ``attenuation`` is the already-upsampled
    ``low_res_certainty`` term of :334-340 (or None).  sigmoid (:360), certainty zeroed where any
    |flow| > 1 (:368-370), clamp (:371), symmetric concat along W (:373-378).
    """
    b, _, g, _ = flow.shape
    fl = flow.permute(0, 2, 3, 1).reshape(-1, g, g, 2)
    cert = cert_logits - (attenuation if attenuation is not None else 0)
    cert = torch.sigmoid(cert)
    wrong = (fl.abs() > 1).sum(dim=-1) > 0
    cert = cert.clone()
    cert[wrong[:, None]] = 0
    fl = torch.clamp(fl, -1, 1)
    if symmetric:
        nb = b // 2
        grid = lattice(g, fl.dtype).to(fl.device).expand(nb, g, g, 2)
        a2b, b2a = fl.chunk(2)
        warp = torch.cat((torch.cat((grid, a2b), dim=-1), torch.cat((b2a, grid), dim=-1)), dim=2)
        cert = torch.cat(cert.chunk(2), dim=3)
    else:
        grid = lattice(g, fl.dtype).to(fl.device).expand(b, g, g, 2)
        warp = torch.cat((grid, fl), dim=-1)
    return warp, cert[:, 0]


def multinomial_from_noise(p, q, n):
    """``torch.multinomial(p, n, replacement=False)`` given its Exp(1) draw ``q``.

    reference call sites: model/network.py:400-402, 411-413.  ATen draws ``q ~ Exp(1)``, forms
    ``p / q`` and returns ``topk(n)`` indices (largest first).  Ties (exactly equal quotients) are
    ordered by ascending index here; ATen leaves their order unspecified.
    """
    v = p / q
    order = torch.sort(v, descending=True, stable=True).indices
    return order[:n]


def balanced_probability(density):
    """p = 1/(rho+1), p[rho < 10] = 1e-7; reference: model/network.py:409-410."""
    p = 1 / (density + 1)
    p[density < 10] = 1e-7
    return p


def sample_port(matches, certainty, num, q1, q2, sample_thresh=0.05, half=False, down=8,
                balanced=True):
    """``GFNet.sample`` for one pair with the two Exp(1) draws given explicitly.

    reference: model/network.py:385-414 ("threshold_balanced": threshold at 0.05 (:391-394),
    multinomial of min(4*num, n) (:399-402), kde std 0.1 (:406-408; on CPU the reference uses
    half=False, down=8), p = 1/(rho+1) with the rho<10 floor (:409-410), multinomial of num (:411-413)).
    Returns (matches[num,4], certainty[num], idx1, idx2, density).
    """
    cert = certainty.clone()
    cert[cert > sample_thresh] = 1
    m = matches.reshape(-1, 4)
    cflat = cert.reshape(-1)
    n1 = min((4 if balanced else 1) * num, cflat.numel())
    idx1 = multinomial_from_noise(cflat, q1, n1)
    gm, gc = m[idx1], cflat[idx1]
    if not balanced:
        return gm, gc, idx1, None, None
    rho = kde_port(gm, std=0.1, half=half, down=down)
    p = balanced_probability(rho.float())
    idx2 = multinomial_from_noise(p, q2, min(num, gc.numel()))
    return gm[idx2], gc[idx2], idx1, idx2, rho
