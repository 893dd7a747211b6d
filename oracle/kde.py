"""Oracle for ``kde`` (reference: utils/kde.py:4-13).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch


def kde_port(x, std=0.1, half=True, down=None):
    """Torch-CPU port: Gaussian kernel density through ``torch.cdist`` as the reference does.

    reference: utils/kde.py:4-13 -- optional ``.half()`` (:6-7), ``cdist(x, x[::down])`` squared,
    divided by ``2 std^2``, negated, ``exp``, summed over the last axis (:8-13).
    """
    if half:
        x = x.half()
    y = x if down is None else x[::down]
    scores = torch.exp(-(torch.cdist(x, y) ** 2) / (2 * std ** 2))
    return scores.sum(dim=-1)


def kde_def(x, std=0.1, down=None, chunk=2048):
    """Definition in float64: density[m] = sum_m' exp(-||x_m - y_m'||^2 / (2 std^2)), y = x[::down].

    reference: utils/kde.py:8-12 with exact pairwise differences (no matmul cancellation).
    """
    xn = x.detach().cpu().double().numpy()
    yn = xn if down is None else xn[::down]
    out = np.empty(xn.shape[0], dtype=np.float64)
    for s in range(0, xn.shape[0], chunk):
        d2 = ((xn[s:s + chunk, None, :] - yn[None, :, :]) ** 2).sum(-1)
        out[s:s + chunk] = np.exp(-d2 / (2.0 * std * std)).sum(-1)
    return torch.from_numpy(out)
