"""Generate the golden fixtures by calling the REFERENCE's own functions.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference in place (a stub ``romatch.utils.utils`` module is
needed because model/transformer/__init__.py:5 imports romatch), evaluates the hot-path functions
on small seeded inputs and writes ``tests/golden/*.npz``.  The fixtures are committed; the tests
compare the oracle and the CUDA path against them.  Nothing here is product code.
"""
import logging
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GFNET_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    import utils.utils as U
    for n in ("romatch", "romatch.utils", "romatch.utils.utils"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["romatch.utils.utils"].get_grid = U.get_grid
    sys.modules["romatch.utils.utils"].get_autocast_params = U.get_autocast_params
    logging.disable(logging.WARNING)
    import model.network as N
    from utils.local_correlation import local_correlation
    from utils.kde import kde
    return N, local_correlation, kde


def rand_h(gen, jitter=0.3):
    """Random homography in normalised coords: unit-square corners jittered (cf. the reference's
    datasets/generate_random_H_large_size.py:6-36 4-corner recipe)."""
    import cv2
    src = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float32)
    dst = src + (torch.rand(4, 2, generator=gen).numpy().astype(np.float32) * 2 - 1) * jitter
    return cv2.getPerspectiveTransform(src, dst)


def lattice_flow(gen, b, g, hs, jitter_px=0.5, adversarial=False):
    if adversarial:
        return (torch.rand(b, 2, g, g, generator=gen) * 2.4 - 1.2).float()
    t = torch.linspace(-1 + 1 / g, 1 - 1 / g, g)
    yy, xx = torch.meshgrid(t, t, indexing="ij")
    out = []
    for _ in range(b):
        H = torch.from_numpy(rand_h(gen)).float()
        p = torch.stack((xx, yy, torch.ones_like(xx)), 0).reshape(3, -1)
        q = H @ p
        q = (q[:2] / q[2:]).reshape(2, g, g)
        out.append(q + torch.randn(2, g, g, generator=gen) * (jitter_px * 2 / hs))
    return torch.stack(out).float()


def main():
    N, local_correlation, kde = import_reference()
    gen = torch.Generator().manual_seed(20251017)

    # ---- local_correlation: (b, c, hs, ws, G, r, kwargs, adversarial)
    cases = [
        (2, 8, 12, 12, 6, 2, {}, False),
        (2, 16, 14, 10, 8, 3, {}, False),                       # non-square hs != ws, G != hs
        (1, 8, 9, 9, 9, 1, {}, True),                           # |flow| > 1: zero padding
        (1, 4, 8, 8, 8, 0, {}, False),                          # r = 0
        (2, 8, 12, 12, 6, 2, {"grid_based_correlation": True}, False),
        (1, 8, 16, 16, 8, 1, {"num_level": 2}, False),
        (1, 8, 10, 10, 5, 2, {"sample_mode": "nearest"}, False),
        (1, 8, 10, 10, 5, 2, {"padding_mode": "border"}, True),
        (1, 6, 7, 7, 7, 2, {"flow": None}, False),              # flow=None: identity lattice h x w
    ]
    lc = {}
    for i, (b, c, hs, ws, g, r, kw, adv) in enumerate(cases):
        f0 = torch.randn(b, c, g, g, generator=gen)
        f1 = torch.randn(b, c, hs, ws, generator=gen)
        kw = dict(kw)
        if "flow" in kw:
            flow = None
            kw.pop("flow")
        else:
            flow = lattice_flow(gen, b, g, hs, adversarial=adv)
        out = local_correlation((b, c, hs, ws), f0, f1, local_radius=r, num_grid=g, flow=flow, **kw)
        lc[f"c{i}_meta"] = np.array([b, c, hs, ws, g, r, int(kw.get("grid_based_correlation", False)),
                                     kw.get("num_level", 1), int(flow is None)], dtype=np.int64)
        lc[f"c{i}_mode"] = np.array([kw.get("sample_mode", "bilinear"), kw.get("padding_mode", "zeros")])
        lc[f"c{i}_f0"], lc[f"c{i}_f1"] = f0.numpy(), f1.numpy()
        if flow is not None:
            lc[f"c{i}_flow"] = flow.numpy()
        lc[f"c{i}_out"] = out.numpy()
    lc["ncases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "local_correlation.npz"), **lc)

    # ---- coarse global match
    cm = {}
    for i, (b, c, h0, w0, h1, w1) in enumerate([(2, 16, 6, 5, 6, 5), (1, 64, 8, 8, 8, 8), (2, 8, 4, 7, 5, 3)]):
        f1 = torch.randn(b, c, h1, w1, generator=gen)
        f0 = torch.randn(b, c, h0, w0, generator=gen) * 1.5
        vol = N.GFNet.corr_volume(None, f0, f1)
        flow = N.GFNet.pos_embed(None, vol)
        cm[f"c{i}_f0"], cm[f"c{i}_f1"], cm[f"c{i}_vol"], cm[f"c{i}_flow"] = f0.numpy(), f1.numpy(), vol.numpy(), flow.numpy()
    cm["ncases"] = np.array(3)
    np.savez_compressed(os.path.join(HERE, "coarse_match.npz"), **cm)

    # ---- kde (fp32 path; the reference's CPU setting is half=False, down=8)
    kd = {}
    for i, (m, down) in enumerate([(300, None), (257, 8), (64, 3)]):
        a = torch.rand(m, 2, generator=gen) * 2 - 1
        H = torch.from_numpy(rand_h(gen)).float()
        q = H @ torch.cat((a, torch.ones(m, 1)), 1).T
        x = torch.cat((a, (q[:2] / q[2:]).T + torch.randn(m, 2, generator=gen) * 0.002), 1).float()
        kd[f"c{i}_x"] = x.numpy()
        kd[f"c{i}_down"] = np.array(-1 if down is None else down)
        kd[f"c{i}_out"] = kde(x, std=0.1, half=False, down=down).numpy()
    kd["ncases"] = np.array(3)
    np.savez_compressed(os.path.join(HERE, "kde.npz"), **kd)

    # ---- GFNet.sample (threshold_balanced) with a recorded generator seed
    class S:
        sample_mode = "threshold_balanced"
        sample_thresh = 0.05
    sm = {}
    g = 24
    t = torch.linspace(-1 + 1 / g, 1 - 1 / g, g)
    xx, yy = torch.meshgrid(t, t, indexing="xy")
    H = torch.from_numpy(rand_h(gen)).float()
    src = torch.stack((xx, yy), -1).reshape(-1, 2)
    q = H @ torch.cat((src, torch.ones(len(src), 1)), 1).T
    warp = torch.cat((src, (q[:2] / q[2:]).T + torch.randn(len(src), 2, generator=gen) * 0.003), 1).reshape(g, g, 4)
    cert = torch.rand(g, g, generator=gen) ** 4
    torch.manual_seed(777)
    gm, gc = N.GFNet.sample(S, warp, cert, 100)
    sm["warp"], sm["cert"], sm["seed"], sm["num"] = warp.numpy(), cert.numpy(), np.array(777), np.array(100)
    sm["good_matches"], sm["good_certainty"] = gm.numpy(), gc.numpy()
    np.savez_compressed(os.path.join(HERE, "sample.npz"), **sm)

    # ---- cv2.findHomography with the reference's arguments (estimation.py:66-72)
    import cv2
    hm = {}
    w = h = 448
    for i, (npts, sigma, outl) in enumerate([(500, 0.0, 0.0), (500, 0.25, 0.1), (800, 0.5, 0.2)]):
        Hn = rand_h(gen)
        a = (torch.rand(npts, 2, generator=gen) * 2 - 1).numpy().astype(np.float64)
        qh = np.concatenate((a, np.ones((npts, 1))), 1) @ Hn.T
        bn = qh[:, :2] / qh[:, 2:]
        pa = np.stack(((w - 1) * (a[:, 0] + 1) / 2, (h - 1) * (a[:, 1] + 1) / 2), -1)
        pb = np.stack(((w - 1) * (bn[:, 0] + 1) / 2, (h - 1) * (bn[:, 1] + 1) / 2), -1)
        pb = pb + torch.randn(npts, 2, generator=gen).numpy() * sigma
        nout = int(outl * npts)
        if nout:
            pb[:nout] = torch.rand(nout, 2, generator=gen).numpy() * (w - 1)
        pa32, pb32 = pa.astype(np.float32), pb.astype(np.float32)
        Hr, mask = cv2.findHomography(pa32, pb32, method=cv2.RANSAC, confidence=0.99999, ransacReprojThreshold=3)
        Hls, _ = cv2.findHomography(pa32[nout:], pb32[nout:], method=0)
        T = np.array([[(w - 1) / 2, 0, (w - 1) / 2], [0, (h - 1) / 2, (h - 1) / 2], [0, 0, 1.0]])
        hm[f"c{i}_pa"], hm[f"c{i}_pb"] = pa32, pb32
        hm[f"c{i}_H_ransac"], hm[f"c{i}_mask"] = Hr, mask
        hm[f"c{i}_H_lsq_inliers"], hm[f"c{i}_nout"] = Hls, np.array(nout)
        hm[f"c{i}_H_true_px"] = T @ Hn @ np.linalg.inv(T)
    hm["ncases"] = np.array(3)
    hm["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "homography_cv2.npz"), **hm)
    print("wrote fixtures to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
