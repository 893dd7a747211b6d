"""Seeded synthetic inputs of the hot path's shapes; data generation only.

The backbone that produces the feature pyramid in the reference (DINOv2 + FPN, model/network.py:
156-201) is out of scope, so the path is fed random-init pyramids whose correlations peak where a
random homography says they should.  The random-H recipe mirrors the reference's 4-corner jitter
(datasets/generate_random_H_large_size.py:6-36).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def pyramid_config(res=448, native=False, upsample_res=None):
    """[(scale, c, hs, G, r)] coarse->fine for the scales that run local_correlation.

    reference shapes: gfnet_configs/basic.json:20,23-26 (feat_chs [64,32,16,8], num_grid
    [32,32,64,128,256], radius [7,6,4,2,0]); model/network.py:185-198; 560 pass: :326-349 (no scale 16).
    The reference always runs at 448 (+ 560); ``res`` = 224 / 672 gives the NATIVE pyramids of SURVEY.md 8(d2)
    (16/28/56/112 with G = 16,16,32,64 and 48/84/168/336 with G = 48,48,96,192: ``num_grid[0] == res / 14`` as
    model/network.py:548 requires); ``native`` is kept for call-site readability only.
    """
    chans = {16: 64, 8: 64, 4: 32, 2: 16}
    radius = {16: 7, 8: 6, 4: 4, 2: 2}
    out = []
    if upsample_res is None:
        g0 = res // 14
        grids = {16: g0, 8: g0, 4: 2 * g0, 2: 4 * g0}
        for s in (16, 8, 4, 2):
            hs = res // 14 if s == 16 else res // s
            out.append((s, chans[s], hs, grids[s], radius[s]))
    else:
        g0 = upsample_res // 14
        grids = {8: g0, 4: 2 * g0, 2: 4 * g0}
        for s in (8, 4, 2):
            out.append((s, chans[s], upsample_res // s, grids[s], radius[s]))
    return out


def final_grid(res=448, upsample_res=560):
    return 8 * ((upsample_res if upsample_res else res) // 14)


def solve_h4(src, dst):
    A = np.zeros((8, 8)); b = np.zeros(8)
    for i in range(4):
        X, Y = src[i]; x, y = dst[i]
        A[2 * i] = [X, Y, 1, 0, 0, 0, -x * X, -x * Y]; b[2 * i] = x
        A[2 * i + 1] = [0, 0, 0, X, Y, 1, -y * X, -y * Y]; b[2 * i + 1] = y
    return np.concatenate((np.linalg.solve(A, b), [1.0])).reshape(3, 3)


def random_homography(gen, jitter=0.15):
    """Normalised-coordinate H: the 4 corners of [-1,1]^2 moved by U(-jitter, jitter) * 2."""
    src = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float64)
    d = (torch.rand(4, 2, generator=gen, dtype=torch.float64).numpy() * 2 - 1) * (2 * jitter)
    return solve_h4(src, src + d)


def to_pixel_homography(Hn, wq, hq, wsup, hsup):
    """Same map between pixel frames under estimation.py:26-45's convention px = (w-1)(x+1)/2."""
    Ta = np.array([[(wq - 1) / 2, 0, (wq - 1) / 2], [0, (hq - 1) / 2, (hq - 1) / 2], [0, 0, 1.0]])
    Tb = np.array([[(wsup - 1) / 2, 0, (wsup - 1) / 2], [0, (hsup - 1) / 2, (hsup - 1) / 2], [0, 0, 1.0]])
    return Tb @ Hn @ np.linalg.inv(Ta)


def lattice(g, device="cpu"):
    t = torch.linspace(-1 + 1 / g, 1 - 1 / g, g, device=device)
    yy, xx = torch.meshgrid(t, t, indexing="ij")
    return torch.stack((xx, yy), 0)                      # [2,G,G] (x,y)


def warp_points(H, pts):
    """pts [2,...] normalised -> H applied, same shape (H: 3x3 tensor)."""
    x, y = pts[0], pts[1]
    d = H[2, 0] * x + H[2, 1] * y + H[2, 2]
    return torch.stack(((H[0, 0] * x + H[0, 1] * y + H[0, 2]) / d, (H[1, 0] * x + H[1, 1] * y + H[1, 2]) / d), 0)


def homography_flow(Hs, g, hs, gen, device, jitter_px=0.5):
    """flow [b,2,G,G]: lattice warped by each H plus N(0, (jitter_px * 2/hs)^2)."""
    lat = lattice(g, device)
    out = []
    for H in Hs:
        Ht = torch.as_tensor(H, dtype=torch.float32, device=device)
        out.append(warp_points(Ht, lat))
    fl = torch.stack(out)
    if jitter_px:
        fl = fl + torch.randn(fl.shape, generator=gen, device=device) * (jitter_px * 2 / hs)
    return fl.float().contiguous()


def scale_inputs(Hs, c, hs, g, gen, device, correlated=True, adversarial=False):
    """(feature0 [b,c,G,G], feature1 [b,c,hs,hs], flow [b,2,G,G]) for one scale."""
    b = len(Hs)
    f1 = torch.randn((b, c, hs, hs), generator=gen, device=device)
    if adversarial:
        flow = (torch.rand((b, 2, g, g), generator=gen, device=device) * 2.4 - 1.2)
    else:
        flow = homography_flow(Hs, g, hs, gen, device)
    if correlated:
        f0 = F.grid_sample(f1, flow.permute(0, 2, 3, 1), mode="bilinear", align_corners=False)
        f0 = f0 + 0.5 * torch.randn(f0.shape, generator=gen, device=device)
    else:
        f0 = torch.randn((b, c, g, g), generator=gen, device=device)
    return f0.contiguous(), f1.contiguous(), flow.contiguous()


def make_matches(H, m, gen, device, sigma=0.002, outlier_frac=0.0):
    """[m,4] normalised matches a -> H(a) + noise; a ~ U(-1,1)^2."""
    a = torch.rand((m, 2), generator=gen, device=device) * 2 - 1
    Ht = torch.as_tensor(H, dtype=torch.float32, device=device)
    bq = warp_points(Ht, a.T).T + torch.randn((m, 2), generator=gen, device=device) * sigma
    nout = int(outlier_frac * m)
    if nout:
        bq[:nout] = torch.rand((nout, 2), generator=gen, device=device) * 2 - 1
    return torch.cat((a, bq), 1).float().contiguous()


WORKLOADS = {   # BASELINE.json configs 2-4: (res, upsample_res, description)
    "visir448": (448, 560, "vis_ir.json 448x448 pairs (pass 1 at 448 + upsample pass at 560), num_itr=2, symmetric"),
    "map224": (224, 280, "map.json googlemap 224x224 pairs, NATIVE pyramid 16/28/56/112 (+ upsample pass at 280 = 1.25 x, the "
                         "reference's 560/448 ratio), num_itr=2, symmetric"),
    "map672": (672, 840, "map.json googlemap 672x672 pairs, NATIVE pyramid 48/84/168/336 (+ upsample pass at 840), num_itr=2, "
                         "symmetric; global match N = 2304"),
}


DISP_DIM = {16: 64, 8: 64, 4: 32, 2: 16}      # gfnet_configs/basic.json:25 displacement_dim (coarse -> fine)


def refiner_features(Hn, c, hs, gen, device, noise=0.5):
    """Feature maps of image A and B of every pair at one scale: B random, A = B warped by the pair's homography
    (+ noise), so that A(p) correlates with B(H p).  Returns (featA, featB) [B,c,hs,hs]."""
    B = len(Hn)
    featB = torch.randn((B, c, hs, hs), generator=gen, device=device)
    lat = lattice(hs, device)
    grid = torch.stack([warp_points(torch.as_tensor(h, dtype=torch.float32, device=device), lat) for h in Hn])
    featA = F.grid_sample(featB, grid.permute(0, 2, 3, 1), mode="bilinear", align_corners=False)
    featA = featA + noise * torch.randn(featA.shape, generator=gen, device=device)
    return featA.contiguous(), featB


class PairBatch:
    """All tensors one step of the hot path consumes for B pairs (symmetric => op batch b = 2B).

    Per scale (reference: model/network.py:213-222 concatenates the two directions, :257-268 runs the refiner):
    ``x`` [2B,c,hs,hs] = cat(featA, featB) is the image-A side of the op batch (UPLOADED), ``f1`` = cat(featB, featA) the
    image-B side (``derive()``: a device-side concatenation of the same maps, as the reference does), ``f0`` =
    grid_sample(x, lattice) the grid features -- what ``ops.refiner_input`` computes on the device; kept here for the CPU
    arms and the kernel-level tools, never uploaded.  ``disp_w`` / ``disp_b``: random-init ``ConvRefiner.disp_emb``.
    """

    def __init__(self, B, res=448, upsample_res=560, num_itr=1, seed=1234, device="cuda", rank=0,
                 pair_offset=0, native=False):
        self.B, self.res, self.upsample_res, self.num_itr = B, res, upsample_res, num_itr
        dev = torch.device(device)
        gen = torch.Generator(device=dev).manual_seed(seed + 1000 * rank + pair_offset)
        cgen = torch.Generator().manual_seed(seed + 1000 * rank + pair_offset)
        self.Hn = [random_homography(cgen) for _ in range(B)]
        Hs = self.Hn + [np.linalg.inv(h) for h in self.Hn]            # A->B then B->A (network.py:213-222)
        self.H_gt = torch.as_tensor(np.stack([to_pixel_homography(h, res, res, res, res) for h in self.Hn]),
                                    dtype=torch.float64, device=dev)
        self.passes = []
        for up in ([None, upsample_res] if upsample_res else [None]):
            scales = []
            for (s, c, hs, g, r) in pyramid_config(res, native, up):
                fa, fb = refiner_features(self.Hn, c, hs, gen, dev)
                x = torch.cat((fa, fb)).contiguous()
                flows = [homography_flow(Hs, g, hs, gen, dev) for _ in range(num_itr)]
                dd = DISP_DIM[s]
                scales.append(dict(scale=s, c=c, hs=hs, G=g, r=r, x=x, f1=torch.empty_like(x), f0=None, flows=flows,
                                   disp_w=(torch.randn((dd, 2), generator=gen, device=dev) * 0.5).contiguous(),
                                   disp_b=(torch.randn((dd,), generator=gen, device=dev) * 0.1).contiguous(),
                                   scale_factor=1.0 if up is None else math.sqrt(up * up / float(res * res))))   # network.py:347
            self.passes.append(scales)
        self.derive(grid_features=True)
        # coarse features for the global match: the scale-16 maps of pass 1 (num_grid[0] == hs there, so the lattice
        # samples the pixel centres and the grid features ARE the map)
        s16 = self.passes[0][0]
        self.coarse_f0, self.coarse_f1 = s16["x"], s16["f1"]
        G = final_grid(res, upsample_res)
        self.G = G
        self.final_flow = homography_flow(Hs, G, G, gen, dev, jitter_px=0.25)
        self.cert_logits = (torch.randn((2 * B, 1, G, G), generator=gen, device=dev) * 2 + 1).contiguous()

    def derive(self, grid_features=False):
        """Image-B side of the op batch from the uploaded maps (in place, same storage every call)."""
        B = self.B
        for scales in self.passes:
            for sc in scales:
                sc["f1"][:B].copy_(sc["x"][B:])
                sc["f1"][B:].copy_(sc["x"][:B])
                if grid_features:
                    lat = lattice(sc["G"], sc["x"].device)[None].expand(2 * B, 2, sc["G"], sc["G"])
                    sc["f0"] = F.grid_sample(sc["x"], lat.permute(0, 2, 3, 1), mode="bilinear", align_corners=False).contiguous()

    def tensors(self):
        """What crosses the host boundary every step: feature maps (once per pair and scale), flows, final flow,
        certainty logits, ground-truth H."""
        out = [self.final_flow, self.cert_logits, self.H_gt]
        for scales in self.passes:
            for sc in scales:
                out += [sc["x"]] + sc["flows"]
        return out
