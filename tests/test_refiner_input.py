"""the refiner input assembly against the reference's own operator sequence (model/network.py:537-555),
run in fp32 on the same GPU inputs: F.grid_sample x 2, the module's own ``disp_emb`` conv, the reference's
``local_correlation`` and ``torch.cat``.  The reference sources come from /root/reference or the staged baseline/_ref.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import reference as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="reference sources not staged (tools/stage_reference.py)")]

# (scale, c, hs, G, r): the 448 pass and the 560 upsample pass (ws = 70 is not a multiple of 4: padded rows)
SHAPES = [(16, 64, 32, 32, 7), (8, 64, 56, 32, 6), (4, 32, 112, 64, 4), (2, 16, 224, 128, 2),
          (8, 64, 70, 40, 6), (4, 32, 140, 80, 4), (2, 16, 280, 160, 2)]


@pytest.fixture(scope="module")
def ref():
    return R.load_reference()


def _reference_d(ref, cr, G, x, y, flow, scale_factor):
    """Lines 537-555 of the reference, fp32 (no autocast), with the reference's own local_correlation."""
    b, c, hs, ws = x.shape
    x_hat = F.grid_sample(y, flow.permute(0, 2, 3, 1).contiguous(), align_corners=False, mode="bilinear")
    t = torch.linspace(-1 + 1 / G, 1 - 1 / G, G, device=x.device)
    gy, gx = torch.meshgrid((t, t), indexing="ij")
    coords = torch.stack((gx, gy))[None].expand(b, 2, G, G)
    grid_feature = F.grid_sample(x, coords.permute(0, 2, 3, 1), align_corners=False, mode="bilinear")
    emb = cr.disp_emb(40 / 32 * scale_factor * (flow - coords))
    lc = ref.local_correlation((b, c, hs, ws), grid_feature, y, local_radius=cr.local_corr_radius, num_grid=G, flow=flow,
                               im_A_coords=None, sample_mode="bilinear", grid_based_correlation=False)
    return torch.cat((grid_feature, x_hat, emb, lc), dim=1)


def _inputs(c, hs, G, b, seed):
    from gfnet_b200 import synth
    gen = torch.Generator(device="cuda").manual_seed(seed)
    cgen = torch.Generator().manual_seed(seed)
    Hs = [synth.random_homography(cgen) for _ in range(b)]
    x = torch.randn((b, c, hs, hs), generator=gen, device="cuda")
    y = torch.randn((b, c, hs, hs), generator=gen, device="cuda")
    flow = synth.homography_flow(Hs, G, hs, gen, "cuda")
    return x, y, flow


@pytest.mark.parametrize("scale,c,hs,G,r", SHAPES)
def test_refiner_input_matches_reference_sequence(ref, scale, c, hs, G, r):
    import gfnet_b200 as gf
    torch.manual_seed(scale)
    cr = R.make_conv_refiner(ref, scale).cuda().eval()
    assert cr.local_corr_radius == r
    x, y, flow = _inputs(c, hs, G, 2, 10 * scale + hs)
    sf = 1.25 if hs in (70, 140, 280) else 1.0                     # the upsample pass runs with scale_factor = 560 / 448
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False        # the reference's 1x1 conv in real fp32 (cuDNN's default is TF32: 2e-4 off)
    try:
        with torch.inference_mode():
            want = _reference_d(ref, cr, G, x, y, flow, sf)
            got = gf.refiner_input(G, x, y, flow, cr.disp_emb.weight, cr.disp_emb.bias, r, sf)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert got.shape == want.shape and got.dtype == torch.float32
    dd = cr.disp_emb.weight.shape[0]
    parts = {"grid_feature": (0, c), "x_hat": (c, 2 * c), "emb": (2 * c, 2 * c + dd), "local_corr": (2 * c + dd, got.shape[1])}
    for name, (a, b_) in parts.items():
        w, g = want[:, a:b_], got[:, a:b_]
        scale_ = float(w.abs().max())
        err = float((g - w).abs().max())
        # grid_sample / conv: fp32 rounding of a handful of operations; correlation: the local-correlation bar
        tol = (4e-5 if name == "local_corr" else 2e-6) * scale_ + (1e-4 * scale_ if name == "local_corr" else 0.0)
        assert err <= tol, f"{name}: {err:.3e} of max {scale_:.3e}"


def test_refiner_input_iterations_reuse_the_prepass(ref):
    """Second refiner iteration of a 64-channel scale with ``prepared=`` (same features, new flow) == a fresh call."""
    import gfnet_b200 as gf
    from gfnet_b200 import synth
    cr = R.make_conv_refiner(ref, 8).cuda().eval()
    x, y, flow = _inputs(64, 56, 32, 3, 5)
    flow2 = (flow + 0.01 * torch.randn_like(flow)).contiguous()
    with torch.inference_mode():
        d, handle = gf.refiner_input(32, x, y, flow, cr.disp_emb.weight, cr.disp_emb.bias, 6, want_prepared=True)
        first = d.clone()
        again = gf.refiner_input(32, x, y, flow2, cr.disp_emb.weight, cr.disp_emb.bias, 6, out=d, prepared=handle)
        fresh = gf.refiner_input(32, x, y, flow2, cr.disp_emb.weight, cr.disp_emb.bias, 6)
        kept = gf.refiner_input(32, x, y, flow2, cr.disp_emb.weight, cr.disp_emb.bias, 6, out=d.clone(), prepared=handle, parts=7)
    assert again.data_ptr() == d.data_ptr()
    assert torch.equal(again, fresh) and torch.equal(kept, fresh)      # parts bit 2: grid features of the first call kept
    assert not torch.equal(first, fresh)


@pytest.mark.parametrize("c,hs,ws,G,scale", [(32, 40, 56, 24, 4), (16, 50, 36, 28, 2), (64, 24, 40, 20, 8)])
def test_refiner_input_non_square_maps(ref, c, hs, ws, G, scale):
    """hs != ws, ws % 4 != 0, G not a multiple of 8: the general shapes of the reference's signature (b, c, hs, ws)."""
    import gfnet_b200 as gf
    from gfnet_b200 import synth
    torch.manual_seed(3)
    cr = R.make_conv_refiner(ref, scale).cuda().eval()
    r = cr.local_corr_radius
    gen = torch.Generator(device="cuda").manual_seed(c + hs)
    x = torch.randn((2, c, hs, ws), generator=gen, device="cuda")
    y = torch.randn((2, c, hs, ws), generator=gen, device="cuda")
    flow = (synth.lattice(G, "cuda")[None] * 0.9 + 0.05 * torch.randn((2, 2, G, G), generator=gen, device="cuda")).contiguous()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.inference_mode():
            want = _reference_d(ref, cr, G, x, y, flow, 1.0)
            got = gf.refiner_input(G, x, y, flow, cr.disp_emb.weight, cr.disp_emb.bias, r)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    k = (2 * r + 1) ** 2
    scale_ = float(want[:, -k:].abs().max())
    assert float((got[:, :-k] - want[:, :-k]).abs().max()) <= 2e-6 * float(want[:, :-k].abs().max())
    assert float((got[:, -k:] - want[:, -k:]).abs().max()) <= 1.4e-4 * scale_


def test_refiner_input_rejects_cpu_tensors():
    import gfnet_b200 as gf
    x = torch.zeros((1, 16, 8, 8))
    with pytest.raises(RuntimeError):
        gf.refiner_input(8, x, x, torch.zeros((1, 2, 8, 8)), torch.zeros((16, 2, 1, 1)), torch.zeros(16), 2)
