"""Drop-in test through the REAL reference module on the GPU: ``patch(model.network)`` and compare the reference's
own ``ConvRefiner.forward`` / ``GFNet.sample`` / ``corr_volume`` + ``pos_embed`` patched against unpatched.

The reference is imported unmodified from /root/reference or from the staged copy baseline/_ref
(tools/stage_reference.py; `__graft_entry__.build()` stages it).  Call sites exercised:
model/network.py:10-11 (module-level names), :251-252 (coarse match), :385-414 (sample), :533-564 (ConvRefiner.forward,
inside the fp16 autocast context of :535-536).
"""
import pytest
import torch

from oracle import reference as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="reference sources not staged (tools/stage_reference.py)")]


@pytest.fixture(scope="module")
def ref():
    return R.load_reference()


@pytest.fixture()
def patched(ref):
    from gfnet_b200.patch import patch, unpatch
    saved = patch(ref.network, ref.utils_local_correlation, ref.utils_kde)
    yield ref
    unpatch(ref.network, saved)
    ref.utils_local_correlation.local_correlation = saved["local_correlation"]
    ref.utils_kde.kde = saved["kde"]


def _inputs(scale, b, seed):
    from gfnet_b200 import synth
    cfg = {16: (64, 32, 32), 8: (64, 56, 32), 4: (32, 112, 64), 2: (16, 224, 128)}[scale]
    c, hs, G = cfg
    gen = torch.Generator(device="cuda").manual_seed(seed)
    cgen = torch.Generator().manual_seed(seed)
    Hs = [synth.random_homography(cgen) for _ in range(b)]
    x = torch.randn((b, c, hs, hs), generator=gen, device="cuda")            # image-A feature map
    y = torch.randn((b, c, hs, hs), generator=gen, device="cuda")            # image-B feature map
    flow = synth.homography_flow(Hs, G, hs, gen, "cuda")
    return x, y, flow, G


@pytest.mark.parametrize("scale", [16, 8, 4, 2])
def test_conv_refiner_forward_patched_vs_unpatched(ref, scale):
    """ConvRefiner.forward (model/network.py:533-564) with random-init weights, eval mode, fp16 autocast as shipped."""
    from gfnet_b200.patch import patch, unpatch
    torch.manual_seed(scale)
    cr = R.make_conv_refiner(ref, scale).cuda().eval()
    x, y, flow, G = _inputs(scale, 2, 100 + scale)
    with torch.inference_mode():                                             # GFNet.match is @torch.inference_mode()
        # reference arm through PyTorch's native convolution kernels: on this image cuDNN's fp16 depth-wise kernel returns
        # non-finite values for these finite inputs once other refiner shapes have run in the process (reproduced with
        # the reference alone, tools/repro_reference_depthwise_nan.py, profiles/r2_reference_cudnn_depthwise_nan.txt)
        with torch.backends.cudnn.flags(enabled=False):
            d0, c0, lc0 = cr(G, x, y, flow)
        assert bool(torch.isfinite(d0.float()).all()) and bool(torch.isfinite(c0.float()).all())
        saved = patch(ref.network)
        try:
            assert ref.network.local_correlation is not saved["local_correlation"]
            d1, c1, lc1 = cr(G, x, y, flow)
        finally:
            unpatch(ref.network, saved)
    assert lc1.dtype == lc0.dtype == torch.float32 and lc1.shape == lc0.shape
    scale_lc = float(lc0.abs().max())
    err = float((lc1 - lc0).abs().max())
    assert err <= 4e-5 * scale_lc + 1e-4 * scale_lc, f"local_corr through ConvRefiner: {err:.3e} of max {scale_lc:.3e}"
    # the 9 depth-wise conv blocks run in fp16 (autocast): outputs agree to fp16 noise on identical inputs
    for a, b_ in ((d0, d1), (c0, c1)):
        tol = 2e-2 * float(a.float().abs().max()) + 1e-3
        assert float((a.float() - b_.float()).abs().max()) <= tol


def test_coarse_match_patched_vs_unpatched(ref):
    """GFNet.corr_volume + GFNet.pos_embed (model/network.py:415-440), the two-call sequence of :251-252."""
    from gfnet_b200.patch import patch, unpatch
    gen = torch.Generator(device="cuda").manual_seed(5)
    f0 = torch.randn((2, 64, 32, 32), generator=gen, device="cuda")
    f1 = torch.randn((2, 64, 32, 32), generator=gen, device="cuda")
    self = R.SampleSelf()
    with torch.inference_mode():
        flow0 = ref.GFNet.pos_embed(self, ref.GFNet.corr_volume(self, f0, f1))
        saved = patch(ref.network)
        try:
            flow1 = ref.network.GFNet.pos_embed(self, ref.network.GFNet.corr_volume(self, f0, f1))
        finally:
            unpatch(ref.network, saved)
    assert float((flow1 - flow0).abs().max()) < 1e-5


def _candidates(seed, G=64):
    from gfnet_b200 import synth
    gen = torch.Generator(device="cuda").manual_seed(seed)
    cgen = torch.Generator().manual_seed(seed)
    H = synth.random_homography(cgen)
    lat = synth.lattice(G, "cuda")
    wb = synth.warp_points(torch.as_tensor(H, dtype=torch.float32, device="cuda"), lat)
    warp = torch.cat((lat, wb), 0).permute(1, 2, 0).contiguous()             # [G,G,4]
    cert = torch.rand((G, G), generator=gen, device="cuda") * 0.9 + 0.05     # all positive: no ties at key = 0
    return warp, cert


def test_sample_threshold_patched_is_bit_exact(ref):
    """First multinomial draw of GFNet.sample (model/network.py:400-402): same CUDA generator state -> same matches."""
    from gfnet_b200.patch import patch, unpatch
    warp, cert = _candidates(3)

    class S(R.SampleSelf):
        sample_mode = "threshold"

    torch.manual_seed(1234)
    m0, c0 = ref.GFNet.sample(S(), warp, cert, 2000)
    saved = patch(ref.network)
    try:
        torch.manual_seed(1234)
        m1, c1 = ref.network.GFNet.sample(S(), warp, cert, 2000)
    finally:
        unpatch(ref.network, saved)
    assert torch.equal(m0, m1) and torch.equal(c0, c1)


def test_sample_balanced_patched_agrees_with_reference(ref):
    """Balanced sampling (:403-414).  The reference's CUDA path runs kde in fp16 (~8 % density error), ours
    in fp32, so the second draw differs where p is borderline: the first draw is bit-identical (same Exp(1) stream),
    the second-draw overlap is reported and bounded."""
    from gfnet_b200.patch import patch, unpatch
    warp, cert = _candidates(4, G=96)
    torch.manual_seed(99)
    m0, c0 = ref.GFNet.sample(R.SampleSelf(), warp, cert, 1500)
    saved = patch(ref.network)
    try:
        torch.manual_seed(99)
        m1, c1 = ref.network.GFNet.sample(R.SampleSelf(), warp, cert, 1500)
    finally:
        unpatch(ref.network, saved)
    assert m1.shape == m0.shape == (1500, 4) and c1.shape == c0.shape
    a = {tuple(r) for r in m0.cpu().numpy().round(6).tolist()}
    b = {tuple(r) for r in m1.cpu().numpy().round(6).tolist()}
    overlap = len(a & b) / 1500.0
    print(f"balanced sample overlap patched vs reference (fp16 kde): {overlap:.3f}")
    cand = {tuple(r) for r in warp.reshape(-1, 4).cpu().numpy().round(6).tolist()}
    assert b <= cand                                                          # every returned match is a candidate
    assert overlap > 0.5


def test_prepared_features_under_inference_mode():
    """PreparedFeatures.matches must not touch ``_version`` of inference tensors."""
    import gfnet_b200 as gf
    from gfnet_b200 import synth
    with torch.inference_mode():
        gen = torch.Generator(device="cuda").manual_seed(0)
        cgen = torch.Generator().manual_seed(0)
        Hs = [synth.random_homography(cgen)]
        f0, f1, flow = synth.scale_inputs(Hs, 64, 32, 32, gen, "cuda")
        prep = gf.local_correlation_prepare((1, 64, 32, 32), f0, f1, 7, 32)
        a = gf.local_correlation((1, 64, 32, 32), f0, f1, 7, 32, flow=flow, prepared=prep)
        b = gf.local_correlation((1, 64, 32, 32), f0, f1, 7, 32, flow=flow)
    assert torch.equal(a, b)
