"""Refiner convolution tail + refinement-loop glue (reference model/network.py:505-531, 557-563, 262-285).

CPU part: the oracle restatement against the golden vector written by the reference's own ConvRefiner modules, and the
host-side batch-norm folding.  GPU part (-m gpu): every kernel through the C ABI against torch on the same inputs, the whole
tail against the oracle / the reference's modules.

Stated tolerance of the tail (the reference runs it under fp16 autocast, rounding to fp16 after every operator; the kernels
store fp16 activations and sum in fp32): against the fp32 evaluation of the same modules the error must not exceed 1.5x the
error of the reference's own fp16-autocast evaluation, and must agree with the float64 model of the kernels' storage
rounding (oracle.refiner_blocks.refiner_tail_fp16_storage) to 2e-3 of max|out|.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import refiner_blocks as ORB

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    z = np.load(os.path.join(HERE, "golden", "refiner_tail.npz"))
    blocks = []
    for i in range(int(z["nblocks"])):
        p = {k: torch.from_numpy(z[f"b{i}_{k}"]) for k in ("w1", "b1", "gamma", "beta", "mean", "var", "w2", "b2")}
        p["eps"] = float(z["eps"])
        blocks.append(p)
    return blocks, torch.from_numpy(z["wout"]), torch.from_numpy(z["bout"]), torch.from_numpy(z["d"]), torch.from_numpy(z["out"])


def _make_blocks(c, nblocks=9, seed=0, out_dim=3):
    """torch modules with the structure of ConvRefiner.create_block (model/network.py:505-531) + out_conv, BN statistics
    randomised (a fresh BatchNorm2d is the identity in eval mode)."""
    torch.manual_seed(seed)
    blocks = []
    for _ in range(nblocks):
        blocks.append(torch.nn.Sequential(torch.nn.Conv2d(c, c, 5, 1, 2, groups=c), torch.nn.BatchNorm2d(c, momentum=0.01),
                                          torch.nn.ReLU(inplace=True), torch.nn.Conv2d(c, c, 1, 1, 0)))
    out_conv = torch.nn.Conv2d(c, out_dim, 1, 1, 0)
    with torch.no_grad():
        for b in blocks:
            b[1].running_mean.normal_(0, 0.2)
            b[1].running_var.uniform_(0.5, 1.5)
            b[1].weight.uniform_(0.7, 1.3)
            b[1].bias.normal_(0, 0.1)
    for m in blocks + [out_conv]:
        m.eval()
    return blocks, out_conv


# ---- CPU ---------------------------------------------------------------------------------------------------------------
def test_oracle_tail_matches_reference_golden():
    blocks, wout, bout, d, out = _golden()
    got = ORB.refiner_tail_port(blocks, wout, bout, d)
    assert got.shape == out.shape == (1, 3, 24, 24)
    assert float((got - out).abs().max()) <= 1e-5 * float(out.abs().max()) + 1e-7
    got64 = ORB.refiner_tail_port(blocks, wout, bout, d, dtype=torch.float64)
    assert float((got64.float() - out).abs().max()) <= 2e-5 * float(out.abs().max())
    sim = ORB.refiner_tail_fp16_storage(blocks, wout, bout, d)               # the kernels' storage rounding, modelled
    assert float((sim.float() - out).abs().max()) <= 2e-2 * float(out.abs().max())


def test_batch_norm_folding_on_host():
    from gfnet_b200.refiner import _fold_block, pad16
    blocks, _ = _make_blocks(21, nblocks=1, seed=3)
    blk = blocks[0]
    cp = pad16(21)
    wf, shift, b2, w2 = _fold_block(blk, cp)
    assert wf.shape == (25, cp) and shift.shape == (cp,) and w2.shape == (cp, cp) and w2.dtype == torch.float16
    assert float(wf[:, 21:].abs().max()) == 0 and float(w2[21:].abs().max()) == 0 and float(w2[:, 21:].abs().max()) == 0
    x = torch.randn(2, 21, 9, 9)
    with torch.no_grad():
        ref = blk[2](blk[1](blk[0](x)))
        got = F.relu(F.conv2d(x, wf[:, :21].t().reshape(21, 1, 5, 5), None, 1, 2, groups=21) + shift[:21].view(1, -1, 1, 1))
    assert float((ref - got).abs().max()) < 1e-5


def test_flow_update_port_zero_rule():
    g = torch.Generator().manual_seed(0)
    df, dc = torch.randn(2, 2, 8, 8, generator=g), torch.randn(2, 1, 8, 8, generator=g)
    flow, cert = torch.rand(2, 2, 8, 8, generator=g), torch.zeros(2, 1, 8, 8)
    pre = torch.zeros_like(flow) + 1e-7
    f1, c1, disp = ORB.flow_update_port(df, dc, flow, cert, pre, 8, 448, 448)
    assert torch.equal(f1, flow + disp) and torch.equal(c1, dc)
    f2, _, disp2 = ORB.flow_update_port(df, dc, f1, c1, disp, 8, 448, 448)     # same delta again: everything is zeroed (:270-271)
    assert float(disp2.abs().max()) == 0 and torch.equal(f2, f1)


def test_no_cpu_path():
    """The product path has no CPU fallback: CPU tensors / modules raise."""
    from gfnet_b200 import refiner as RF
    blocks, oc = _make_blocks(16, nblocks=1)
    with pytest.raises(RuntimeError):
        RF.RefinerBlocks(blocks, oc)
    with pytest.raises(RuntimeError):
        RF.flow_update(torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4), torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4), 8, 448, 448)
    with pytest.raises(RuntimeError):
        RF.upsample_bilinear(torch.zeros(1, 2, 4, 4), 8)
    with pytest.raises(RuntimeError):
        RF.pack_nhwc_f16(torch.zeros(1, 3, 4, 4))


def test_decoder_module_imports_and_signature():
    import inspect
    from gfnet_b200 import decoder
    sig = inspect.signature(decoder.gfnet_forward(lambda *a, **k: None))
    assert list(sig.parameters)[:3] == ["self", "batch", "symmetric"]         # GFNet.forward, model/network.py:203
    assert {"upsample", "scale_factor", "pre_corresps"} <= set(sig.parameters)


# ---- GPU ---------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("c,G,b", [(73, 24, 2), (24, 40, 1), (417, 32, 2), (177, 18, 3)])
def test_pack_nhwc_f16(c, G, b):
    from gfnet_b200 import refiner as RF
    d = torch.randn(b, c, G, G, device="cuda")
    h = RF.pack_nhwc_f16(d)
    cp = RF.pad16(c)
    assert h.shape == (b, G * G, cp) and h.dtype == torch.float16
    ref = d.permute(0, 2, 3, 1).reshape(b, G * G, c).half()
    assert torch.equal(h[..., :c], ref) and float(h[..., c:].abs().max() if cp > c else 0) == 0


@gpu
@pytest.mark.parametrize("c,G,b", [(73, 24, 2), (24, 50, 1), (417, 32, 1), (361, 40, 1), (177, 64, 2), (24, 256, 1)])
def test_dw5_bn_relu_kernel(c, G, b):
    from gfnet_b200 import refiner as RF
    blocks, _ = _make_blocks(c, nblocks=1, seed=c)
    blk = blocks[0].cuda()
    cp = RF.pad16(c)
    wf, shift, _, _ = RF._fold_block(blk, cp)
    d = torch.randn(b, c, G, G, device="cuda")
    h = RF.pack_nhwc_f16(d)
    got = RF.dw5_bn_relu(h, wf, shift, G)
    x = h[..., :c].float().reshape(b, G, G, c).permute(0, 3, 1, 2).double()
    with torch.no_grad():
        ref = blk[2](blk[1].double()(blk[0].double()(x)))
    ref = ref.permute(0, 2, 3, 1).reshape(b, G * G, c)
    err = float((got[..., :c].double() - ref).abs().max())
    assert err <= 1e-3 * float(ref.abs().max()) + 1e-4, err                   # one fp16 rounding of an fp32 sum
    if cp > c:
        assert float(got[..., c:].abs().max()) == 0


@gpu
@pytest.mark.parametrize("c,G,b", [(24, 256, 1), (73, 128, 2), (24, 40, 3), (177, 64, 1), (5, 8, 2)])
def test_dw5_planar_tensor_core_experiment(c, G, b):
    """The channel-planar depth-wise stage as banded MMAs (gfb_debug_refiner_dw5_planar_f16) against torch in
    float64: taps rounded to fp16 like the reference's autocast, fp32 sums, one fp16 rounding of the output."""
    from gfnet_b200 import refiner as RF
    from gfnet_b200._lib import lib, check, ptr, stream_ptr
    cp = RF.pad16(c)
    g = torch.Generator(device="cuda").manual_seed(c + G)
    w = torch.randn((c, 25), generator=g, device="cuda") * 0.2
    wf = torch.zeros((25, cp), device="cuda"); wf[:, :c] = w.t()
    shift = torch.zeros(cp, device="cuda"); shift[:c] = torch.randn(c, generator=g, device="cuda") * 0.1
    x = torch.randn((b, c, G, G), generator=g, device="cuda").half().contiguous()
    out = torch.full_like(x, float("nan"))
    check(lib.gfb_debug_refiner_dw5_planar_f16(ptr(x), ptr(wf), ptr(shift), ptr(out), b, c, cp, G, stream_ptr("cuda")), "dw planar")
    ref = F.relu(F.conv2d(x.double(), w.double().reshape(c, 1, 5, 5), shift[:c].double(), 1, 2, 1, c))
    assert bool(torch.isfinite(out.float()).all())
    assert float((out.double() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-4


@gpu
@pytest.mark.parametrize("algo", [1, 0, 2])
@pytest.mark.parametrize("cp,P", [(32, 1000), (80, 4096 + 37), (192, 128 * 300), (368, 5000), (432, 2048 + 5), (432, 128 * 700),
                                  (256, 777), (16, 130), (512, 640), (48, 3333), (64, 70001), (96, 515), (80, 9), (32, 16 * 5000)])
def test_pointwise_gemm(cp, P, algo):
    """1x1 convolution as a GEMM against torch: algo 0 = auto (streaming mma.sync kernel for Cp <= 96, tcgen05 above), 2 = the
    tcgen05 kernel at every width, 1 = the CUDA-core cross-check."""
    from gfnet_b200 import refiner as RF
    g = torch.Generator(device="cuda").manual_seed(cp + P)
    act = torch.randn(P, cp, generator=g, device="cuda").half()
    w2 = (torch.randn(cp, cp, generator=g, device="cuda") / cp ** 0.5).half()
    bias = torch.randn(cp, generator=g, device="cuda")
    got = RF.pointwise(act, w2, bias, algo=algo)
    ref = act.double() @ w2.double().t() + bias.double()
    err = float((got.double() - ref).abs().max())
    assert err <= 1e-3 * float(ref.abs().max()) + 1e-4, (cp, P, algo, err)


@gpu
@pytest.mark.parametrize("cp,G,b", [(80, 24, 2), (32, 40, 3), (432, 16, 1)])
def test_out_conv(cp, G, b):
    from gfnet_b200 import refiner as RF
    g = torch.Generator(device="cuda").manual_seed(cp)
    act = torch.randn(b, G * G, cp, generator=g, device="cuda").half()
    w = torch.randn(3, cp, generator=g, device="cuda") / cp ** 0.5
    bias = torch.randn(4, generator=g, device="cuda")
    got = RF.out_conv(act, w, bias, b, G)
    ref = (act.double() @ w.double().t() + bias[:3].double()).permute(0, 2, 1).reshape(b, 3, G, G)
    assert float((got.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-6


@gpu
@pytest.mark.parametrize("algo", [1, 0])
def test_tail_vs_reference_golden(algo):
    """The golden vector of the reference's own ConvRefiner modules (scale-2 refiner, C = 73)."""
    from gfnet_b200 import refiner as RF
    blocks, wout, bout, d, out = _golden()
    mods, oc = _make_blocks(73, nblocks=len(blocks), seed=0)
    with torch.no_grad():
        for m, p in zip(mods, blocks):
            m[0].weight.copy_(p["w1"]); m[0].bias.copy_(p["b1"])
            m[1].weight.copy_(p["gamma"]); m[1].bias.copy_(p["beta"]); m[1].running_mean.copy_(p["mean"]); m[1].running_var.copy_(p["var"])
            m[3].weight.copy_(p["w2"]); m[3].bias.copy_(p["b2"])
        oc.weight.copy_(wout); oc.bias.copy_(bout)
    rb = RF.RefinerBlocks([m.cuda() for m in mods], oc.cuda())
    got = rb(d.cuda(), algo=algo).cpu()
    sim = ORB.refiner_tail_fp16_storage(blocks, wout, bout, d).float()
    mx = float(out.abs().max())
    print(f"tail (algo {algo}) vs reference fp32: {float((got - out).abs().max()) / mx:.2e} of max; vs fp16-storage model: "
          f"{float((got - sim).abs().max()) / mx:.2e}")
    assert float((got - sim).abs().max()) <= 2e-3 * mx
    assert float((got - out).abs().max()) <= 2e-2 * mx


@gpu
@pytest.mark.parametrize("c,G,b", [(417, 32, 3), (361, 40, 2), (177, 64, 2), (73, 128, 2), (24, 256, 1)])
def test_tail_vs_torch_fp32_and_autocast(c, G, b):
    """Real refiner widths (model/network.py:79-153).  Truth = the same modules in fp32; the bar is the reference's own
    fp16-autocast error against that truth."""
    from gfnet_b200 import refiner as RF
    mods, oc = _make_blocks(c, seed=c)
    mods = [m.cuda() for m in mods]
    oc = oc.cuda()
    d = torch.randn(b, c, G, G, device="cuda")
    seq = torch.nn.Sequential(*mods)
    # reference arms on PyTorch's native convolution kernels: cuDNN's fp16 depth-wise kernel returned non-finite values for
    # finite inputs on the test boxes (tools/repro_reference_depthwise_nan.py)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=False):
        truth = oc(seq(d.clone()))
        with torch.autocast("cuda", dtype=torch.float16):
            h16 = seq(d.clone())
        ref16 = oc(h16.float())
    assert bool(torch.isfinite(truth).all()) and bool(torch.isfinite(ref16).all())
    rb = RF.RefinerBlocks(mods, oc)
    got = rb(d)
    got1 = rb(d, chunk=1)
    assert torch.equal(got, got1)                                            # chunking does not change a bit
    mx = float(truth.abs().max())
    e_ours, e_ref = float((got - truth).abs().max()), float((ref16 - truth).abs().max())
    print(f"C={c} G={G}: ours vs fp32 {e_ours / mx:.2e}, reference autocast vs fp32 {e_ref / mx:.2e} (of max |out|)")
    assert e_ours <= 1.5 * e_ref + 1e-4 * mx
    assert float((got - ref16).abs().max()) <= 2.5 * e_ref + 1e-4 * mx


@gpu
def test_flow_update_matches_torch_sequence():
    """model/network.py:265-274 evaluated by torch on the same device, two iterations (the second exercises the zero rule)."""
    from gfnet_b200 import refiner as RF
    g = torch.Generator(device="cuda").manual_seed(7)
    B, G = 3, 40
    flow = torch.rand(B, 2, G, G, generator=g, device="cuda") * 2 - 1
    cert = torch.zeros(B, 1, G, G, device="cuda")
    pre = torch.zeros_like(flow) + 1e-7
    f_ref, c_ref, p_ref = flow.clone(), cert.clone(), pre.clone()
    for it in range(3):
        delta = torch.randn(B, 3, G, G, generator=g, device="cuda")
        if it == 2:
            delta = last                                                     # identical update: displacement zeroed
        last = delta
        f_ref, c_ref, p_ref = ORB.flow_update_port(delta[:, :2], delta[:, 2:3], f_ref, c_ref, p_ref, 8, 448, 448)
        RF.flow_update(delta, flow, cert, pre, 8, 448, 448)
        nd = int((flow != f_ref).sum())
        print(f"flow_update iteration {it}: {nd} of {flow.numel()} elements differ from torch")
        assert float((flow - f_ref).abs().max()) <= 1e-7 and torch.equal(cert, c_ref)
        assert float((pre - p_ref).abs().max()) <= 1e-9
    assert float(p_ref.abs().max()) == 0 and float(pre.abs().max()) == 0


@gpu
@pytest.mark.parametrize("hi,ho", [(32, 64), (64, 128), (40, 80), (128, 256), (24, 50)])
def test_upsample_bilinear_matches_interpolate(hi, ho):
    from gfnet_b200 import refiner as RF
    g = torch.Generator(device="cuda").manual_seed(hi)
    x = torch.randn(2, 3, hi, hi, generator=g, device="cuda")
    got = RF.upsample_bilinear(x, ho)
    ref = F.interpolate(x, size=ho, mode="bilinear")
    nd = int((got != ref).sum())
    print(f"upsample {hi}->{ho}: {nd} of {ref.numel()} elements differ from F.interpolate")
    assert float((got - ref).abs().max()) <= 1e-6


@gpu
def test_refiner_blocks_rejects_what_it_does_not_cover():
    from gfnet_b200 import refiner as RF
    mods, oc = _make_blocks(16, nblocks=2)
    mods[0].train()
    with pytest.raises(NotImplementedError):
        RF.RefinerBlocks([m.cuda() for m in mods], oc.cuda())
    mods, oc = _make_blocks(16, nblocks=2)
    with pytest.raises(RuntimeError):
        RF.RefinerBlocks(mods, oc)                                           # CPU module: no CPU path
