"""CPU restatement of the refiner's convolution tail and of the refinement-loop glue (TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by gfnet_b200).

reference: ConvRefiner.create_block (model/network.py:505-531: Conv2d(k=5, groups=in_dim) -> BatchNorm2d -> ReLU -> Conv2d 1x1),
the tail of ConvRefiner.forward (:557-563: block1, hidden_blocks under autocast, out_conv on d.float()) and the loop body of
GFNet.decoder (:262-285).  On the CPU the reference's autocast is off (utils/utils.py:306-320), so its CPU path is fp32: that is
``refiner_tail_port``.  Pinned against the reference's own ConvRefiner modules by tests/golden/refiner_tail.npz
(tests/golden/make_golden_refiner.py).
"""
import torch
import torch.nn.functional as F


def block_params(block):
    """Tensors of one create_block Sequential (conv1, norm, relu, conv2) as a dict (state-dict names of model/network.py:528-531)."""
    conv1, norm, _, conv2 = block[0], block[1], block[2], block[3]
    return {"w1": conv1.weight.detach(), "b1": None if conv1.bias is None else conv1.bias.detach(),
            "gamma": norm.weight.detach(), "beta": norm.bias.detach(), "mean": norm.running_mean.detach(),
            "var": norm.running_var.detach(), "eps": norm.eps, "w2": conv2.weight.detach(), "b2": conv2.bias.detach()}


def module_params(refiner):
    """(list of block dicts, out_conv weight, out_conv bias) of a reference ConvRefiner."""
    blocks = [block_params(refiner.block1)] + [block_params(b) for b in refiner.hidden_blocks]
    return blocks, refiner.out_conv.weight.detach(), refiner.out_conv.bias.detach()


def block_port(p, h, round_fp16=False):
    """create_block applied to h [b,c,G,G] (model/network.py:528-531), eval-mode batch norm."""
    c = h.shape[1]
    h = F.conv2d(h, p["w1"].to(h.dtype), None if p["b1"] is None else p["b1"].to(h.dtype), stride=1, padding=2, groups=c)
    h = F.batch_norm(h, p["mean"].to(h.dtype), p["var"].to(h.dtype), p["gamma"].to(h.dtype), p["beta"].to(h.dtype), False, 0.0, p["eps"])
    h = F.relu(h)
    if round_fp16:
        h = h.half().to(h.dtype)
    w2 = p["w2"].half().to(h.dtype) if round_fp16 else p["w2"].to(h.dtype)
    h = F.conv2d(h, w2, p["b2"].to(h.dtype))
    if round_fp16:
        h = h.half().to(h.dtype)
    return h


def refiner_tail_port(blocks, wout, bout, d, dtype=torch.float32):
    """out_conv(hidden_blocks(block1(d)).float()) as the reference evaluates it on the CPU (fp32; float64 = the definition)."""
    h = d.to(dtype)
    for p in blocks:
        h = block_port(p, h)
    return F.conv2d(h, wout.to(dtype), bout.to(dtype))


def refiner_tail_fp16_storage(blocks, wout, bout, d):
    """The numerics of gfnet_b200's kernels modelled on the CPU in float64: activations rounded to fp16 where the kernels
    store them (refiner input, after ReLU, after the 1x1 convolution), 1x1 weights rounded to fp16, exact sums between."""
    h = d.half().double()
    for p in blocks:
        h = block_port(p, h, round_fp16=True)
    return F.conv2d(h, wout.double(), bout.double())


def flow_update_port(delta_flow, delta_certainty, flow, certainty, displacement_pre, scale, H0, W0, training=False):
    """One iteration of the refinement loop, model/network.py:265-274; returns (flow, certainty, displacement)."""
    displacement = int(scale) * torch.stack((delta_flow[:, 0].float() / (4 * W0), delta_flow[:, 1].float() / (4 * H0)), dim=1)
    if not training:
        displacement[((displacement - displacement_pre).abs() / (displacement_pre).abs()) < 1e-6] = 0
    return flow + displacement, certainty + delta_certainty, displacement


def upsample_port(x, size):
    """model/network.py:276-285."""
    return F.interpolate(x, size=size, mode="bilinear")
