"""Refiner convolution blocks and decoder-loop glue on the device.

reference: ``ConvRefiner.create_block`` / the tail of ``ConvRefiner.forward`` (model/network.py:505-531, 557-563) and the
refinement loop of ``GFNet.decoder`` (model/network.py:262-285).  Host side only packs weights and owns buffers; the
arithmetic is in ``csrc/refiner_blocks.cu`` behind the C ABI.  CUDA only, no fallback.
"""
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr, require_cuda_f32


def pad16(c):
    return (int(c) + 15) // 16 * 16


@torch.no_grad()
def _fold_block(block, cp):
    """One ``create_block`` Sequential(conv1 dw 5x5, BatchNorm2d, ReLU, conv2 1x1) -> (wf [25,cp], shift [cp], b2 [cp],
    w2 [cp,cp] fp16).  Eval-mode batch norm folded in float64: y = conv(x) * s + (bias - mean) * s + beta."""
    conv1, norm, relu, conv2 = block[0], block[1], block[2], block[3]
    c = conv1.in_channels
    if not (isinstance(conv1, torch.nn.Conv2d) and conv1.groups == c and conv1.out_channels == c
            and conv1.kernel_size == (5, 5) and conv1.padding == (2, 2) and conv1.stride == (1, 1)
            and conv1.padding_mode == "zeros"):
        raise NotImplementedError("refiner blocks: conv1 must be a depth-wise 5x5 convolution (dw=True, kernel_size=5)")
    if not isinstance(norm, torch.nn.BatchNorm2d) or norm.training or norm.running_mean is None:
        raise NotImplementedError("refiner blocks: BatchNorm2d in eval mode with running statistics only")
    if not isinstance(relu, torch.nn.ReLU):
        raise NotImplementedError("refiner blocks: the activation must be ReLU")
    if not (isinstance(conv2, torch.nn.Conv2d) and conv2.kernel_size == (1, 1) and conv2.groups == 1
            and conv2.in_channels == c and conv2.out_channels == c):
        raise NotImplementedError("refiner blocks: conv2 must be a 1x1 convolution hidden -> hidden")
    dev = conv1.weight.device
    f64 = torch.float64
    s = (norm.weight.to(f64) if norm.affine else torch.ones(c, dtype=f64, device=dev)) / torch.sqrt(norm.running_var.to(f64) + norm.eps)
    beta = norm.bias.to(f64) if norm.affine else torch.zeros(c, dtype=f64, device=dev)
    b1 = conv1.bias.to(f64) if conv1.bias is not None else torch.zeros(c, dtype=f64, device=dev)
    wf = torch.zeros((25, cp), dtype=torch.float32, device=dev)
    wf[:, :c] = (conv1.weight.to(f64).reshape(c, 25) * s[:, None]).t().float()
    shift = torch.zeros(cp, dtype=torch.float32, device=dev)
    shift[:c] = ((b1 - norm.running_mean.to(f64)) * s + beta).float()
    b2 = torch.zeros(cp, dtype=torch.float32, device=dev)
    if conv2.bias is not None:
        b2[:c] = conv2.bias.float()
    w2 = torch.zeros((cp, cp), dtype=torch.float16, device=dev)
    w2[:c, :c] = conv2.weight.reshape(c, c).to(torch.float16)          # autocast casts the weights the same way
    return wf, shift, b2, w2


class RefinerBlocks:
    """The convolution tail of one ``ConvRefiner`` with its weights packed for ``gfb_refiner_blocks_f16``.

    ``RefinerBlocks.from_module(refiner)(d)`` == ``refiner.out_conv(refiner.hidden_blocks(refiner.block1(d)).float())`` of the
    reference under its fp16 autocast, up to fp16 rounding of the activations (sums are fp32 here)."""

    def __init__(self, blocks, out_conv):
        blocks = list(blocks)
        self.c = int(blocks[0][0].in_channels)
        self.cp = pad16(self.c)
        self.nblocks = len(blocks)
        self.out_dim = int(out_conv.out_channels)
        if not (isinstance(out_conv, torch.nn.Conv2d) and out_conv.kernel_size == (1, 1) and out_conv.in_channels == self.c
                and self.out_dim <= 4):
            raise NotImplementedError("refiner blocks: out_conv must be a 1x1 convolution hidden -> (<= 4)")
        dev = out_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("refiner blocks: the module must live on a CUDA device (no CPU path)")
        _lib.require_sm100()
        parts, self.folded = [], []
        with torch.no_grad():
            for blk in blocks:
                wf, shift, b2, w2 = _fold_block(blk, self.cp)
                self.folded.append((wf, shift, b2, w2))
                parts += [wf.reshape(-1).view(torch.uint8), shift.view(torch.uint8), b2.view(torch.uint8),
                          w2.reshape(-1).view(torch.uint8)]
            wout = torch.zeros((self.out_dim, self.cp), dtype=torch.float32, device=dev)
            wout[:, :self.c] = out_conv.weight.reshape(self.out_dim, self.c).float()
            bout = torch.zeros(4, dtype=torch.float32, device=dev)
            if out_conv.bias is not None:
                bout[:self.out_dim] = out_conv.bias.float()
            self.wout, self.bout = wout, bout
            parts += [wout.reshape(-1).view(torch.uint8), bout.view(torch.uint8)]
            self.blob = torch.cat(parts).contiguous()
        assert self.blob.numel() == int(lib.gfb_refiner_blocks_weight_bytes(self.c, self.nblocks, self.out_dim))
        self._ws = None

    @classmethod
    def from_module(cls, refiner):
        return cls([refiner.block1] + list(refiner.hidden_blocks), refiner.out_conv)

    def launches(self, B, G):
        chunk = int(lib.gfb_refiner_blocks_chunk(B, self.c, G))
        return ((B + chunk - 1) // chunk) * (2 + 2 * self.nblocks)

    def __call__(self, d, chunk=0, algo=0):
        d = require_cuda_f32("d", d)
        if d.dim() != 4 or d.shape[1] != self.c or d.shape[2] != d.shape[3]:
            raise ValueError(f"d must be [B,{self.c},G,G]")
        B, _, G, _ = (int(v) for v in d.shape)
        out = torch.empty((B, self.out_dim, G, G), device=d.device, dtype=torch.float32)
        nws = int(lib.gfb_refiner_blocks_workspace_bytes(B, self.c, G, int(chunk)))
        if self._ws is None or self._ws.numel() < nws or self._ws.device != d.device:
            self._ws = torch.empty(nws, device=d.device, dtype=torch.uint8)
        with torch.cuda.device(d.device):
            check(lib.gfb_refiner_blocks_f16(ptr(d), ptr(self.blob), ptr(out), B, self.c, G, self.nblocks, self.out_dim,
                                             ptr(self._ws), self._ws.numel(), int(chunk), int(algo), stream_ptr(d.device)),
                  "refiner_blocks")
        return out


def refiner_blocks(refiner, d):
    """``out_conv(hidden_blocks(block1(d)).float())`` of a reference ``ConvRefiner`` (weights packed once per module)."""
    rb = getattr(refiner, "_gfb_blocks", None)
    if rb is None or rb.blob.device != d.device:
        rb = RefinerBlocks.from_module(refiner)
        refiner._gfb_blocks = rb
    return rb(d)


# ---- single kernels (tests, benches) -----------------------------------------------------------------------------------
def pack_nhwc_f16(d):
    d = require_cuda_f32("d", d)
    B, C = int(d.shape[0]), int(d.shape[1])
    P = int(d.shape[2]) * int(d.shape[3])
    out = torch.empty((B, P, pad16(C)), device=d.device, dtype=torch.float16)
    with torch.cuda.device(d.device):
        check(lib.gfb_refiner_pack_f16(ptr(d), ptr(out), B, C, P, stream_ptr(d.device)), "refiner_pack")
    return out


def dw5_bn_relu(h, wf, shift, G):
    """h [B,G*G,Cp] fp16 NHWC -> relu(dwconv5x5(h) with folded batch norm) in the same layout."""
    B, P, cp = (int(v) for v in h.shape)
    if h.dtype != torch.float16 or not h.is_cuda or P != G * G:
        raise ValueError("h must be a CUDA fp16 [B,G*G,Cp] tensor")
    out = torch.empty_like(h)
    with torch.cuda.device(h.device):
        check(lib.gfb_refiner_dw5_f16(ptr(h.contiguous()), ptr(wf.contiguous()), ptr(shift.contiguous()), ptr(out), B, G, cp,
                                      stream_ptr(h.device)), "refiner_dw5")
    return out


def pointwise(act, w2, bias, algo=0):
    """act [..., Cp] fp16 -> fp16(act @ w2.T + bias) (the 1x1 convolution as a GEMM)."""
    cp = int(act.shape[-1])
    if act.dtype != torch.float16 or w2.dtype != torch.float16 or not act.is_cuda or w2.shape != (cp, cp):
        raise ValueError("act [...,Cp] fp16 and w2 [Cp,Cp] fp16 on CUDA expected")
    a = act.contiguous()
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        check(lib.gfb_refiner_pw_f16(ptr(a), ptr(w2.contiguous()), ptr(bias.contiguous()), ptr(out), a.numel() // cp, cp, int(algo),
                                     stream_ptr(a.device)), "refiner_pw")
    return out


def out_conv(act, w, bias, B, G):
    cp = int(act.shape[-1])
    oc = int(w.shape[0])
    out = torch.empty((B, oc, G, G), device=act.device, dtype=torch.float32)
    with torch.cuda.device(act.device):
        check(lib.gfb_refiner_out_f32(ptr(act.contiguous()), ptr(w.contiguous()), ptr(bias.contiguous()), ptr(out), B, G * G, cp, oc,
                                      stream_ptr(act.device)), "refiner_out")
    return out


def flow_update(delta, flow, certainty, disp_pre, scale, H0, W0, zero_rule=True):
    """In place: the body of the refinement loop, model/network.py:265-274 (``delta`` = cat(delta_flow, delta_certainty))."""
    delta = require_cuda_f32("delta", delta)
    for name, t in (("flow", flow), ("certainty", certainty), ("disp_pre", disp_pre)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous CUDA fp32 tensor (updated in place)")
    B, _, G, _ = (int(v) for v in flow.shape)
    if delta.shape != (B, 3, G, G) or certainty.shape != (B, 1, G, G) or disp_pre.shape != flow.shape:
        raise ValueError("delta [B,3,G,G], flow [B,2,G,G], certainty [B,1,G,G], disp_pre [B,2,G,G] expected")
    with torch.cuda.device(flow.device):
        check(lib.gfb_flow_update_f32(ptr(delta), ptr(flow), ptr(certainty), ptr(disp_pre), B, G, int(scale), int(H0), int(W0),
                                      int(bool(zero_rule)), stream_ptr(flow.device)), "flow_update")
    return flow, certainty


def upsample_bilinear(x, size):
    """``F.interpolate(x, size=size, mode="bilinear", align_corners=False)`` for [B,C,H,W] fp32 (model/network.py:276-285)."""
    x = require_cuda_f32("x", x)
    ho, wo = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    B, C, hi, wi = (int(v) for v in x.shape)
    out = torch.empty((B, C, ho, wo), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib.gfb_upsample_bilinear_f32(ptr(x), ptr(out), B * C, hi, wi, ho, wo, stream_ptr(x.device)), "upsample_bilinear")
    return out


def refiner_blocks_flops(B, C, G, nblocks=9):
    """Useful FLOPs of one tail: per block 2*25 (depth-wise) + 2*C (1x1) per channel and pixel."""
    return nblocks * B * G * G * C * (50 + 2 * C)


def refiner_blocks_bytes(B, C, G, out_dim=3):
    """Algorithmic HBM bytes of one tail when the activations stay on chip: read d once, write the output planes once."""
    return 4 * B * G * G * (C + out_dim)


__all__ = ["RefinerBlocks", "refiner_blocks", "pack_nhwc_f16", "dw5_bn_relu", "pointwise", "out_conv", "flow_update",
           "upsample_bilinear", "refiner_blocks_flops", "refiner_blocks_bytes", "pad16"]
