"""Rebind the reference's hot-path names to the B200 kernels (drop-in switch).

The reference binds with ``from utils.kde import kde`` / ``from utils.local_correlation import
local_correlation`` (model/network.py:10-11), so the *module attributes of model.network* must be
replaced, not only utils.*; ``GFNet.corr_volume`` / ``pos_embed`` / ``sample`` are methods.
"""
import torch

from . import ops, matcher, refiner as _refiner


def _blocks_on_device(module):
    """The packed convolution tail of a ConvRefiner, or None when the module is not the configuration the kernels cover
    (dw=True, kernel_size=5, BatchNorm2d in eval mode, fp16 / bf16 autocast): decided once per module and mode."""
    w0, w1 = module.block1[0].weight, module.out_conv.weight          # load_state_dict / optimiser steps bump _version in place
    key = (module.training, bool(module.amp), w0.data_ptr(), w1.data_ptr(),
           None if w0.is_inference() else w0._version, None if w1.is_inference() else w1._version)
    cached = getattr(module, "_gfb_blocks_state", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    rb = None
    if not module.training and module.amp and module.amp_dtype in (torch.float16, torch.bfloat16):
        try:
            rb = _refiner.RefinerBlocks.from_module(module)
        except NotImplementedError:
            rb = None
    module._gfb_blocks_state = (key, rb)
    return rb


def _refiner_forward(original, blocks=True):
    """ConvRefiner.forward (model/network.py:533-564) with lines 537-555 -- the two grid_samples, the displacement
    embedding, local_correlation and the concatenation -- replaced by ``ops.refiner_input`` (one buffer, written in
    place) and, with ``blocks``, the convolution tail (:557-563) by ``refiner.RefinerBlocks`` (fp16
    activations like the reference's autocast, fp32 sums); otherwise the tail stays the reference's own modules under the
    reference's autocast."""
    def forward(self, num_grid, x, y, flow, scale_factor=1, logits=None):
        if not (self.has_displacement_emb and self.sample_mode == "bilinear" and x.is_cuda
                and x.dtype == torch.float32 and y.dtype == torch.float32 and flow.dtype == torch.float32):
            return original(self, num_grid, x, y, flow, scale_factor=scale_factor, logits=logits)
        r = self.local_corr_radius if self.corr_in_other else None            # scale 1: no correlation channels (:557-558)
        d = ops.refiner_input(num_grid, x, y, flow, self.disp_emb.weight, self.disp_emb.bias, r, scale_factor)
        local_corr = d[:, d.shape[1] - (2 * r + 1) ** 2:] if r is not None else None
        rb = _blocks_on_device(self) if blocks else None
        if rb is not None:
            h = rb(d)
            return h[:, :2], h[:, 2:3], local_corr
        with torch.autocast("cuda", enabled=bool(self.amp), dtype=self.amp_dtype):
            h = self.block1(d)
            h = self.hidden_blocks(h)
        h = self.out_conv(h.float())
        return h[:, :2], h[:, 2:3], local_corr
    return forward


def patch(network_module, utils_local_correlation=None, utils_kde=None, refiner=True, refiner_blocks=True, forward=False):
    """``patch(model.network)``; returns a dict of the originals so ``unpatch`` can restore them.  ``refiner=True`` also
    replaces the input assembly of ``ConvRefiner.forward``, ``refiner_blocks=True`` its convolution tail
    (inference with the reference's autocast only, anything else keeps the reference's modules);
    ``forward=True`` additionally replaces ``GFNet.forward`` by ``decoder.gfnet_forward`` (the reference's backbone, then the
    whole refinement loop of model/network.py:224-287 on the device kernels: iterations of a scale share the refiner-input
    buffer, flow update and between-scale upsampling are kernels)."""
    saved = {
        "refiner_forward": network_module.ConvRefiner.forward,
        "local_correlation": network_module.local_correlation,
        "kde": network_module.kde,
        "corr_volume": network_module.GFNet.corr_volume,
        "pos_embed": network_module.GFNet.pos_embed,
        "sample": network_module.GFNet.sample,
        "forward": network_module.GFNet.forward,
    }
    network_module.local_correlation = ops.local_correlation
    network_module.kde = ops.kde
    network_module.GFNet.corr_volume = lambda self, feat0, feat1: ops.corr_volume(feat0, feat1)
    network_module.GFNet.pos_embed = lambda self, corr_volume: ops.pos_embed(corr_volume)

    def _sample(self, matches, certainty, num=5000):
        return matcher.sample(matches, certainty, num, sample_mode=self.sample_mode, sample_thresh=self.sample_thresh)
    network_module.GFNet.sample = _sample
    if refiner:
        network_module.ConvRefiner.forward = _refiner_forward(saved["refiner_forward"], blocks=refiner_blocks)
    if forward:
        from .decoder import gfnet_forward
        network_module.GFNet.forward = gfnet_forward(saved["forward"])
    if utils_local_correlation is not None:
        utils_local_correlation.local_correlation = ops.local_correlation
    if utils_kde is not None:
        utils_kde.kde = ops.kde
    return saved


def unpatch(network_module, saved):
    network_module.local_correlation = saved["local_correlation"]
    network_module.kde = saved["kde"]
    network_module.GFNet.corr_volume = saved["corr_volume"]
    network_module.GFNet.pos_embed = saved["pos_embed"]
    network_module.GFNet.sample = saved["sample"]
    network_module.ConvRefiner.forward = saved["refiner_forward"]
    if "forward" in saved:
        network_module.GFNet.forward = saved["forward"]
