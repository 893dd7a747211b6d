"""The refinement loop of ``GFNet.forward`` on the device kernels (SURVEY.md 8 f1 + f4 + the loop glue).

reference: model/network.py:203-287 -- coarse global match at the first scale (or the upsampled previous result in the
560 pass), then per scale ``num_itr`` refiner iterations (input assembly + local correlation + convolution tail, displacement
scaling, eval-mode zeroing rule, flow / certainty update) and a bilinear upsampling to the next lattice.  The feature
pyramids come from the reference's own ``extract_features`` (the backbone is out of scope); everything after it runs here.
Inference only (eval mode, CUDA fp32 features); no fallback.
"""
import torch

from . import ops, refiner as RF
from .patch import _blocks_on_device


def refiner_delta(module, num_grid, x, y, flow, scale_factor=1, *, buf=None, handle=None, first=True):
    """One ``ConvRefiner.forward`` (model/network.py:533-564) as ``delta [B,3,G,G] = (dx, dy, d_certainty)``.

    ``buf`` / ``handle`` / ``first``: the iterations of one scale share the refiner-input buffer, the grid features and the
    feature pre-pass of the tcgen05 correlation kernel (the features do not change between them, :257-268).  Returns
    ``(delta, buf, handle)``."""
    r = module.local_corr_radius if module.corr_in_other else None
    w, b = module.disp_emb.weight, module.disp_emb.bias
    if first or buf is None:
        buf, handle = ops.refiner_input(num_grid, x, y, flow, w, b, r, scale_factor, want_prepared=True)
    else:
        ops.refiner_input(num_grid, x, y, flow, w, b, r, scale_factor, out=buf, prepared=handle, parts=1 | 2 | 4)
    rb = _blocks_on_device(module)
    if rb is None:
        raise NotImplementedError("decoder.refine covers the refiners GFNet builds (dw=True, kernel_size=5, BatchNorm2d, amp=True) "
                                  "in eval mode")
    return rb(buf), buf, handle


def refine(features0, features1, conv_refiner, num_grid, num_itr, H0, W0, *, scale_factor=1, upsample=False, pre_corresps=None,
           zero_rule=True):
    """``corresps`` of model/network.py:224-287 for feature pyramids ``{scale: [B,c,h,w]}`` (coarse to fine, keys as the
    reference's: "16", "8", "4", "2", "1"; without "16" in the upsample pass).  ``zero_rule`` = the reference's
    ``not self.training`` (:270-271); the refiner modules themselves must be in eval mode."""
    scales = list(features0.keys())
    corresps = {}
    flow = certainty = None
    for idx, scale in enumerate(scales):
        f0, f1 = features0[scale], features1[scale]
        if idx == 0:
            if upsample:
                if pre_corresps is None:
                    raise ValueError("you should provide a pre_corresps for upsampling refine.")
                flow = RF.upsample_bilinear(pre_corresps["flow"], num_grid[0])
                certainty = RF.upsample_bilinear(pre_corresps["certainty"], num_grid[0])
            else:
                flow = ops.coarse_match(f0, f1)                              # corr_volume + pos_embed (:251-252)
                certainty = torch.zeros((flow.shape[0], 1) + tuple(flow.shape[2:]), device=flow.device, dtype=flow.dtype)
        corresps[scale] = {}
        disp_pre = torch.full_like(flow, 1e-7)
        buf = handle = None
        for itr in range(num_itr[idx]):
            delta, buf, handle = refiner_delta(conv_refiner[scale], num_grid[idx], f0, f1, flow, scale_factor,
                                               buf=buf, handle=handle, first=itr == 0)
            flow, certainty = flow.clone(), certainty.clone()                # every iteration's result is kept in corresps
            RF.flow_update(delta, flow, certainty, disp_pre, int(scale), H0, W0, zero_rule=zero_rule)
            corresps[scale][itr + 1] = {"flow": flow, "certainty": certainty}
        if scale != "1":
            flow = RF.upsample_bilinear(flow, num_grid[idx + 1])
            certainty = RF.upsample_bilinear(certainty, num_grid[idx + 1])
    return corresps


class GraphedRefine:
    """``refine`` for fixed shapes captured in a CUDA graph: one replay instead of ~400 kernel launches issued from Python per
    pass -- what matters at the reference's own operating point of ONE pair per call (test.py), where the eager loop is bound by
    host launch overhead, not by the GPU.  Inputs are copied into static buffers, the returned ``corresps`` tensors are static too
    (valid until the next call)."""

    def __init__(self, features0, features1, conv_refiner, num_grid, num_itr, H0, W0, *, scale_factor=1, upsample=False,
                 pre_corresps=None, zero_rule=True):
        self.f0 = {s: t.detach().clone() for s, t in features0.items()}
        self.f1 = {s: t.detach().clone() for s, t in features1.items()}
        self.pre = None if pre_corresps is None else {k: v.detach().clone() for k, v in pre_corresps.items()}
        args = (self.f0, self.f1, conv_refiner, list(num_grid), list(num_itr), H0, W0)
        kw = dict(scale_factor=scale_factor, upsample=upsample, pre_corresps=self.pre, zero_rule=zero_rule)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up outside the capture: packs weights, sets kernel attributes
            for _ in range(2):
                refine(*args, **kw)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = refine(*args, **kw)

    def __call__(self, features0, features1, pre_corresps=None):
        for s in self.f0:
            self.f0[s].copy_(features0[s])
            self.f1[s].copy_(features1[s])
        if self.pre is not None:
            for k in self.pre:
                self.pre[k].copy_(pre_corresps[k])
        self.graph.replay()
        return self.out


def gfnet_forward(original):
    """Replacement for ``GFNet.forward`` (model/network.py:203-287): the reference's own ``extract_features``, then ``refine``.
    Refiners in training mode, CPU tensors or visualisation requests go to the original."""
    def forward(self, batch, symmetric=False, upsample=False, scale_factor=1, pre_corresps=None, visualization=False):
        im0, im1 = batch["im_A"], batch["im_B"]
        if not im0.is_cuda or visualization or any(m.training for m in self.conv_refiner.values()):
            return original(self, batch, symmetric=symmetric, upsample=upsample, scale_factor=scale_factor,
                            pre_corresps=pre_corresps, visualization=visualization)
        H0, W0 = int(im0.shape[2]), int(im0.shape[3])
        features0, features1 = self.extract_features(torch.cat([im0, im1], dim=0), upsample)
        if symmetric:                                                        # both directions in one op batch (:213-222)
            fq = {s: torch.cat((features0[s], features1[s]), dim=0) for s in features0.keys()}
            fs = {s: torch.cat((features1[s], features0[s]), dim=0) for s in features0.keys()}
            features0, features1 = fq, fs
        features0 = {s: t.float().contiguous() for s, t in features0.items()}
        features1 = {s: t.float().contiguous() for s, t in features1.items()}
        num_grid, num_itr = (self.num_grid_up, self.num_itr_up) if upsample else (self.num_grid, self.num_itr)
        return refine(features0, features1, self.conv_refiner, num_grid, num_itr, H0, W0, scale_factor=scale_factor,
                      upsample=upsample, pre_corresps=pre_corresps, zero_rule=not self.training)
    return forward


__all__ = ["refine", "refiner_delta", "gfnet_forward", "GraphedRefine"]
