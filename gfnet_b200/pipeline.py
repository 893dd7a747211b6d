"""The hot path end to end for a batch of pairs: global match -> local correlation at every scale /
iteration / pass (as the refiner-input assembly of SURVEY.md 8 f1: grid features, warped features, displacement
embedding and the correlation written into one buffer) -> match post-process -> balanced sampling (kde) -> homography ->
corner error.

Order and shapes follow the reference's inference call stack: model/network.py:251-259
(coarse match, refiner's local_correlation per scale), :326-349 (560 pass), :358-414 (post-process,
sample), estimation.py:60-92 (H + metric).  The refiner's conv blocks between the local-correlation
calls are out of scope, so each call gets its flow from the input batch.
"""
import torch

from . import ops, matcher, estimation


class HotPath:
    def __init__(self, num_samples=5000, n_hyp=None, precision=0, seed=0):
        # n_hyp None = OpenCV's own RANSAC loop on the device (cv2.findHomography parity), K > 0 = fixed hash-drawn budget
        self.num_samples, self.n_hyp, self.precision, self.seed = num_samples, n_hyp, precision, seed
        self._corr = {}
        self.timing = None    # list of (key, start_event, end_event) when bench.py wants per-launch device times

    def _corr_buf(self, key, shape, device):
        buf = self._corr.get(key)
        if buf is None or buf.shape != shape or buf.device != device:
            buf = torch.empty(shape, device=device, dtype=torch.float32)
            self._corr[key] = buf
        return buf

    def kernel_launches(self, batch):
        """How many of our kernels one run() launches (bench.py's gpu_launches claim)."""
        n = 1                                                   # global match
        for p in batch.passes:                                  # refiner input: assemble + local correlation per iteration
            for sc in p:
                b = sc["f1"].shape[0]
                n += ops.refiner_input_launches(b, sc["c"], sc["hs"], sc["hs"], sc["G"], sc["r"], calls=len(sc["flows"]))
        n += 1 + 1 + 1 + 1 + 5 + 1 + 1 + 1                      # postprocess, keys, topk, gather, kde (keys, sort, gather+boxes, symmetric, finish), balance, topk, gather
        n += (1 if self.n_hyp is None else 2 if self.n_hyp > 0 else 0) + 1 + 1   # (cv ransac | init + ransac), refit, corner error
        return n

    def run(self, batch, generator=None, noise=None):
        out = self.run_front(batch)
        out.update(self.run_back(batch, generator=generator, noise=noise))
        return out

    def run_front(self, batch):
        """Dense part: coarse global match + the refiner input of every scale / iteration / pass (kernels that fill the GPU)."""
        out = {}
        out["coarse_flow"] = ops.coarse_match(batch.coarse_f0, batch.coarse_f1, precision=self.precision)
        for pi, scales in enumerate(batch.passes):
            for sc in scales:
                b, c, hs, G, r = sc["f1"].shape[0], sc["c"], sc["hs"], sc["G"], sc["r"]
                dd = sc["disp_w"].shape[0]
                # refiner input d = cat(grid_feature, x_hat, emb, local_corr) written in place (model/network.py:537-555);
                # the features of a scale are the same in every refiner iteration (:230-281), so the tcgen05 pre-pass of
                # the 64-channel scales is hoisted out of the iteration loop (and timed with the first call)
                buf = self._corr_buf((pi, sc["scale"]), (b, 2 * c + dd + (2 * r + 1) ** 2, G, G), sc["x"].device)
                prep = None
                for it, flow in enumerate(sc["flows"]):
                    args = (G, sc["x"], sc["f1"], flow, sc["disp_w"], sc["disp_b"], r, sc["scale_factor"])
                    ops.refiner_input(*args, out=buf, parts=1 if it == 0 else 5)     # grid features: first iteration only
                    if self.timing is not None:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                    if it == 0:
                        _, prep = ops.refiner_input(*args, out=buf, parts=2, want_prepared=True)
                    else:
                        ops.refiner_input(*args, out=buf, parts=2, prepared=prep)
                    if self.timing is not None:
                        e1.record()
                        self.timing.append((f"pass{pi + 1}_scale{sc['scale']}", e0, e1))
        out["last_corr"] = buf
        return out

    def run_back(self, batch, generator=None, noise=None):
        """Sparse tail: match post-process, balanced sampling (top-k, kde), homography, corner error -- per-pair kernels that
        leave most SMs idle (32 CTAs), so a throughput loop may run it on a second stream under the next batch's dense part."""
        out = {}
        warp, cert = matcher.match_postprocess(batch.final_flow, batch.cert_logits, symmetric=True)
        m, c = matcher.sample_batched(warp, cert, self.num_samples, generator=generator, noise=noise)
        res = batch.res
        H, status, ninl = estimation.estimate_homography(m, res, res, res, res, n_hyp=self.n_hyp, seed=self.seed)
        err = estimation.corner_error(H, batch.H_gt, res, res)
        out.update(H=H, status=status, n_inliers=ninl, err=err, matches=m)
        # [B,12] result row: 9 H entries + error + inlier count + status (what the ranks gather)
        out["result"] = torch.cat((H.reshape(-1, 9), err[:, None].double(), ninl[:, None].double(), status[:, None].double()), 1)
        return out
