"""Evidence for a library problem met while testing the drop-in (none of gfnet_b200's kernels is launched here): after the
reference's own refinement loop has run once at another resolution, the reference's own ConvRefiner (model/network.py:533-564,
fp16 autocast) returns NON-FINITE values for finite inputs at [2, 417, 32, 32]; the first bad tensor is the output of the
depth-wise 5x5 convolution (cuDNN), and the same call with ``torch.backends.cudnn.enabled = False`` is finite.  The drop-in
tests therefore evaluate their reference arm with cuDNN disabled (PyTorch's native convolution kernels).

    python tools/repro_reference_depthwise_nan.py          # needs the staged reference (baseline/_ref)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import reference as R          # noqa: E402
import test_dropin_reference as TR         # noqa: E402
import test_decoder_dropin as TD           # noqa: E402

print("torch", torch.__version__, "cudnn", torch.backends.cudnn.version(), torch.cuda.get_device_name(0))
ref = R.load_reference()
warm = len(sys.argv) < 2 or sys.argv[1] != "cold"
if warm:   # the reference's own GFNet.forward (unpatched) on a 224-pixel synthetic pyramid
    s = TD._stand_in(ref, 224, [1, 1, 1, 1, 1], seed=11)
    batch = {"im_A": torch.zeros((1, 3, 224, 224), device="cuda"), "im_B": torch.zeros((1, 3, 224, 224), device="cuda")}
    with torch.inference_mode():
        ref.network.GFNet.forward(s, batch, symmetric=True)
    print("ran the reference's GFNet.forward once (unpatched, torch operators only)")
torch.manual_seed(16)
cr = R.make_conv_refiner(ref, 16).cuda().eval()
x, y, flow, G = TR._inputs(16, 2, 116)


def mk(name):
    def hook(mod, inp, out):
        i = inp[0]
        print(f"  {name:18s} input finite={bool(torch.isfinite(i.float()).all())} |max|={float(i.float().nan_to_num().abs().max()):.3g} ({i.dtype})"
              f"  output finite={bool(torch.isfinite(out.float()).all())} |max|={float(out.float().nan_to_num().abs().max()):.3g}")
    return hook


cr.block1[0].register_forward_hook(mk("block1 depth-wise"))
cr.block1[3].register_forward_hook(mk("block1 1x1"))
cr.out_conv.register_forward_hook(mk("out_conv"))
for enabled in (True, False, True):
    with torch.backends.cudnn.flags(enabled=enabled), torch.inference_mode():
        d, c, _ = cr(G, x, y, flow)
    print(f"cudnn enabled={enabled}: reference ConvRefiner.forward output finite={bool(torch.isfinite(d.float()).all())}")
