"""CPU simulation of shared-memory wavefronts per window-read instruction for candidate layouts of the C = 16
point kernel (no GPU needed).  python tools/sim/bank_sim.py"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfnet_b200 import synth

def geom(flow, hs, R):
    sx = ((flow[:, 0].double() + 1) * hs - 1) / 2
    sy = ((flow[:, 1].double() + 1) * hs - 1) / 2
    return (torch.floor(sx).long() - R).numpy(), (torch.floor(sy).long() - R).numpy()

def wavefronts(addr):
    """addr [n_instr, lanes] word addresses (-1 = inactive) -> wavefronts per instruction."""
    out = np.zeros(addr.shape[0], dtype=np.int64)
    for i, a in enumerate(addr):
        a = np.unique(a[a >= 0])
        if a.size == 0: continue
        out[i] = np.bincount(a % 32, minlength=32).max()
    return out

def sim_planes(xb, yb, TY, pitch_fn, NPL=3, W=6, rowrot=False):
    b, G, _ = xb.shape
    tot = cnt = 0
    for e in range(b):
        for ty in range(0, G, TY):
            for tx in range(0, G, 32):
                X = xb[e, ty:ty+TY, tx:tx+32]; Y = yb[e, ty:ty+TY, tx:tx+32]
                X0 = X.min() & ~3; Y0 = Y.min()
                shear = np.mean(Y[:, -1] - Y[:, 0])
                pitch = pitch_fn(shear)
                for w in range(TY):
                    u = X[w] - X0; oy = Y[w] - Y0
                    uq, um = u // NPL, u % NPL
                    instr = []
                    for j in range(W):
                        for q in range(NPL):
                            for s in range(2):
                                if rowrot:
                                    row = oy + ((j - oy) % W)
                                else:
                                    row = oy + j
                                instr.append(row * pitch + q * 32 + uq + (q < um) + s)
                    wf = wavefronts(np.array(instr))
                    tot += wf.sum(); cnt += len(instr)
    return tot / cnt

def sim_dense64(xb, yb, W=6):
    """old kernel: 16x8 tile, LDS.64 at aligned x, pitch 64, per half-warp"""
    b, G, _ = xb.shape
    tot = cnt = 0
    for e in range(b):
        for ty in range(0, G, 8):
            for tx in range(0, G, 16):
                X = xb[e, ty:ty+8, tx:tx+16]; Y = yb[e, ty:ty+8, tx:tx+16]
                X0 = X.min() & ~3; Y0 = Y.min()
                for w in range(8):
                    ox = (X[w] - X0) & ~1; oy = Y[w] - Y0
                    for j in range(W):
                        for h in range(4):
                            a = (oy + j) * 64 + ox + 2 * h
                            addr = np.concatenate([a, a + 1])
                            wf = wavefronts(addr[None])[0]
                            tot += wf; cnt += 1
    return tot / cnt   # wavefronts per half-warp LDS.64 (ideal 1)

gen = torch.Generator().manual_seed(0)
cgen = torch.Generator().manual_seed(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 6
Hn = [synth.random_homography(cgen) for _ in range(nb)]
Hs = Hn + [np.linalg.inv(h) for h in Hn]
flow = synth.homography_flow(Hs, 128, 224, gen, "cpu")
xb, yb = geom(flow, 224, 2)
print("dense LDS.64 (old), wavefronts per half-warp (ideal 1):", sim_dense64(xb, yb))
print("planes pitch 96:", sim_planes(xb, yb, 8, lambda s: 96))
for s in (1, 2, 3, 4, 5, 6, 7):
    print(f"planes pitch +-{s} adaptive:", sim_planes(xb, yb, 8, lambda sh, s=s: 96 + s if sh >= 0 else 128 - s),
          "fixed +:", sim_planes(xb, yb, 8, lambda sh, s=s: 96 + s))
print("planes pitch 96 row-rotation:", sim_planes(xb, yb, 8, lambda s: 96, rowrot=True))
for pp in (96, 97, 98, 99, 100, 101, 102, 104, 107, 112):
    print(f"row-rotation pitch {pp}:", sim_planes(xb, yb, 8, lambda s, pp=pp: pp, rowrot=True))
