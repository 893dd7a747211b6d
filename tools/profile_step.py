"""One warm step + one profiled step of the hot path (for ncu --profile-from-start off).
Usage: ncu ... --profile-from-start off python tools/profile_step.py [--pairs 32] [--itr 2]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfnet_b200 import synth
from gfnet_b200.pipeline import HotPath


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--itr", type=int, default=2)
    ap.add_argument("--warm", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    batch = synth.PairBatch(args.pairs, num_itr=args.itr, seed=1234, device=dev)
    hp = HotPath()
    for _ in range(args.warm):
        hp.run(batch, generator=gen)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    hp.run(batch, generator=gen)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
