"""World-size-2 gloo test (CPU) of the N>1 host logic: shard ranges, rank-seeded synthetic pairs that
do not depend on the world size, and the single final gather of the [B,12] result rows."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gfnet_b200.dist import shard_range, gather_results
    from gfnet_b200 import synth
    G = 5
    lo, hi = shard_range(G, rank, world)
    rows = []
    for pair in range(lo, hi):          # inputs are seeded by global pair index: world-size invariant
        cg = torch.Generator().manual_seed(1234 + pair)
        H = synth.random_homography(cg)
        rows.append(np.concatenate((H.reshape(9), [float(pair), 1.0, 1.0])))
    local = torch.tensor(np.stack(rows), dtype=torch.float64)
    allr = gather_results(local, G)
    if rank == 0:
        q.put(allr.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    from gfnet_b200.dist import shard_range
    from gfnet_b200 import synth
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_range(256, r, 8) for r in range(8)][-1] == (224, 256)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == (5, 12)
    for pair in range(5):
        H = synth.random_homography(torch.Generator().manual_seed(1234 + pair))
        assert np.allclose(got[pair, :9], H.reshape(9)) and got[pair, 9] == pair
