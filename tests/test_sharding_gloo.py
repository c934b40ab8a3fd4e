"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard ranges tile the global batch, each rank's synthetic
inputs are exactly its slice of the global input set, and the end-of-solve gather returns counters in problem order."""
import importlib
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pddp = importlib.import_module("parallel-ddp_b200")
sharding = importlib.import_module("parallel-ddp_b200.sharding")


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, N = 6, 32
    lo, hi = sharding.shard_range(rank, world, G)
    x0, u0, xg = pddp.make_inputs_kuka(N, hi - lo, seed0=lo)
    gx, gu, gg = pddp.make_inputs_kuka(N, G, seed0=0)
    ok = np.array_equal(x0, gx[lo:hi]) and np.array_equal(u0, gu[lo:hi]) and np.array_equal(xg, gg[lo:hi])
    counters = torch.arange(lo, hi, dtype=torch.int32) * 10 + 3          # stand-in for per-problem iteration counters
    allc = sharding.gather_counters(counters)
    ok = ok and torch.equal(allc, torch.arange(G, dtype=torch.int32) * 10 + 3)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_ranges_tile_the_batch():
    for world in (1, 2, 4, 8):
        for G in (8, 64, 512, 13):
            r = [sharding.shard_range(k, world, G) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == G and all(r[i][1] == r[i+1][0] for i in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_sharding_gloo():
    if not os.path.exists(pddp.LIB_PATH):
        import subprocess
        subprocess.check_call([os.path.join(os.path.dirname(pddp.LIB_PATH), "build.sh")])
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29531 + os.getpid() % 200, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
