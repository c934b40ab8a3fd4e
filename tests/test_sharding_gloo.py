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


# ---------------------------------------------------------------------------------------------------------------------
# step sizes of one line search sharded over ranks (libpddp: pddp_alpha_shard_*): host-side statement on CPU.
# Each rank runs the oracle's iteration but only "owns" the (J, defect) pairs of its alpha range; one all-gather (gloo here, NCCL in
# the library) and the reference's sequential scan on every rank must give the unsharded decision, iteration after iteration.
# ---------------------------------------------------------------------------------------------------------------------
def _alpha_worker(rank, world, port, ret):
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as ol
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, iters = 32, 6
    x0, u0, xg = pddp.make_inputs_kuka(N, 1, seed0=4)
    L = ol.lib(True); cfg = ol.kuka_cfg(N, fma=True); cfg.max_iter = iters; cp = C.byref(cfg)
    A = cfg.n_alpha; a0, cnt = sharding.alpha_range(rank, world, A)
    ok = True
    # unsharded oracle: the decisions to reproduce
    rJ = np.full(iters + 1, np.nan, np.float32); ra = np.full(iters + 1, -99, np.int32)
    rx = np.zeros((N, 14), np.float32); ru = np.zeros((N, 7), np.float32)
    L.orc_solve(cp, ol.fptr(x0[0]), ol.fptr(u0[0]), ol.fptr(xg[0]), ol.fptr(rx), ol.fptr(ru), ol.fptr(rJ), ol.iptr(ra))
    W = ol.WsView(L, cfg); Jout = np.full(iters + 1, np.nan, np.float32); aOut = np.full(iters + 1, -99, np.int32)
    L.orc_load(cp, W.ptr, ol.fptr(x0[0]), ol.fptr(u0[0]), ol.fptr(xg[0])); L.orc_init(cp, W.ptr, ol.fptr(Jout), ol.iptr(aOut))
    while True:
        L.orc_backward_pass(cp, W.ptr); L.orc_forward_sweep(cp, W.ptr); L.orc_forward_sim(cp, W.ptr); W.xp2[:] = W.xp
        L.orc_cost_defect(cp, W.ptr)
        # this rank only trusts its own candidates: everything else is poisoned, then filled by the exchange
        pairs = torch.tensor([[[W.s.J[a], W.s.dT[a]] for a in range(a0, a0 + cnt)]], dtype=torch.float32)
        for a in range(A):
            W.s.J[a] = float("nan"); W.s.dT[a] = float("nan")
        full = sharding.merge_selection_inputs(pairs)
        assert full.shape == (1, A, 2)
        for a in range(A):
            W.s.J[a] = float(full[0, a, 0]); W.s.dT[a] = float(full[0, a, 1])
        L.orc_line_search(cp, W.ptr)
        if L.orc_accept_reject(cp, W.ptr, ol.fptr(Jout), ol.iptr(aOut)):
            break
        L.orc_next_iteration_setup(cp, W.ptr)
    ok = ok and np.array_equal(aOut, ra) and np.array_equal(Jout, rJ, equal_nan=True) and np.array_equal(W.x[W.s.alphaIndex], rx)
    W.free()
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_alpha_ranges_tile_the_candidates():
    for world in (1, 2, 4, 8):
        r = [sharding.alpha_range(k, world, 16) for k in range(world)]
        assert r[0][0] == 0 and sum(c for _, c in r) == 16 and all(r[i][0] + r[i][1] == r[i+1][0] for i in range(world - 1))
    import pytest
    with pytest.raises(ValueError):
        sharding.alpha_range(0, 3, 16)


def test_alpha_sharded_selection_gloo():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_alpha_worker, args=(world, 29731 + os.getpid() % 200, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _uneven_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = 7                                                      # 4 + 3 problems: shards of different sizes
    lo, hi = sharding.shard_range(rank, world, G)
    allc = sharding.gather_counters(torch.arange(lo, hi, dtype=torch.int32) + 100)
    ret[rank] = bool(torch.equal(allc, torch.arange(G, dtype=torch.int32) + 100))
    dist.destroy_process_group()


def test_gather_counters_with_uneven_shards_gloo():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_uneven_worker, args=(world, 29931 + os.getpid() % 200, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
