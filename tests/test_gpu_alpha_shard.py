"""Line search sharded over GPUs (north_star): needs at least two devices (run with `gpurun --gpus 2`); skipped otherwise.
Every rank solves the same problems with its share of the step sizes; results must equal the single-GPU solve bit for bit."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
pddp = importlib.import_module("parallel-ddp_b200")


def _worker(rank, world, port, ret, N, B, iters, tol):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x0, u0, xg = pddp.make_inputs_kuka(N, B, seed0=17)
    uid = [pddp.alpha_shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s = pddp.Solver(pddp.default_config_kuka(N, B, device=rank, max_iter=iters, tol_cost=tol))
    s.alpha_shard_init(rank, world, uid[0])
    st = s.alpha_shard_stats()
    assert (st["a_first"], st["a_cnt"]) == (rank * 16 // world, 16 // world)
    o = s.runiLQR_GPU(x0, u0, xg)
    o2 = s.runiLQR_GPU(x0, u0, xg)                               # a second solve on the same communicator
    ref = pddp.Solver(pddp.default_config_kuka(N, B, device=rank, max_iter=iters, tol_cost=tol)).runiLQR_GPU(x0, u0, xg)
    ok = all(np.array_equal(o[k], ref[k], equal_nan=True) and np.array_equal(o2[k], ref[k], equal_nan=True) for k in ("x", "u", "Jout", "alphaOut", "iters"))
    ret[rank] = (bool(ok), s.alpha_shard_stats()["exchange_us_per_iteration"], [int(v) for v in o["iters"]])
    s.freeMemory_GPU()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,B,iters,tol", [(32, 3, 20, 0.0), (128, 1, 12, 0.0), (32, 5, 40, 1e-4)])
def test_alpha_sharded_solve_equals_single_gpu(N, B, iters, tol):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2 if world < 4 else 4
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29331 + os.getpid() % 200, ret, N, B, iters, tol), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), dict(ret)
    print("exchange us/iteration per rank:", [round(ret[r][1], 1) for r in range(world)], "iters", ret[0][2])
