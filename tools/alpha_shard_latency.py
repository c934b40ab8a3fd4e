"""Latency of ONE problem (Kuka N=128, alpha=16) per iLQR iteration with the line search on 1 GPU and sharded over all visible GPUs:
python tools/alpha_shard_latency.py   (spawns one process per GPU; prints one JSON line)"""
import importlib, json, os, sys, time
import numpy as np
import torch, torch.distributed as dist, torch.multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pddp = importlib.import_module("parallel-ddp_b200")
N, B, IT = 128, int(os.environ.get("PDDP_LAT_BATCH", "1")), 50


def timed(s, x0, u0, xg, reps=5):
    s.runiLQR_GPU(x0, u0, xg); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); o = s.runiLQR_GPU(x0, u0, xg); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3, o


def worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x0, u0, xg = pddp.make_inputs_kuka(N, B, 0)
    uid = [pddp.alpha_shard_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
    s = pddp.Solver(pddp.default_config_kuka(N, B, device=rank, max_iter=IT)); s.alpha_shard_init(rank, world, uid[0])
    dist.barrier(); ms, o = timed(s, x0, u0, xg)
    ret[rank] = (ms, s.alpha_shard_stats()["exchange_us_per_iteration"], float(o["Jout"][0, IT]))
    s.freeMemory_GPU(); dist.destroy_process_group()


if __name__ == "__main__":
    x0, u0, xg = pddp.make_inputs_kuka(N, B, 0)
    out = {"problem": f"Kuka N={N} alpha=16 batch={B}, {IT} iterations, host buffers", "per_gpus": {}}
    s1 = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=IT)); ms1, o1 = timed(s1, x0, u0, xg); s1.freeMemory_GPU()
    out["per_gpus"]["1"] = {"ms_per_iteration": ms1 / IT, "exchange_us_per_iteration": 0.0}
    for world in [w for w in (2, 4, 8) if w <= torch.cuda.device_count()]:
        ret = mp.Manager().dict()
        mp.spawn(worker, args=(world, 29431 + world, ret), nprocs=world, join=True)
        out["per_gpus"][str(world)] = {"ms_per_iteration": max(ret[r][0] for r in range(world)) / IT, "exchange_us_per_iteration": max(ret[r][1] for r in range(world)),
                                       "same_final_cost_as_1_gpu": all(ret[r][2] == float(o1["Jout"][0, IT]) for r in range(world))}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open(f"gpurun_out/alpha_shard_latency_b{B}.json", "w"), indent=1)
