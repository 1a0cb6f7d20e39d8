"""Is the numpy VecEnv at 8 GPUs bound by the HOST?  Run under torchrun with one rank per GPU.
 (1) device->pinned-host copies of 10 MB (one step's outputs of a 65,536-env batch), rank 0 alone and then all ranks at once:
     per-rank and aggregate GB/s;
 (2) host->device copies of 1 MB (one step's actions), the same way;
 (3) dc_host_apply_pairs of 133 k pairs into a 266 MB array on this rank's core slice, alone and all at once;
 (4) DroneChaseVecEnv.step alone on rank 0 and on all ranks at once.
Everything is timed on the host clock around a device synchronise, 100 repetitions after 10 of warm-up."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
from dronechase_b200.vec_env import DroneChaseVecEnv, _pin_to_rank_cores
from dronechase_b200 import _lib
import ctypes as C
_pin_to_rank_cores()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, n=100, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def phase(label, fn, unit_bytes=None):
    """rank 0 alone, then everybody"""
    out = {}
    barrier()
    alone = timed(fn) if rank == 0 else None
    barrier()
    allr = timed(fn)
    barrier()
    t = torch.tensor([allr], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out = {"alone_ms": alone * 1e3, "all_ranks_ms_max": float(t.item()) * 1e3}
        if unit_bytes:
            out["alone_GBps"] = unit_bytes / alone / 1e9
            out["all_ranks_aggregate_GBps"] = world * unit_bytes / float(t.item()) / 1e9
        print(label, json.dumps(out), flush=True)
    return out


n = 10 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
phase("d2h_10MB", lambda: (h.copy_(d, non_blocking=True), torch.cuda.synchronize()), n)
d1 = torch.empty(1 << 20, dtype=torch.uint8, device="cuda"); h1 = torch.empty(1 << 20, dtype=torch.uint8, pin_memory=True)
phase("h2d_1MB", lambda: (d1.copy_(h1, non_blocking=True), torch.cuda.synchronize()), 1 << 20)
E = 65536
dense = np.ones((E, 3, 13, 26), dtype=np.float32)
rng = np.random.RandomState(rank)
idx = rng.choice(dense.size, 133000, replace=False).astype(np.int32)
pairs = np.stack([idx, np.zeros_like(idx)], axis=1).copy()
threads = max(1, min(16, len(os.sched_getaffinity(0))))
L = _lib.lib()
phase(f"host_apply_133k_pairs_{threads}_threads", lambda: L.dc_host_apply_pairs(C.c_void_p(dense.ctypes.data), C.c_void_p(pairs.ctypes.data), len(idx), threads))
v = DroneChaseVecEnv("exp02_v2_full", n_envs=E, seed=1, device=local, env_offset=rank * E, terminal_observation=True)
v.reset()
acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
k = [0]
def step():
    v.step(acts[k[0] % 4]); k[0] += 1
for _ in range(60):
    step()
phase("vecenv_step", step)
if rank == 0:
    print("cores of rank 0:", sorted(os.sched_getaffinity(0)), "threads", v._threads, flush=True)
v.close()
if world > 1:
    dist.destroy_process_group()
