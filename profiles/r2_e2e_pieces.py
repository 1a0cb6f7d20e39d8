"""The pieces of DroneChaseVecEnv.step (change-list transfer) timed one by one: python profiles/r2_e2e_pieces.py [preset] [envs]"""
import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200.vec_env import DroneChaseVecEnv
from dronechase_b200 import _lib
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_v2_full"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rng = np.random.RandomState(0)
acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
v = DroneChaseVecEnv(name, n_envs=E, seed=1, terminal_observation=True, pairs_lidar=True, host_threads=16)
v.reset()
for i in range(160): v.step(acts[i % 4])
def t(label, fn, n=50, sync=True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
        if sync: torch.cuda.synchronize()
    print(f"  {label:70s} {(time.perf_counter() - t0) / n * 1e3:.3f} ms")
print(f"{name} E={E}")
t("np.copyto(actions -> pinned)", lambda i: np.copyto(v._h_actions_np, acts[i % 4]), sync=False)
t("H2D actions + sync", lambda i: v._dev_actions.copy_(v._h_actions, non_blocking=True))
t("dc_step enqueue (host time only)", lambda i: v.sim.step(v._dev_actions), sync=False)
t("dc_step + sync", lambda i: v.sim.step(v._dev_actions))
g_pairs, g_rest = v._graphs[0]
t("graph: diff kernel + head of the change list D2H, + sync", lambda i: g_pairs.replay())
t("graph: other obs + reward/done/info + terminal rows D2H, + sync", lambda i: g_rest.replay())
n_pairs = []
def one(i):
    v.sim.step(v._dev_actions); g_pairs.replay(); torch.cuda.synchronize(); n_pairs.append(int(v._pairs_h[0]))
t("dc_step + diff + head copy + sync", one)
n = int(np.median(n_pairs)); print(f"  pairs per step: median {n} ({n / E:.2f} per env), first copy holds {v._pairs_fast}")
dense = v._h[0]["obs"]["lidar"]
for th in (1, 4, 8, 16):
    t(f"dc_host_apply_pairs x{th} ({n} pairs)", lambda i: _lib.lib().dc_host_apply_pairs(C.c_void_p(dense.data_ptr()), C.c_void_p(v._pairs_h.data_ptr() + 8), n, th), sync=False)
t("np.nonzero(dones) + views", lambda i: np.nonzero(v._h[0]["done"].numpy().view(np.bool_))[0], sync=False)
t("full v.step", lambda i: v.step(acts[i % 4]), sync=False)
v.close()
