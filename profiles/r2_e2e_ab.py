"""A/B of the sphere transfers of DroneChaseVecEnv in ONE process, alternating blocks of steps (a shared box's host speed drifts):
python profiles/r2_e2e_ab.py [preset] [envs]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200.vec_env import DroneChaseVecEnv
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_v2_full"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rng = np.random.RandomState(0)
acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
modes = {"mapped": dict(mapped_lidar=True), "pairs x16 eager": dict(pairs_lidar=True, host_threads=16, transfer_graphs=False),
         "pairs x4": dict(pairs_lidar=True, host_threads=4), "pairs x8": dict(pairs_lidar=True, host_threads=8),
         "pairs x16": dict(pairs_lidar=True, host_threads=16)}
envs = {}
for k, kw in modes.items():
    envs[k] = DroneChaseVecEnv(name, n_envs=E, seed=1, terminal_observation=True, **kw)
    envs[k].reset()
    for i in range(160): envs[k].step(acts[i % 4])
res = {k: [] for k in modes}
for rep in range(6):
    for k, v in envs.items():
        t0 = time.perf_counter()
        for i in range(40): v.step(acts[i % 4])
        res[k].append((time.perf_counter() - t0) / 40 * 1e3)
for k, x in res.items():
    print(f"{name} E={E} {k:16s} median {np.median(x):.3f} ms/step  blocks {[round(t, 3) for t in x]}  -> {E / np.median(x) * 1e3:.3e} env-steps/s")
