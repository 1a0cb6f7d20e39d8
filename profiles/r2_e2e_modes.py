"""End-to-end step of DroneChaseVecEnv (numpy in / numpy out, terminal observations on) with the three sphere transfers:
python profiles/r2_e2e_modes.py [preset] [envs]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200.vec_env import DroneChaseVecEnv
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_vFinal"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rng = np.random.RandomState(0)
acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
import os
NOINFO = os.environ.get("NO_TERMINAL_DICTS") == "1"      # time v.step alone (bench.py's e2e arm does not build the lazy dicts either)
for label, kw in (("mapped (dc_mirror_hits)", dict(mapped_lidar=True)), ("change list x16 (dc_diff_hits)", dict(pairs_lidar=True, host_threads=16)),
                  ("change list x8", dict(pairs_lidar=True, host_threads=8)), ("change list x4", dict(pairs_lidar=True, host_threads=4)),
                  ("hit list + host scatter x16", dict(mapped_lidar=False, host_threads=16)),
                  ("hit list + host scatter x4", dict(mapped_lidar=False, host_threads=4)), ("dense D2H", dict(sparse_lidar=False))):
    if name.startswith("level5") and (kw.get("mapped_lidar") or kw.get("pairs_lidar")):
        continue
    v = DroneChaseVecEnv(name, n_envs=E, seed=1, terminal_observation=True, **kw)
    v.reset()
    for i in range(160): v.step(acts[i % 4])
    n, n_done, t_info = 60, 0, 0.0
    t0 = time.perf_counter()
    for i in range(n):
        obs, rew, dones, infos = v.step(acts[i % 4])
        t1 = time.perf_counter()
        for j in (() if NOINFO else np.nonzero(dones)[0]):
            infos[int(j)]["terminal_observation"]; n_done += 1
        t_info += time.perf_counter() - t1
    tot = time.perf_counter() - t0
    print(f"{name} E={E} {label:32s} {tot / n * 1e3:.3f} ms/step  {E * n / tot:.3e} env-steps/s  "
          f"(terminal dicts: {n_done / n:.0f} per step, {t_info / n * 1e3:.3f} ms)", flush=True)
    v.close()
