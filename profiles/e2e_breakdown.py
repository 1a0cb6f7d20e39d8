"""Where the end-to-end step of DroneChaseVecEnv goes (host side): python profiles/e2e_breakdown.py [preset] [threads]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200.vec_env import DroneChaseVecEnv
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_vFinal"
E = 65536
for threads in ([int(sys.argv[2])] if len(sys.argv) > 2 else [4, 8, 16, 24]):
    v = DroneChaseVecEnv(name, n_envs=E, seed=1, terminal_observation=False, host_threads=threads)
    v.reset()
    rng = np.random.RandomState(0)
    acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
    for i in range(60): v.step(acts[i % 4])
    T = {"async": 0.0, "enqueue": 0.0, "sync": 0.0, "densify": 0.0, "rest": 0.0}
    n = 60
    t_all = time.perf_counter()
    for i in range(n):
        t0 = time.perf_counter(); v.step_async(acts[i % 4]); t1 = time.perf_counter()
        s = v.sim; v._flip ^= 1; h = v._h[v._flip]
        v._enqueue_obs(h); h["reward"].copy_(s.reward, non_blocking=True); h["done"].copy_(s.done, non_blocking=True); h["info"].copy_(s.info, non_blocking=True)
        t2 = time.perf_counter(); torch.cuda.current_stream(s.device).synchronize(); t3 = time.perf_counter()
        v._densify(h); t4 = time.perf_counter()
        T["async"] += t1 - t0; T["enqueue"] += t2 - t1; T["sync"] += t3 - t2; T["densify"] += t4 - t3
    tot = time.perf_counter() - t_all
    print(name, "threads", threads, "ms/step %.3f" % (tot / n * 1e3), {k: round(x / n * 1e3, 3) for k, x in T.items()}, "env-steps/s %.3e" % (E * n / tot))
    v.close()
