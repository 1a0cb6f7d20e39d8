"""Where DroneChaseVecEnv.step spends its wall time (host clock around each stage; the GPU stages show up as the waits):
python profiles/r2_e2e_where.py [preset] [envs] [mode: mapped|pairs]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200.vec_env import DroneChaseVecEnv
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_v2_full"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
mode = sys.argv[3] if len(sys.argv) > 3 else "mapped"
kw = dict(pairs_lidar=True, host_threads=8) if mode == "pairs" else dict(mapped_lidar=True)
rng = np.random.RandomState(0)
acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
v = DroneChaseVecEnv(name, n_envs=E, seed=1, terminal_observation=True, **kw)
v.reset()
for i in range(160): v.step(acts[i % 4])
T = {}
def timed(obj, attr, label):
    f = getattr(obj, attr)
    def w(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T[label] = T.get(label, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, attr, w)
timed(v, "step_async", "step_async (copy actions to pinned, H2D, dc_step enqueue)")
timed(v, "_enqueue_obs", "  _enqueue_obs (mirror / diff launch + obs D2H enqueue)")
timed(v, "_wait_and_densify", "  _wait_and_densify (wait for the GPU, host apply)")
timed(v, "step_wait", "step_wait (all of it)")
n = 100
dev = []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
for i in range(n):
    v.step(acts[i % 4])
tot = time.perf_counter() - t0
print(f"{name} E={E} {mode}: {tot / n * 1e3:.3f} ms per step")
for k, x in T.items(): print(f"  {k:70s} {x / n * 1e3:.3f} ms")
# GPU-side: the step alone, then step + transfers, by events
torch.cuda.synchronize()
for label, fn in (("dc_step only", lambda: v.sim.step(v._dev_actions)),):
    ev0.record()
    for i in range(50): fn()
    ev1.record(); torch.cuda.synchronize()
    print(f"  GPU {label}: {ev0.elapsed_time(ev1) / 50:.3f} ms")
v.close()
