import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200.vec_env import DroneChaseVecEnv
E=65536
v = DroneChaseVecEnv("level5_c1", n_envs=E, seed=1, terminal_observation=False)
v.reset()
rng = np.random.RandomState(0)
acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
for i in range(30): v.step(acts[i % 4])
T=[0,0,0,0]
for i in range(30):
    a=acts[i%4]
    t0=time.perf_counter(); v._h_actions.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
    t1=time.perf_counter(); v._dev_actions.copy_(v._h_actions, non_blocking=True)
    t2=time.perf_counter(); v.sim.step(v._dev_actions)
    t3=time.perf_counter(); v.step_wait(); t4=time.perf_counter()
    T[0]+=t1-t0; T[1]+=t2-t1; T[2]+=t3-t2; T[3]+=t4-t3
print("host copy %.3f  h2d enqueue %.3f  sim.step %.3f  step_wait %.3f ms" % tuple(x/30*1e3 for x in T))
