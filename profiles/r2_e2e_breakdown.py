"""Where the end-to-end step goes with the mapped sphere transfer: python profiles/r2_e2e_breakdown.py [preset] [envs]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200.vec_env import DroneChaseVecEnv
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_vFinal"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rng = np.random.RandomState(0)
acts = [np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32) for _ in range(4)]
v = DroneChaseVecEnv(name, n_envs=E, seed=1, terminal_observation=True)
v.reset()
for i in range(160): v.step(acts[i % 4])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
T = {"host step_async": 0, "gpu h2d+step": 0, "gpu mirror": 0, "gpu d2h": 0, "host wait": 0, "host terminal": 0, "total": 0}
n = 60
for i in range(n):
    t0 = time.perf_counter()
    ev[0].record(); v.step_async(acts[i % 4]); ev[1].record()
    t1 = time.perf_counter()
    s = v.sim; v._flip ^= 1; h = v._h[v._flip]
    if v.mapped:
        import ctypes as C
        from dronechase_b200 import _lib
        f = v._flip
        _lib.check(_lib.lib().dc_mirror_hits(C.c_void_p(v._shown[f].data_ptr()), C.c_void_p(s.lidar_hits.data_ptr()), E, v.cfg.n_drones, v.cfg.n_lw,
                                             v.cfg.lidar_channels, C.c_void_p(v._dense_dev[f]), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "m")
    ev[2].record()
    for k, t in s.obs.items():
        if k != v._lidar_key: h["obs"][k].copy_(t, non_blocking=True)
    h["reward"].copy_(s.reward, non_blocking=True); h["done"].copy_(s.done, non_blocking=True); h["info"].copy_(s.info, non_blocking=True)
    ev[3].record()
    t2 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t3 = time.perf_counter()
    dones = h["done"].numpy().view(np.bool_)
    idx = np.nonzero(dones)[0]; m = len(idx)
    if m:
        didx = v._done_idx[:m]; didx.copy_(torch.from_numpy(idx), non_blocking=True)
        for k, t in s.terminal_obs.items(): v._h_term[k][:m].copy_(t.index_select(0, didx), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t4 = time.perf_counter()
    T["host step_async"] += t1 - t0; T["host wait"] += t3 - t2; T["host terminal"] += t4 - t3; T["total"] += t4 - t0
    T["gpu h2d+step"] += ev[0].elapsed_time(ev[1]) * 1e-3; T["gpu mirror"] += ev[1].elapsed_time(ev[2]) * 1e-3; T["gpu d2h"] += ev[2].elapsed_time(ev[3]) * 1e-3
print(name, E, "mapped" if v.mapped else "host scatter", {k: round(x / n * 1e3, 3) for k, x in T.items()}, "ms per step")
v.close()
