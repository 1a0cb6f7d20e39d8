"""Launches every auxiliary kernel of the library once at a bench-like size, for an `ncu --set full` capture:
lidar_kernel, raycast_kernel (16,384 envs x 37 entities x 7 observers), scatter_hits / scatter_stack (DeviceRollout rebuilds),
diff_hits (numpy VecEnv change list), lw_obs_kernel (exp05: a wingman flown by a policy), stack_kernel (level5_c1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dronechase_b200 import BatchedThreatEngageEnv, DeviceRollout, lidar_project, lidar_raycast
from dronechase_b200.vec_env import DroneChaseVecEnv

SHORT = os.environ.get("AUX_SHORT") == "1"       # one or two launches per kernel: the ncu capture
dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(0)
E, N, O = 16384, 37, 7
pos = (torch.rand(E, N, 3, device=dev, generator=g) * 2 - 1) * 8
q = torch.randn(E, N, 4, device=dev, generator=g); q = q / q.norm(dim=-1, keepdim=True)
types = torch.tensor([3] * O + [1] * (N - O), dtype=torch.int32)
alive = (torch.rand(E, N, device=dev, generator=g) > 0.15).to(torch.uint8)
obs_slot = torch.arange(O, dtype=torch.int32)
for _ in range(1 if SHORT else 3):
    lidar_project(pos, q, types, alive, obs_slot, "fused", 40.0)
    lidar_raycast(pos, q, torch.full((N,), 0.15), types, alive, obs_slot, 40.0)


def rand_actions(n):
    a = torch.rand(n, 4, device=dev, generator=g); a[:, :3] = a[:, :3] * 2 - 1
    return a


for name, n in (("exp02_v2_full", 65536), ("level5_c1", 16384)):
    env = BatchedThreatEngageEnv(name, n_envs=n, seed=1, device=0, with_hits=True)
    env.reset()
    for _ in range(2 if SHORT else 40):
        env.step(rand_actions(n))
    ro = DeviceRollout(env, 4)
    for _ in range(1 if SHORT else 4):
        ro.add(rand_actions(n))
    for t in range(1 if SHORT else 4):
        ro.lidar(t)
    torch.cuda.synchronize(); env.close()
env = BatchedThreatEngageEnv("exp05_vFinal", n_envs=16384, seed=1, device=0)
env.reset()
for _ in range(2 if SHORT else 20):
    if hasattr(env, "lw_observe"):
        env.lw_observe()
    env.step(rand_actions(16384))
torch.cuda.synchronize(); env.close()
venv = DroneChaseVecEnv("exp02_v2_full", n_envs=65536, seed=1, device=0, transfer_graphs=False)
venv.reset()
rng = np.random.RandomState(0)
a = np.concatenate([rng.uniform(-1, 1, (65536, 3)), rng.uniform(0, 1, (65536, 1))], axis=1).astype(np.float32)
for _ in range(2 if SHORT else 12):
    venv.step(a)
venv.close()
print("aux kernels done")
