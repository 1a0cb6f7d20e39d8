"""Per-warp clock stamps of env_kernel's phases (build with -DDC_PROFILE_PHASES): python profiles/phase_clocks.py [preset]
Prints, per phase, the mean and the max over warps in microseconds at the nominal SM clock, and the slowest warps."""
import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import os
from dronechase_b200 import BatchedThreatEngageEnv, _lib
if os.environ.get('DC_LIB'): _lib.LIB_PATH = os.path.abspath(os.environ['DC_LIB'])      # e.g. build/libdc_phases.so
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_vFinal"
E = 65536
env = BatchedThreatEngageEnv(name, n_envs=E, seed=1234, device=0, sub_batches=1)
env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(1)
bank = torch.rand(8, E, 4, device='cuda', generator=g); bank[..., :3] = bank[..., :3] * 2 - 1
bank = [bank[i].contiguous() for i in range(8)]
L = _lib.lib()
NW = 2048
names = ["P0", "P3", "spawn", "P5", "P4"]
for i in range(160): env.step(bank[i % 8])
acc = []
for i in range(8):
    env.step(bank[i % 8]); torch.cuda.synchronize()
    buf = np.zeros((NW, 8), dtype=np.int64)
    rc = L.dc_debug_phase_clocks(buf.ctypes.data_as(C.c_void_p), NW); assert rc == 0
    d = np.diff(buf[:, :6], axis=1) / 1965.0     # us at 1965 MHz (clock64 counts SM cycles)
    if buf[:, 6].any():      # stamps inside P4: end of the slot pass (6), end of the projection (7)
        sub = np.stack([buf[:, 6] - buf[:, 4], buf[:, 7] - buf[:, 6], buf[:, 5] - buf[:, 7]], 1) / 1965.0
        if i == 7: print("P4 split (slot pass, projection, winners): mean", sub.mean(0).round(2), "max", sub.max(0).round(2))
    tot = (buf[:, 5] - buf[:, 0]) / 1965.0
    acc.append((d.mean(0), d.max(0), tot.mean(), tot.max(), (buf[:, 5].max() - buf[:, 0].min()) / 1965.0))
    if i == 7:
        worst = np.argsort(-tot)[:5]
        print("slowest warps:", [(int(w), [round(float(x), 1) for x in d[w]]) for w in worst])
m = np.mean([a[0] for a in acc], 0); mx = np.mean([a[1] for a in acc], 0)
for k, n in enumerate(names): print(f"{n:6s} mean {m[k]:7.2f} us   max over warps {mx[k]:7.2f} us")
print("warp total: mean %.2f us, max %.2f us" % (np.mean([a[2] for a in acc]), np.mean([a[3] for a in acc])))
