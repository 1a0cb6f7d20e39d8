"""Launch timeline of the step kernels (build with -DDC_PROFILE_PHASES, DC_LIB=build/libdc_phases.so):
python profiles/timeline.py [preset] [envs] [K] -> per step, start/end of every dyn_kernel / env_kernel launch in us."""
import os, sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
from dronechase_b200 import BatchedThreatEngageEnv, _lib
if os.environ.get('DC_LIB'): _lib.LIB_PATH = os.path.abspath(os.environ['DC_LIB'])
name = sys.argv[1] if len(sys.argv) > 1 else "exp02_v2_full"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
env = BatchedThreatEngageEnv(name, n_envs=E, seed=1234, device=0, sub_batches=K)
env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(1)
bank = torch.rand(8, E, 4, device='cuda', generator=g); bank[..., :3] = bank[..., :3] * 2 - 1
bank = [bank[i].contiguous() for i in range(8)]
L = _lib.lib()
for i in range(160): env.step(bank[i % 8])
torch.cuda.synchronize()
buf = np.zeros((256, 4), dtype=np.int64)
L.dc_debug_timeline(buf.ctypes.data_as(C.c_void_p), 1)
NS = 24
for i in range(NS): env.step(bank[i % 8])
torch.cuda.synchronize()
L.dc_debug_timeline(buf.ctypes.data_as(C.c_void_p), 0)
rows = buf[:NS * K * 2]
t0 = rows[:, 2].min()
sims = sorted(set(rows[:, 0].tolist()))
per_step = rows.reshape(NS, K * 2, 4)
prev_end = None
durs = []
for st in per_step[4:]:
    s0 = st[:, 2].min(); e1 = st[:, 3].max()
    line = " ".join(f"{'dyn' if r[1] == 0 else 'env'}{sims.index(r[0])}[{(r[2]-s0)/1e3:5.1f},{(r[3]-s0)/1e3:5.1f}]" for r in st)
    gap = (s0 - prev_end) / 1e3 if prev_end is not None else 0.0
    print(f"step span {(e1-s0)/1e3:6.1f} us, gap before {gap:5.1f} us | {line}")
    prev_end = e1
    durs.append((e1 - s0) / 1e3)
print("mean span %.1f us; steps/s implied by first->last start: %.1f us per step" % (np.mean(durs), (per_step[-1][:, 2].min() - per_step[4][:, 2].min()) / 1e3 / (NS - 5)))
