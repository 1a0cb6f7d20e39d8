"""Time one build variant of the library: python profiles/r2_variants.py <lib.so | pkg_root> [preset] [envs] [K ...]
(<pkg_root> = a directory holding another copy of the dronechase_b200 package, e.g. the previous commit's)."""
import os, sys
arg = sys.argv[1]
if os.path.isdir(arg):
    sys.path.insert(0, arg)
else:
    sys.path.insert(0, '.')
import torch
import dronechase_b200._lib as _lib
if not os.path.isdir(arg):
    _lib.LIB_PATH = os.path.abspath(arg)
from dronechase_b200 import BatchedThreatEngageEnv
name = sys.argv[2] if len(sys.argv) > 2 else "exp02_vFinal"
E = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
Ks = [int(k) for k in sys.argv[4:]] or [2]
g = torch.Generator(device='cuda'); g.manual_seed(1)
bank = torch.rand(8, E, 4, device='cuda', generator=g); bank[..., :3] = bank[..., :3] * 2 - 1
bank = [bank[i].contiguous() for i in range(8)]
for K in Ks:
    env = BatchedThreatEngageEnv(name, n_envs=E, seed=1234, device=0, sub_batches=K)
    env.reset()
    for i in range(150): env.step(bank[i % 8])
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): env.step(bank[i % 8])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200)
    print(f"{os.path.basename(arg):16s} {name} E={E} K={K}: {best:.4f} ms/step", flush=True)
    env.close()
