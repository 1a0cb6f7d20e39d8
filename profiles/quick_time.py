import sys, torch, time
sys.path.insert(0,'.')
from dronechase_b200 import BatchedThreatEngageEnv
E=65536
env=BatchedThreatEngageEnv("exp02_vFinal", n_envs=E, seed=1234, device=0)
env.reset()
g=torch.Generator(device='cuda'); g.manual_seed(1)
bank=torch.rand(8,E,4,device='cuda',generator=g); bank[...,:3]=bank[...,:3]*2-1
bank=[bank[i].contiguous() for i in range(8)]
for i in range(150): env.step(bank[i%8])
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200): env.step(bank[i%8])
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1)/200)
t0=time.perf_counter()
for i in range(200): env.step(bank[i%8])
t1=time.perf_counter(); torch.cuda.synchronize()
print("host enqueue us/step", (t1-t0)/200*1e6)
