"""Experiment: does splitting the env batch into K independent sub-sims on K streams (captured in one CUDA graph)
let the latency-bound env_kernel of one chunk run under the issue-bound dyn_kernel of another?
python profiles/chunk_overlap.py [preset] [E_total]"""
import sys
sys.path.insert(0, '.')
import torch
from dronechase_b200 import BatchedThreatEngageEnv

name = sys.argv[1] if len(sys.argv) > 1 else "exp02_vFinal"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
NB = 8


def run(K, graph=True, steps=200):
    per = E // K
    sims = [BatchedThreatEngageEnv(name, n_envs=per, seed=1234, device=0, env_offset=k * per) for k in range(K)]
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    banks = []
    for k in range(K):
        b = torch.rand(NB, per, 4, device='cuda', generator=g); b[..., :3] = b[..., :3] * 2 - 1
        banks.append([b[i].contiguous() for i in range(NB)])
    for s in sims: s.reset()
    for i in range(152):
        for k, s in enumerate(sims): s.step(banks[k][i % NB])
    torch.cuda.synchronize()
    side = [torch.cuda.Stream() for _ in range(K)]

    def body():
        cur = torch.cuda.current_stream()
        for k, s in enumerate(sims):
            side[k].wait_stream(cur)
            with torch.cuda.stream(side[k]):
                for i in range(NB): s.step(banks[k][i])
        for k in range(K): cur.wait_stream(side[k])

    if graph:
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            body()
        fn = cg.replay
    else:
        fn = body
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = steps // NB
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (n * NB)
    print(f"{name} E={E} K={K} graph={graph}: ms/step {ms:.4f}  env-steps/s {E / ms * 1e3:.3e}", flush=True)
    for s in sims: s.close()


run(1, graph=False)
KS = [int(k) for k in sys.argv[3].split(",")] if len(sys.argv) > 3 else (1, 2, 3, 4, 6, 8, 16)
for K in KS:
    if E % K == 0:
        run(K)
