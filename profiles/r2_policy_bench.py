"""Policy inference at the bench size: torch float32 module (cuDNN / cuBLAS), the same with TF32 allowed, and the fused kernel
(csrc/policy_kernel.cu) in both precisions; CUDA events, L2 flushed by the 266 MB sphere tensor itself."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if len(sys.argv) > 2:                                  # A/B of kernel builds: a variant of the library (profiles/micro/variants/)
    from dronechase_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(sys.argv[2])
from dronechase_b200 import BatchedThreatEngageEnv
from dronechase_b200.policy import LidarInertialActionPolicy

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = BatchedThreatEngageEnv("exp02_v2_full", n_envs=E, seed=1, device=0, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for _ in range(80):
    a = torch.rand(E, 4, device="cuda", generator=g); a[:, :3] = a[:, :3] * 2 - 1
    env.step(a)
pol = LidarInertialActionPolicy(env, seed=0)
obs = env.obs


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"envs": E, "lib": sys.argv[2] if len(sys.argv) > 2 else "in-tree"}
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
ref = pol(obs)
if len(sys.argv) <= 2:
    res["torch_fp32_ms"] = timeit(lambda: pol(obs))
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    res["torch_tf32_ms"] = timeit(lambda: pol(obs))
    res["torch_tf32_err"] = float((pol(obs) - ref).abs().max())
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
flop = 2 * E * (12 * 48 * 32 + 3 * 128 * 64 + 15 * 128 + 4 * 128 + 4 * 128 * 128 + 448 * 256 + 256 * 128 + 128 * 256 + 256 * 512 + 512 * 4)
for prec in ("3xtf32", "tf32"):
    f = pol.fused(prec)
    out = torch.empty(E, 4, device="cuda")
    res[f"fused_{prec}_err"] = float((f(obs) - ref).abs().max())
    ms = min(timeit(lambda: f(obs, out), n=40), timeit(lambda: f(obs, out), n=40))
    res[f"fused_{prec}_ms"] = ms
    res[f"fused_{prec}_tflops"] = flop / (ms * 1e-3) / 1e12
    f.close()
res["gflop_per_call"] = flop / 1e9
print(json.dumps(res))
