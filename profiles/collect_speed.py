"""Data-collection speed: collect_data_multiobs over the batched Level5DumbMultiObs, parts written
to local disk by the background thread.  The reference prints "Avg speed ... obs/sec" (collect_and_save.py:197-203,
io_data.py:92-104) and never records it."""
import json, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402,F401
from dronechase_b200 import BatchedThreatEngageEnv
from dronechase_b200.io_data import DatasetWriter, collect_data_multiobs

E, N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, int(sys.argv[2]) if len(sys.argv) > 2 else 40000
out = tempfile.mkdtemp(prefix="dc_collect_")
env = BatchedThreatEngageEnv("level5_dumb_multiobs", n_envs=E, seed=1, auto_reset=True)
t0 = time.time()
with DatasetWriter(out, samples_per_file=1000, backend="npz", file_stem="part") as w:
    res = collect_data_multiobs(env, w, max_observations_collected=N)
    t_loop = time.time() - t0
res["seconds_incl_flush"] = time.time() - t0
res["seconds_loop"] = t_loop
res["obs_per_sec_incl_flush"] = res["observations"] / res["seconds_incl_flush"]
res["envs"] = E
res["bytes_per_observation"] = 6 * 3 * 13 * 26 * 4 + 6 + 60 + 16 + 16
print(json.dumps(res))
shutil.rmtree(out, ignore_errors=True)
