"""N > 1 host logic on CPU (gloo, world size 2): env sharding by index range reproduces the
single-process run exactly (one Philox key space), and the rollout-end all-reduce of the episode
statistics sums the shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.env_oracle import EnvOracle, PRESETS


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _rollout(n_envs, offset, steps, seed):
    orc = EnvOracle(PRESETS["exp02_vFinal"], n_envs, seed=seed, env_offset=offset, auto_reset=True)
    orc.reset()
    rets = np.zeros(n_envs); episodes = 0
    for t in range(steps):
        rng = np.random.RandomState(1000 * t)                      # actions are a function of (t, global env)
        a_all = np.concatenate([rng.uniform(-1, 1, (64, 3)), rng.uniform(0, 1, (64, 1))], axis=1)
        obs, r, done, info = orc.step(a_all[offset:offset + n_envs])
        rets += r; episodes += int(done.sum())
    return rets, episodes, orc.pos.copy()


def _worker(rank, world, port, steps, seed, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = 4
    rets, episodes, pos = _rollout(per, rank * per, steps, seed)
    stats = torch.tensor([episodes, rets.sum(), per * steps], dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)                   # the only collective of the path
    gathered = [torch.zeros(per, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(rets))
    if rank == 0:
        out.put((stats.numpy(), torch.cat(gathered).numpy()))
    dist.destroy_process_group()


def test_two_rank_shards_match_single_process():
    steps, seed, world = 12, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, steps, seed, q)) for r in range(world)]
    for p in procs: p.start()
    stats, rets = q.get(timeout=300)
    for p in procs: p.join(timeout=60)
    ref_rets, ref_episodes, _ = _rollout(8, 0, steps, seed)
    assert np.allclose(rets, ref_rets, rtol=0, atol=1e-12), "sharded returns differ from the single-process run"
    assert stats[0] == ref_episodes and np.isclose(stats[1], ref_rets.sum()) and stats[2] == 8 * steps
