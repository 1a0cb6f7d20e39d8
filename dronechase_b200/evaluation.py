"""Batched evaluation drivers (the reference's evaluation apps over one GPU batch instead of one env).

``evaluate_2bt`` is apps/threatsense_runner/evaluation_2bt.py: N episodes of ``Level52BTEvaluationEnvironment`` (two
behaviour-tree wingmen vs 5 -> 30 munitions), per episode the kills of each wingman at termination
(``info["kills_per_drone"]``, level5_2bt_evaluation_task.py:470-477) and their sum, then mean / std per column
(the ``raw_results`` and ``summary_stats`` sheets of results_2bt.xlsx).  Here the episodes run side by side, and the
sample is a FIXED QUOTA PER ENV: with ``n_envs`` envs every env contributes exactly its first
``ceil(n_episodes / n_envs)`` episodes (by default ``n_envs == n_episodes``, one episode each), whatever their length, and
the batch is stepped until the slowest env has delivered its quota.  Keeping "the first n episodes that finish" instead
would be length-biased -- envs whose wingmen die early recycle and contribute second and third short, low-kill episodes
before the 1301-step time-out episodes end -- and the reference runs N independent episodes to completion.  The rows are
ordered by (episode index of the env, env index): a deterministic function of ``seed`` and ``n_envs`` that does not
depend on the outcome.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

WINGMAN_NAMES = ("loyalwingman_0", "loyalwingman_1")


def episodes_from_steps(done: np.ndarray, info: np.ndarray, step_index: int, rows: List[Dict], quota: int,
                        names=WINGMAN_NAMES, counts: Optional[np.ndarray] = None) -> np.ndarray:
    """Append one row per env that finished at this step and has not yet delivered ``quota`` episodes.
    ``info`` = the [E, 8] counters of the step: agent_kills (slot 0), allies_kills (slot 1), deads, current_wave, ...
    ``counts`` ([E] int, episodes taken from every env so far) is updated and returned; pass the returned array back in."""
    if counts is None:
        counts = np.zeros(done.shape[0], dtype=np.int64)
    for e in np.nonzero(done)[0]:
        if counts[e] >= quota:
            continue
        k0, k1 = int(info[e, 0]), int(info[e, 1])
        rows.append({names[0]: k0, names[1]: k1, "total_kills": k0 + k1, "deads": int(info[e, 2]),
                     "current_wave": int(info[e, 3]), "episode_steps": int(info[e, 7]), "env": int(e), "step": step_index,
                     "episode": int(counts[e])})
        counts[e] += 1
    return counts


def select_rows(rows: List[Dict], n_episodes: int) -> List[Dict]:
    """The evaluation sample: rows in (episode-of-the-env, env) order, cut to ``n_episodes`` (the cut depends on the env index only)."""
    return sorted(rows, key=lambda r: (r["episode"], r["env"]))[:n_episodes]


def summarise(rows: List[Dict], names=WINGMAN_NAMES) -> Tuple[Dict[str, List], Dict[str, Dict[str, float]]]:
    """(raw_results columns, summary_stats) as evaluation_2bt.py builds them with pandas (mean and sample std, ddof=1)."""
    cols = list(names) + ["total_kills"]
    raw = {c: [r[c] for r in rows] for c in cols}
    stats = {"mean": {}, "std": {}}
    for c in cols:
        v = np.asarray(raw[c], dtype=np.float64)
        stats["mean"][c] = float(v.mean()) if len(v) else float("nan")
        stats["std"][c] = float(v.std(ddof=1)) if len(v) > 1 else float("nan")
    return raw, stats


def evaluate_2bt(n_episodes: int = 100, n_envs: Optional[int] = None, seed: int = 0, device=0, max_steps: int = 100_000,
                 output_file: Optional[str] = None, **preset_overrides):
    """Run the 2BT evaluation on the GPU batch.  Returns (rows, raw_results, summary_stats); with ``output_file`` the two
    tables are also written -- ``.xlsx`` through pandas when it has an Excel engine, else two ``.csv`` files."""
    import torch
    from . import preset
    from .sim import BatchedThreatEngageEnv
    n_envs = int(n_envs or min(n_episodes, 65536))
    quota = -(-int(n_episodes) // n_envs)          # ceil: the same number of episodes from every env
    env = BatchedThreatEngageEnv(preset("level5_eval_2bt", **preset_overrides), n_envs=n_envs, seed=seed, device=device,
                                 auto_reset=True)
    env.reset()
    rows: List[Dict] = []
    counts = np.zeros(n_envs, dtype=np.int64)
    t = 0
    while int(counts.min()) < quota and t < max_steps:
        _, _, done, info = env.step(None)
        t += 1
        if bool(done.any()):                       # one small D2H per step; the counters only when an episode ended
            counts = episodes_from_steps(done.cpu().numpy().astype(bool), info.cpu().numpy(), t, rows, quota, counts=counts)
    env.close()
    rows = select_rows(rows, n_episodes)
    raw, stats = summarise(rows)
    if output_file:
        write_results(output_file, raw, stats)
    return rows, raw, stats


def write_results(output_file: str, raw, stats) -> List[str]:
    import os
    os.makedirs(os.path.dirname(os.path.abspath(output_file)), exist_ok=True)
    try:
        import pandas as pd
        df, df_stats = pd.DataFrame(raw), pd.DataFrame(stats)
        if output_file.endswith(".xlsx"):
            try:
                with pd.ExcelWriter(output_file) as writer:
                    df.to_excel(writer, sheet_name="raw_results", index=False)
                    df_stats.to_excel(writer, sheet_name="summary_stats")
                return [output_file]
            except Exception:                      # noqa: BLE001 -- no Excel engine in this image: fall through to csv
                pass
        stem = output_file.rsplit(".", 1)[0]
        df.to_csv(stem + "_raw_results.csv", index=False)
        df_stats.to_csv(stem + "_summary_stats.csv")
        return [stem + "_raw_results.csv", stem + "_summary_stats.csv"]
    except ImportError:
        stem = output_file.rsplit(".", 1)[0]
        cols = list(raw)
        with open(stem + "_raw_results.csv", "w") as f:
            f.write(",".join(cols) + "\n")
            for i in range(len(raw[cols[0]]) if cols else 0):
                f.write(",".join(str(raw[c][i]) for c in cols) + "\n")
        with open(stem + "_summary_stats.csv", "w") as f:
            f.write(",mean,std\n")
            for c in cols:
                f.write(f"{c},{stats['mean'][c]},{stats['std'][c]}\n")
        return [stem + "_raw_results.csv", stem + "_summary_stats.csv"]


# ------------------------------------------------------------------------------------------------ level4 evaluation apps
def evaluate_level4(configuration: Dict, n_episodes: int = 100, n_envs: Optional[int] = None, seed: int = 0, device=0,
                    policies: Optional[Dict[int, object]] = None, max_steps: int = 200_000):
    """apps/threatengage_runner/stage03/experiments/*/evaluation_exp0*_app_ready.py over one GPU batch: N episodes of
    ``EvaluationEnvironment(configuration)``; per episode and per wingman the LAST info row the episode showed for it
    (``update_data`` :52-64 keeps the row with the largest ``step``), flattened like ``preprocess_data`` (:27-49) into the
    columns ["kills", "alive", "munitions", "wave", "step", "name", "episode"].  Same fixed quota of episodes per env as
    ``evaluate_2bt``.  ``policies``: wingman slot -> policy (a callable on the observation dict of device tensors, or an
    object with SB3's ``predict``) for the "nn" drivers whose configuration entry has no loadable ``path``.
    Returns (rows, columns): rows = list of [kills, alive, munitions, wave, step, name, episode]."""
    import torch
    from .config import evaluation_preset
    from .drivers import TaskDrivers, sb3_policy
    from .policy import LidarInertialActionPolicy
    from .sim import BatchedThreatEngageEnv
    cfg = evaluation_preset(configuration)
    n_envs = int(n_envs or min(n_episodes, 65536))
    quota = -(-int(n_episodes) // n_envs)
    env = BatchedThreatEngageEnv(cfg, n_envs=n_envs, seed=seed, device=device, auto_reset=True)
    names = [str(d.get("name", f"lw_{j}")) for j, d in enumerate(configuration["drivers"])]
    pol = {}
    for j in cfg.policy_slots:
        p = (policies or {}).get(j) or configuration["drivers"][j].get("policy")
        if p is None:
            p = LidarInertialActionPolicy.from_sb3_zip(configuration["drivers"][j]["path"], env=env)
        pol[j] = sb3_policy(p) if hasattr(p, "predict") else p
    drivers = TaskDrivers(env, pol) if pol else None
    env.reset()
    L = cfg.n_lw
    last = torch.zeros(n_envs, L, 5, dtype=torch.int64, device=env.device)        # kills, alive, munitions, wave, step
    seen = torch.zeros(n_envs, L, dtype=torch.bool, device=env.device)
    counts = np.zeros(n_envs, dtype=np.int64)
    episodes: List[List] = []          # (env, episode-of-env, rows)
    t = 0
    while int(counts.min()) < quota and t < max_steps:
        if drivers is not None:
            drivers.step(None)
        else:
            env.step(None)
        t += 1
        li = env.lw_info.long()
        armed = li[..., 1] > 0
        row = torch.stack([li[..., 0], li[..., 1], li[..., 2], env.info[:, 3:4].long().expand(-1, L), env.info[:, 5:6].long().expand(-1, L)], dim=-1)
        last = torch.where(armed[..., None], row, last)
        seen |= armed
        done = env.done.bool()
        if bool(done.any()):
            idx = torch.nonzero(done).view(-1)
            rows_np, seen_np = last[idx].cpu().numpy(), seen[idx].cpu().numpy()
            for k, e in enumerate(idx.cpu().numpy()):
                if counts[e] < quota:
                    episodes.append((int(counts[e]), int(e), [[int(v) for v in rows_np[k, j, :5]] + [names[j]] for j in range(L) if seen_np[k, j]]))
                    counts[e] += 1
            last[idx] = 0
            seen[idx] = False
    env.close()
    episodes.sort(key=lambda x: (x[0], x[1]))
    rows = []
    for i, (_, _, ep_rows) in enumerate(episodes[:n_episodes]):
        for r in ep_rows:
            rows.append([r[0], bool(r[1]), r[2], r[3], r[4], r[5], i + 1])
    return rows, ["kills", "alive", "munitions", "wave", "step", "name", "episode"]
