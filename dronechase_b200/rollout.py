"""Device-resident rollout storage.

The reference collects rollouts through SB3's ``RolloutBuffer``: every env step the observation dict crosses to the
host as numpy, is copied into the buffer, and goes back to the device for the policy update
(src/core/rl_framework/utils/pipeline.py:214-241, stable-baselines3 ``OnPolicyAlgorithm.collect_rollouts``).  Here the
rollout stays in HBM.  The (3,13,26) sphere -- 4 KB per env step, 99 % of it 1.0 -- is stored as the hit list the
simulator already maintains (``dc_buffers.lidar_hits``: 8 B per drone slot) and rebuilt by ``dc_scatter_hits`` for the
time steps or the minibatch rows that are asked for: 128 steps x 65,536 envs take 1.1 GB instead of 35 GB.

    env = BatchedThreatEngageEnv("exp02_vFinal", n_envs=65536, with_hits=True)
    ro = DeviceRollout(env, n_steps=128)
    ro.collect(policy)                      # policy: obs dict of CUDA tensors -> actions [E,4] CUDA tensor
    batch = ro.minibatch(torch.randperm(ro.n_steps * env.n_envs, device="cuda")[:4096])

level5 (threatsense): the stacked observation (6,3,13,26) -- 24 KB per env step -- is stored as ITS hit list (at most
8 * (5 D + 1) bytes, -1 terminated; ``dc_scatter_stack`` rebuilds it) with the 6-byte validity mask beside it; the batch
key is ``stacked_spheres`` + ``validity_mask`` instead of ``lidar``.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import torch

from . import _lib
from .sim import BatchedThreatEngageEnv


class DeviceRollout:
    def __init__(self, env: BatchedThreatEngageEnv, n_steps: int):
        if env.lidar_hits is None:
            raise _lib.DroneChaseError("create the env with with_hits=True: the rollout stores the sphere as its hit list")
        self.env, self.n_steps = env, int(n_steps)
        T, E, D, dev = self.n_steps, env.n_envs, env.cfg.n_drones, env.device
        self.channels = env.cfg.lidar_channels
        self.level5 = env.cfg.family == "level5"
        f32 = dict(dtype=torch.float32, device=dev)
        # observation BEFORE the action of step t, as its hit list
        self.hits = torch.empty(T, E, 5 * D + 1 if self.level5 else D, 2, dtype=torch.int32, device=dev)
        self.mask = torch.empty(T, E, _lib.DC_LIDAR_STACK, dtype=torch.bool, device=dev) if self.level5 else None
        self.inertial = torch.empty(T, E, 15, **f32)
        self.last_action = torch.empty(T, E, 4, **f32)
        self.actions = torch.empty(T, E, 4, **f32)
        self.rewards = torch.empty(T, E, **f32)
        self.dones = torch.empty(T, E, dtype=torch.uint8, device=dev)
        self.pos = 0

    @property
    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.hits, self.inertial, self.last_action, self.actions,
                                                          self.rewards, self.dones, self.mask) if t is not None)

    def add(self, actions: torch.Tensor):
        """Store the env's CURRENT observation with `actions`, step the env, store reward and done (SB3 order)."""
        e, t = self.env, self.pos
        self.hits[t].copy_(e.lidar_hits, non_blocking=True)
        if self.level5:
            self.mask[t].copy_(e.obs["validity_mask"], non_blocking=True)
        self.inertial[t].copy_(e.obs["inertial_data"], non_blocking=True)
        self.last_action[t].copy_(e.obs["last_action"], non_blocking=True)
        if actions.data_ptr() != self.actions[t].data_ptr():      # a policy may have written the slab itself (FusedPolicy, out=)
            self.actions[t].copy_(actions, non_blocking=True)
        e.step(self.actions[t])                                   # zero copy: the kernel reads the stored slab
        self.rewards[t].copy_(e.reward, non_blocking=True)
        self.dones[t].copy_(e.done, non_blocking=True)
        self.pos += 1

    def collect(self, policy: Callable[[Dict[str, torch.Tensor]], torch.Tensor], n_steps: Optional[int] = None):
        """Fill the buffer from position 0: actions = policy(env.obs) on the device, nothing touches the host."""
        self.pos = 0
        direct = hasattr(policy, "precision") and hasattr(policy, "refresh")      # policy.FusedPolicy: one kernel, writes the slab
        for _ in range(n_steps or self.n_steps):
            self.add(policy(self.env.obs, out=self.actions[self.pos]) if direct else policy(self.env.obs))
        return self

    def _scatter(self, hits: torch.Tensor, index: Optional[torch.Tensor], n_rows: int) -> torch.Tensor:
        e = self.env
        ip = C.c_void_p(index.data_ptr()) if index is not None else None
        if self.level5:
            out = torch.empty(n_rows, _lib.DC_LIDAR_STACK, 3, _lib.N_THETA, _lib.N_PHI, dtype=torch.float32, device=e.device)
            with torch.cuda.device(e.device):
                _lib.check(_lib.lib().dc_scatter_stack(C.c_void_p(hits.data_ptr()), ip, n_rows, e.cfg.n_drones, C.c_void_p(out.data_ptr()),
                                                       C.c_void_p(torch.cuda.current_stream(e.device).cuda_stream)), "dc_scatter_stack")
            return out
        out = torch.empty(n_rows, self.channels, _lib.N_THETA, _lib.N_PHI, dtype=torch.float32, device=e.device)
        with torch.cuda.device(e.device):
            _lib.check(_lib.lib().dc_scatter_hits(C.c_void_p(hits.data_ptr()), C.c_void_p(index.data_ptr()) if index is not None else None,
                                                  n_rows, e.cfg.n_drones, e.cfg.n_lw, self.channels, C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream(e.device).cuda_stream)), "dc_scatter_hits")
        return out

    def lidar(self, t: int) -> torch.Tensor:
        """Dense [E, C, 13, 26] (level5: [E, 6, 3, 13, 26]) observation of time step t."""
        return self._scatter(self.hits[t], None, self.env.n_envs)

    def minibatch(self, flat_index: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Rows ``flat_index`` (int64 CUDA tensor over the flattened [T * E] axis) as a training batch."""
        idx = flat_index.to(device=self.env.device, dtype=torch.int64).contiguous()
        T, E = self.n_steps, self.env.n_envs
        lidar = self._scatter(self.hits, idx, idx.numel())
        obs = ({"stacked_spheres": lidar, "validity_mask": self.mask.view(T * E, -1)[idx]} if self.level5 else {"lidar": lidar})
        return {**obs,
                "inertial_data": self.inertial.view(T * E, 15)[idx], "last_action": self.last_action.view(T * E, 4)[idx],
                "actions": self.actions.view(T * E, 4)[idx], "rewards": self.rewards.view(T * E)[idx],
                "dones": self.dones.view(T * E)[idx]}
