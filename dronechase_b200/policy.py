"""Device-resident policy inference (second half).

The reference drives the agent with an SB3 PPO whose feature extractor is ``LidarInertialActionExtractor``
(src/core/rl_framework/agents/policies/ppo_policies.py:234-342): Conv2d(C,32,k4,s4)-ReLU-Conv2d(32,64,k2,s2)-ReLU-Flatten
over the (C,13,26) sphere, two 3 x 128 MLPs over ``inertial_data`` and ``last_action``, Linear(448, features_dim)-ReLU; SB3 then
applies ``mlp_extractor.policy_net`` (``net_arch["pi"]``, Tanh by default -- ``create_policy_kwargs`` :150-156 passes
``net_arch=dict(pi=hiddens, vf=hiddens)``) and ``action_net``; ``predict(deterministic=True)`` returns the Gaussian mean
clipped to the action box.  Through SubprocVecEnv every step of that costs an observation round trip host -> device ->
host.  ``LidarInertialActionPolicy`` is the same network as a plain ``torch.nn.Module`` that reads the simulator's
observation tensors where they are (HBM) and writes the action tensor ``dc_step`` consumes: with ``DeviceRollout`` nothing
crosses PCIe during collection.  ``from_sb3_zip`` loads the weights of a model the reference trained (the ``policy.pth``
inside an SB3 ``.zip``) without needing stable-baselines3.
"""
from __future__ import annotations

import io
import zipfile
from typing import Dict, Optional, Sequence

import torch
from torch import nn


def _mlp3(n_in: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(n_in, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU())


class LidarInertialActionPolicy(nn.Module):
    def __init__(self, env=None, lidar_channels: int = 3, features_dim: int = 256, pi: Sequence[int] = (128, 256, 512),
                 activation: str = "tanh", seed: Optional[int] = 0, device=None):
        super().__init__()
        if env is not None:
            lidar_channels, device = env.cfg.lidar_channels, env.device
        if seed is not None:
            torch.manual_seed(int(seed))
        self.lidar_feature_extractor = nn.Sequential(nn.Conv2d(lidar_channels, 32, kernel_size=4, stride=4), nn.ReLU(),
                                                     nn.Conv2d(32, 64, kernel_size=2, stride=2), nn.ReLU(), nn.Flatten())
        self.inertial_feature_extractor = _mlp3(15)
        self.action_feature_extractor = _mlp3(4)
        n_lidar = 64 * ((13 // 4) // 2) * ((26 // 4) // 2)                 # (32,3,6) -> (64,1,3) = 192
        self.final_layer = nn.Sequential(nn.Linear(n_lidar + 128 + 128, features_dim), nn.ReLU())
        act = {"tanh": nn.Tanh, "relu": nn.ReLU}[activation]
        layers, n = [], features_dim
        for h in pi:
            layers += [nn.Linear(n, int(h)), act()]
            n = int(h)
        self.policy_net = nn.Sequential(*layers)
        self.action_net = nn.Linear(n, 4)
        self.register_buffer("low", torch.tensor([-1.0, -1.0, -1.0, 0.0]))
        self.register_buffer("high", torch.tensor([1.0, 1.0, 1.0, 1.0]))
        self.pi, self.features_dim = tuple(int(h) for h in pi), int(features_dim)
        if device is not None:
            self.to(device)
        self.eval()

    def features(self, obs: Dict[str, torch.Tensor]) -> torch.Tensor:
        f = torch.cat((self.lidar_feature_extractor(obs["lidar"]), self.inertial_feature_extractor(obs["inertial_data"].flatten(1)),
                       self.action_feature_extractor(obs["last_action"])), dim=1)
        return self.final_layer(f)

    @torch.no_grad()
    def forward(self, obs: Dict[str, torch.Tensor]) -> torch.Tensor:
        """``model.predict(obs, deterministic=True)``: the mean action, clipped to Box([-1,-1,-1,0],[1,1,1,1])."""
        a = self.action_net(self.policy_net(self.features(obs)))
        return torch.maximum(torch.minimum(a, self.high), self.low).contiguous()

    def describe(self) -> str:
        n = sum(p.numel() for p in self.parameters())
        return (f"LidarInertialActionExtractor (ppo_policies.py:234-342) + pi {list(self.pi)} + action_net, {n} parameters, float32, "
                "deterministic mean action")

    # ------------------------------------------------------------------ weights of a reference-trained SB3 model
    def load_sb3_state_dict(self, sd: Dict[str, torch.Tensor]) -> "LidarInertialActionPolicy":
        """``sd`` = ``model.policy.state_dict()`` of an SB3 PPO (keys ``features_extractor.*`` or ``pi_features_extractor.*``,
        ``mlp_extractor.policy_net.*``, ``action_net.*``); value-function and log_std entries are ignored."""
        fx = "pi_features_extractor." if any(k.startswith("pi_features_extractor.") for k in sd) else "features_extractor."
        mine = {}
        for k, v in sd.items():
            if k.startswith(fx):
                mine[k[len(fx):]] = v
            elif k.startswith("mlp_extractor.policy_net."):
                mine["policy_net." + k[len("mlp_extractor.policy_net."):]] = v
            elif k.startswith("action_net."):
                mine[k] = v
        want = {k for k in self.state_dict() if k not in ("low", "high")}
        missing = want - set(mine)
        if missing:
            raise KeyError(f"SB3 state dict lacks {sorted(missing)[:4]} ...: not a LidarInertialActionExtractor PPO with pi={list(self.pi)}")
        self.load_state_dict({**{k: mine[k] for k in want}, "low": self.low, "high": self.high})
        return self

    @classmethod
    def from_sb3_zip(cls, path: str, env=None, **kw) -> "LidarInertialActionPolicy":
        """Build from an SB3 ``model.save`` archive: reads ``policy.pth``, infers pi sizes and features_dim from the shapes."""
        with zipfile.ZipFile(path) as z:
            sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        pi, i = [], 0
        while f"mlp_extractor.policy_net.{i}.weight" in sd:
            pi.append(sd[f"mlp_extractor.policy_net.{i}.weight"].shape[0])
            i += 2
        fx = "pi_features_extractor." if any(k.startswith("pi_features_extractor.") for k in sd) else "features_extractor."
        kw.setdefault("features_dim", sd[fx + "final_layer.0.weight"].shape[0])
        kw.setdefault("lidar_channels", sd[fx + "lidar_feature_extractor.0.weight"].shape[1])
        return cls(env=env, pi=pi, seed=None, **kw).load_sb3_state_dict(sd)
