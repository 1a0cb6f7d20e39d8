"""dronechase_b200/policy.py on CPU: the network of the reference's LidarInertialActionExtractor PPO
(src/core/rl_framework/agents/policies/ppo_policies.py:234-342) evaluated by hand from an SB3-style state dict."""
import io
import zipfile

import torch
import torch.nn.functional as F

from dronechase_b200.policy import LidarInertialActionPolicy


def _sb3_state_dict(pi=(16, 24), features_dim=32, channels=3, prefix="features_extractor."):
    g = torch.Generator().manual_seed(3)
    r = lambda *s: torch.randn(*s, generator=g) * 0.2      # noqa: E731
    sd = {}
    for pre in (prefix, "vf_features_extractor."):
        sd.update({pre + "lidar_feature_extractor.0.weight": r(32, channels, 4, 4), pre + "lidar_feature_extractor.0.bias": r(32),
                   pre + "lidar_feature_extractor.2.weight": r(64, 32, 2, 2), pre + "lidar_feature_extractor.2.bias": r(64)})
        for name, n_in in (("inertial_feature_extractor", 15), ("action_feature_extractor", 4)):
            for j, (a, b) in enumerate(((n_in, 128), (128, 128), (128, 128))):
                sd[pre + f"{name}.{2 * j}.weight"] = r(b, a); sd[pre + f"{name}.{2 * j}.bias"] = r(b)
        sd[pre + "final_layer.0.weight"] = r(features_dim, 448); sd[pre + "final_layer.0.bias"] = r(features_dim)
    n = features_dim
    for j, h in enumerate(pi):
        sd[f"mlp_extractor.policy_net.{2 * j}.weight"] = r(h, n); sd[f"mlp_extractor.policy_net.{2 * j}.bias"] = r(h)
        sd[f"mlp_extractor.value_net.{2 * j}.weight"] = r(h, n); sd[f"mlp_extractor.value_net.{2 * j}.bias"] = r(h)
        n = h
    sd.update({"action_net.weight": r(4, n), "action_net.bias": r(4), "value_net.weight": r(1, n), "value_net.bias": r(1),
               "log_std": torch.zeros(4)})
    return sd


def _by_hand(sd, obs, pre, n_pi):
    x = F.relu(F.conv2d(obs["lidar"], sd[pre + "lidar_feature_extractor.0.weight"], sd[pre + "lidar_feature_extractor.0.bias"], stride=4))
    x = F.relu(F.conv2d(x, sd[pre + "lidar_feature_extractor.2.weight"], sd[pre + "lidar_feature_extractor.2.bias"], stride=2)).flatten(1)
    def mlp(v, name):
        for j in range(3):
            v = F.relu(F.linear(v, sd[pre + f"{name}.{2 * j}.weight"], sd[pre + f"{name}.{2 * j}.bias"]))
        return v
    f = torch.cat((x, mlp(obs["inertial_data"], "inertial_feature_extractor"), mlp(obs["last_action"], "action_feature_extractor")), 1)
    f = F.relu(F.linear(f, sd[pre + "final_layer.0.weight"], sd[pre + "final_layer.0.bias"]))
    for j in range(n_pi):
        f = torch.tanh(F.linear(f, sd[f"mlp_extractor.policy_net.{2 * j}.weight"], sd[f"mlp_extractor.policy_net.{2 * j}.bias"]))
    a = F.linear(f, sd["action_net.weight"], sd["action_net.bias"])
    return torch.clamp(a, torch.tensor([-1.0, -1, -1, 0]), torch.tensor([1.0, 1, 1, 1]))


def test_policy_matches_the_sb3_network_by_hand(tmp_path):
    sd = _sb3_state_dict()
    g = torch.Generator().manual_seed(0)
    obs = {"lidar": torch.rand(5, 3, 13, 26, generator=g), "inertial_data": torch.rand(5, 15, generator=g) * 2 - 1,
           "last_action": torch.rand(5, 4, generator=g)}
    pol = LidarInertialActionPolicy(lidar_channels=3, features_dim=32, pi=(16, 24), seed=None).load_sb3_state_dict(sd)
    want = _by_hand(sd, obs, "features_extractor.", 2)
    got = pol(obs)
    assert got.shape == (5, 4) and torch.allclose(got, want, atol=1e-6)
    assert (got[:, 3] >= 0).all() and (got.abs() <= 1).all()
    # an SB3 archive: policy.pth inside a zip; sizes are inferred from the tensors (separate pi/vf extractors: pi is taken)
    sd2 = {("pi_" + k if k.startswith("features_extractor.") else k): v for k, v in sd.items()}
    path = tmp_path / "t2_PPO_r4427.63.zip"
    buf = io.BytesIO(); torch.save(sd2, buf)
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("policy.pth", buf.getvalue()); z.writestr("data", "{}")
    pol2 = LidarInertialActionPolicy.from_sb3_zip(str(path))
    assert pol2.pi == (16, 24) and pol2.features_dim == 32 and torch.allclose(pol2(obs), want, atol=1e-6)
    assert "parameters" in pol2.describe()


def test_fragment_model_of_the_fused_kernel_matches_the_module():
    """oracle/policy_fragment_model.py restates csrc/policy_kernel.cu lane by lane (packed weight order, fragment addresses of
    every layer, in-place layers, streamed last layer + head): it must evaluate the same function as the torch module.  A wrong
    offset in the kernel's index arithmetic fails here, on the CPU."""
    import numpy as np
    from oracle.policy_fragment_model import forward
    for channels, features_dim, pi, E in ((3, 256, (128, 256, 512), 37), (2, 128, (64,), 64), (3, 192, (), 5)):
        pol = LidarInertialActionPolicy(lidar_channels=channels, features_dim=features_dim, pi=pi, seed=11).double()
        with torch.no_grad():
            for p_ in pol.parameters():                  # larger weights than the default init: saturating tanh / clipping get exercised
                p_.mul_(1.7)
        g = torch.Generator().manual_seed(E)
        obs = {"lidar": torch.rand(E, channels, 13, 26, generator=g, dtype=torch.float64),
               "inertial_data": torch.rand(E, 15, generator=g, dtype=torch.float64) * 2 - 1,
               "last_action": torch.rand(E, 4, generator=g, dtype=torch.float64)}
        obs["lidar"][obs["lidar"] > 0.3] = 1.0           # a mostly empty sphere, like the simulator's
        want = pol(obs).numpy()
        w = {k: ([t.detach().numpy() for t in v] if isinstance(v, list) else v.detach().numpy()) for k, v in pol.weight_arrays().items()}
        got = forward(w, obs["lidar"].numpy(), obs["inertial_data"].numpy(), obs["last_action"].numpy(),
                      pol.low.numpy(), pol.high.numpy(), activation=2)
        assert got.shape == (E, 4)
        assert np.abs(got - want).max() < 1e-12, f"C={channels} F={features_dim} pi={pi}: fragment model differs by {np.abs(got - want).max()}"
        assert (np.abs(want) < 1).any() and want.std() > 0.05, "test network saturates everywhere: the comparison would be vacuous"


def test_policy_c_abi_argument_checks_and_no_cpu_fallback():
    """dc_policy_create validates the shape description before touching a device, and without a GPU it refuses (DC_ERR_NO_DEVICE):
    the fused policy has no CPU path; FusedPolicy raises for a module that lives on the CPU."""
    import ctypes as C
    import pytest
    from dronechase_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    w = _lib.dc_policy_weights()
    assert L.dc_policy_create(C.byref(w), 0, C.byref(h)) == -1 and b"lidar_channels" in L.dc_last_error()
    w.lidar_channels, w.features_dim, w.n_pi, w.activation = 3, 256, 2, 2
    w.pi[0], w.pi[1] = 128, 300
    assert L.dc_policy_create(C.byref(w), 0, C.byref(h)) == -1 and b"multiples of 64" in L.dc_last_error()
    w.pi[1] = 512
    assert L.dc_policy_create(C.byref(w), 0, C.byref(h)) == -1 and b"null weight pointer" in L.dc_last_error()
    assert L.dc_policy_forward(None, None, None, None, 4, None, 0, None) == -1
    pol = LidarInertialActionPolicy(lidar_channels=3, seed=0)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DroneChaseError, match="no CPU fallback"):
            pol.fused()
        buf = (C.c_float * 4)()
        p = C.cast(buf, C.POINTER(C.c_float))
        for name, _ in _lib.dc_policy_weights._fields_:
            if name.endswith(("_w", "_b")):
                f = getattr(w, name)
                if isinstance(f, C.Array):
                    for i in range(len(f)):
                        f[i] = p
                else:
                    setattr(w, name, p)
        assert L.dc_policy_create(C.byref(w), 0, C.byref(h)) == -4 and b"no CPU fallback" in L.dc_last_error()
