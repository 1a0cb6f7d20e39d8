"""DeviceRollout: the rollout kept in HBM with the sphere stored as its hit list must give back, bit for bit, the dense
observations the simulator showed (dc_scatter_hits == the incremental sphere of env_kernel)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["exp02_vFinal", "exp03_vFinal", "stage02"])
def test_rollout_rebuilds_the_observations(name):
    from dronechase_b200 import BatchedThreatEngageEnv, DeviceRollout
    E, T = 384, 48
    env = BatchedThreatEngageEnv(name, n_envs=E, seed=3, device=0, with_hits=True)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)

    def rand_policy(obs):
        a = torch.rand(E, 4, device="cuda", generator=g); a[:, :3] = a[:, :3] * 2 - 1
        return a
    for _ in range(70):                                # get past the empty first spheres
        env.step(rand_policy(env.obs))
    ro = DeviceRollout(env, T)
    dense, inertial, rewards, dones, acts = [], [], [], [], []

    def recording_policy(obs):
        dense.append(obs["lidar"].clone()); inertial.append(obs["inertial_data"].clone())
        a = rand_policy(obs); acts.append(a.clone())
        return a
    ro.pos = 0
    for t in range(T):
        ro.add(recording_policy(env.obs))
        rewards.append(env.reward.clone()); dones.append(env.done.clone())
    marked = 0
    for t in range(T):
        got = ro.lidar(t)
        assert torch.equal(got, dense[t]), f"step {t}: rebuilt sphere differs"
        marked += int((got < 1).sum())
        assert torch.equal(ro.inertial[t], inertial[t]) and torch.equal(ro.actions[t], acts[t])
        assert torch.equal(ro.rewards[t], rewards[t]) and torch.equal(ro.dones[t], dones[t])
    assert marked > 1000
    idx = torch.randperm(T * E, device="cuda", generator=g)[:1000]
    mb = ro.minibatch(idx)
    all_dense = torch.stack(dense).view(T * E, *dense[0].shape[1:])
    assert torch.equal(mb["lidar"], all_dense[idx])
    assert torch.equal(mb["rewards"], torch.stack(rewards).view(-1)[idx])
    assert ro.bytes < 0.08 * (T * E * dense[0][0].numel() * 4)        # sparse storage: a few % of the dense rollout
    env.close()


@pytest.mark.parametrize("name", ["level5_c1", "level5_fusion"])
def test_rollout_rebuilds_the_stacked_observations(name):
    """level5: the (6,3,13,26) stack is kept as its hit list + validity mask; dc_scatter_stack gives back the bits."""
    from dronechase_b200 import BatchedThreatEngageEnv, DeviceRollout
    E, T = 160, 40
    env = BatchedThreatEngageEnv(name, n_envs=E, seed=5, device=0, with_hits=True)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)

    def rand_policy(obs):
        a = torch.rand(E, 4, device="cuda", generator=g); a[:, :3] = a[:, :3] * 2 - 1
        return a
    for _ in range(30):
        env.step(rand_policy(env.obs))
    ro = DeviceRollout(env, T)
    dense, masks = [], []
    for t in range(T):
        dense.append(env.obs["stacked_spheres"].clone()); masks.append(env.obs["validity_mask"].clone())
        ro.add(rand_policy(env.obs))
    marked = 0
    for t in range(T):
        got = ro.lidar(t)
        assert torch.equal(got, dense[t]), f"step {t}: rebuilt stack differs"
        assert torch.equal(ro.mask[t], masks[t])
        marked += int((got < 1).sum())
    assert marked > 1000
    idx = torch.randperm(T * E, device="cuda", generator=g)[:700]
    mb = ro.minibatch(idx)
    assert torch.equal(mb["stacked_spheres"], torch.stack(dense).view(T * E, 6, 3, 13, 26)[idx])
    assert torch.equal(mb["validity_mask"], torch.stack(masks).view(T * E, 6)[idx])
    assert ro.bytes < 0.2 * (T * E * 6 * 3 * 338 * 4)
    env.close()


def test_policy_driven_rollout_against_the_oracle():
    """An INDEPENDENT source for the rollout: the float64 oracle is stepped with the very actions the device-resident
    policy (dronechase_b200.policy, the reference's LidarInertialActionExtractor PPO network) produced on the GPU; what the
    rollout stored -- sphere rebuilt from its hit list, inertial vector, reward, done -- must equal what the oracle shows at
    every time step.  f64 build of the simulator, so events and LiDAR cells are exact."""
    from dronechase_b200 import BatchedThreatEngageEnv, DeviceRollout
    from dronechase_b200.policy import LidarInertialActionPolicy
    from oracle.env_oracle import EnvOracle
    from tests.util import oracle_cfg
    E, T, seed = 32, 90, 17
    env = BatchedThreatEngageEnv("exp02_vFinal", n_envs=E, seed=seed, device=0, auto_reset=True, precision="f64", with_hits=True)
    orc = EnvOracle(oracle_cfg("exp02_vFinal"), E, seed=seed, auto_reset=True)
    env.reset()
    ref = orc.reset()
    pol = LidarInertialActionPolicy(env, seed=5)
    # an untrained network barely moves the drone: add a deterministic push towards the nearest munition seen in the sphere
    def policy(obs):
        a = pol(obs)
        a[:, 3] = 1.0
        return a.contiguous()
    ro = DeviceRollout(env, T).collect(policy)
    marked = kills = 0
    for t in range(T):
        assert np.abs(ro.lidar(t).cpu().numpy() - ref["lidar"]).max() < 1e-6, f"step {t}: sphere stored in the rollout"
        assert np.array_equal(ro.lidar(t).cpu().numpy() < 1, ref["lidar"] < 1), f"step {t}: marked cells"
        assert np.abs(ro.inertial[t].cpu().numpy() - ref["inertial_data"]).max() < 1e-6, f"step {t}: inertial"
        marked += int((ref["lidar"] < 1).sum())
        ref, r_ref, d_ref, i_ref = orc.step(ro.actions[t].cpu().numpy().astype(np.float64))
        assert np.array_equal(ro.dones[t].cpu().numpy().astype(bool), d_ref), f"step {t}: done"
        assert np.allclose(ro.rewards[t].cpu().numpy(), r_ref, rtol=1e-6, atol=1e-5), f"step {t}: reward"
        kills = max(kills, int(i_ref["agent_kills"].max()))
    assert marked > 500
    # the policy's output is the clipped mean action of the network: inside the action box, last component pinned above
    assert float(ro.actions[..., :3].abs().max()) <= 1.0 and float(ro.actions[..., 3].min()) == 1.0
    env.close()
