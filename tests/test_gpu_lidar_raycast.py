"""Opt-in ray-cast LiDAR (dc_lidar_raycast) against oracle/lidar_raycast.py and against the
reference-pinned projection it must degenerate to."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(seed, E=24, N=16, O=6, spread=6.0):
    rng = np.random.RandomState(seed)
    pos = rng.uniform(-spread, spread, (E, N, 3)).astype(np.float32)
    q = rng.normal(size=(E, N, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
    types = np.array([3] * O + [1] * (N - O), dtype=np.int32)
    alive = (rng.rand(E, N) > 0.15).astype(np.uint8)
    return pos, q.astype(np.float32), types, alive, np.arange(O, dtype=np.int32)


def test_raycast_zero_radius_is_the_projection():
    from dronechase_b200 import lidar_project, lidar_raycast
    pos, q, types, alive, obs_slot = _scene(1)
    args = (torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda())
    tail = (torch.from_numpy(types), torch.from_numpy(alive), torch.from_numpy(obs_slot))
    ref, ref_ids = lidar_project(*args, *tail, "fused", 40.0, with_ids=True)
    got, got_ids = lidar_raycast(*args, torch.zeros(pos.shape[1]), *tail, 40.0, with_ids=True)
    assert torch.equal(got_ids, ref_ids)
    assert torch.equal(got, ref)


def test_raycast_matches_oracle():
    from dronechase_b200 import lidar_raycast
    from oracle.lidar_raycast import lidar_raycast as oracle_raycast
    E, N, O = 6, 16, 3
    pos, q, types, alive, obs_slot = _scene(2, E=E, N=N, O=O, spread=2.5)
    radius = np.full(N, 0.35, dtype=np.float32); radius[O:] = 0.2
    sph, ids = lidar_raycast(torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(radius),
                             torch.from_numpy(types), torch.from_numpy(alive), torch.from_numpy(obs_slot), 40.0, with_ids=True)
    sph, ids = sph.cpu().numpy(), ids.cpu().numpy()
    checked = hits = 0
    for e in range(E):
        for o in range(O):
            if not alive[e, o]:
                assert (sph[e, o] == 1).all() and (ids[e, o] == -1).all()
                continue
            others = [k for k in range(N) if k != o and alive[e, k]]
            s_ref, i_ref, margin = oracle_raycast(pos[e, o], q[e, o], pos[e, others], radius[others], types[others], others, 40.0)
            ok = margin > 1e-4             # grazing rays may flip between float32 and float64
            assert np.array_equal(ids[e, o][ok], i_ref[ok]), f"env {e} obs {o}: hit ids"
            assert np.array_equal(sph[e, o][1:][:, ok], s_ref[1:][:, ok])
            # stated tolerance for LiDAR distances: 1e-5 (normalised) away from grazing incidence
            assert np.abs(sph[e, o][0][ok] - s_ref[0][ok]).max() <= 1e-5
            checked += int(ok.sum()); hits += int((i_ref[ok] >= 0).sum())
    assert checked > 0.95 * E * O * 338 * alive[:, :O].mean() and hits > 200   # bodies cover many cells, unlike centres


def test_raycast_occlusion():
    """A big near body hides a small far one on the same ray; the projection would report both cells' centres."""
    from dronechase_b200 import lidar_raycast
    pos = np.array([[[0, 0, 0], [2, 0, 0], [4, 0, 0.05]]], dtype=np.float32)
    q = np.tile(np.array([0, 0, 0, 1], dtype=np.float32), (1, 3, 1))
    types = np.array([3, 1, 1], dtype=np.int32)
    sph, ids = lidar_raycast(torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda(), torch.tensor([0.1, 0.5, 0.1]),
                             torch.from_numpy(types), torch.ones(1, 3, dtype=torch.uint8), torch.tensor([0]), 40.0, with_ids=True)
    ids = ids.cpu().numpy()[0, 0]; sph = sph.cpu().numpy()[0, 0]
    # +x is theta = pi/2 (row 6), phi = 0 (col 13): both centres fall in that cell, the near body wins it
    assert ids[6, 13] == 1 and 1.5 / 40 < sph[0, 6, 13] < 2.0 / 40, "surface hit, nearer than the centre"
    assert (ids == 2).sum() == 0, "the far munition is completely occluded"
    assert (ids == 1).sum() >= 2, "a 0.5 m body at 2 m subtends more than one 0.24 rad cell"
