"""TEST INFRASTRUCTURE, build container only: golden vectors for the classic LiDAR from the reference's OWN class.

`LIDAR` (src/core/entities/quadcopters/components/sensors/lidar.py) is driven through its own sensor interface -- the parent's
and the other publishers' inertial messages go in through `buffer_inertial_data` (:237-254), `update_data` (:263-280) rebuilds the
sphere (`_add_end_position_for_entity` -> `_add_end_position` -> `_rotate_position` / `LidarMath.cartesian_to_spherical` ->
`_add_spherical` with `_normalize_angle`'s round() modulo n), `read_data` (:283) hands it out -- unmodified from /root/reference/src,
with oracle/refshim.py standing in for the absent pybullet (`getMatrixFromQuaternion` only).  Scenes: random ones plus the corner
cases the index rule has: exact ties in one cell (the later publisher wins: '>' rejects), entities at and beyond the radius, an
entity on the observer, directions along +-z and -x (theta = pi -> row 13 % 13 = 0, phi = +-pi -> column 26 % 26 = 0), directions
on and next to the rounding borders of a cell (Python round() is half-to-even).

    python -m oracle.make_golden_classic_lidar        # writes tests/golden/classic_lidar.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np

N_MAX = 12


def scenes(rng):
    out = []

    def quat():
        q = rng.normal(size=4)
        return q / np.linalg.norm(q)
    for k in range(160):                                  # random scenes, a few entities beyond the 40 m radius
        n = rng.randint(1, N_MAX + 1)
        own = rng.uniform(-6, 6, 3)
        ents = own + rng.normal(size=(n, 3)) * rng.choice([0.5, 3.0, 12.0, 30.0])
        out.append((own, quat(), ents, rng.choice([1, 3, 4], size=n)))
    ident = np.array([0.0, 0, 0, 1])
    for own_q in (ident, quat(), quat()):                # corner cases, in the observer's frame and rotated
        own = rng.uniform(-2, 2, 3)
        from oracle import dynamics as dy
        R = dy.rot_from_quat(own_q)                       # body -> world: an entity at own + R d is seen along d

        def at(d):
            return own + R @ np.asarray(d, dtype=np.float64)
        tie = at([1.0, 2.0, 0.5])
        out.append((own, own_q, np.stack([tie, tie, tie]), np.array([1, 3, 4])))                       # same spot: the last wins
        out.append((own, own_q, np.stack([at([2, 0, 0]), at([1, 0, 0]), at([1, 0, 0]), at([3, 0, 0])]), np.array([1, 3, 1, 4])))
        out.append((own, own_q, np.stack([at([40.0, 0, 0]), at([39.999, 0, 0]), at([0, 41.0, 0]), own.copy()]), np.array([1, 1, 3, 3])))
        out.append((own, own_q, np.stack([at([0, 0, 2.0]), at([0, 0, -2.0]), at([-3.0, 0, 0]), at([-3.0, 1e-12, 0]), at([-3.0, -1e-12, 0])]),
                    np.array([1, 3, 1, 3, 4])))
        border = []
        for i in (0, 1, 6, 12):                           # theta on / next to the rounding border between rows i and i+1
            th = (i + 0.5) * np.pi / 13
            for eps in (-1e-9, 0.0, 1e-9):
                border.append(at(5.0 * np.array([np.sin(th + eps), 0.0, np.cos(th + eps)])))
        out.append((own, own_q, np.stack(border), np.array([1] * len(border))))
        border = []
        for j in (0, 1, 12, 24, 25):                      # phi borders, equator
            ph = -np.pi + (j + 0.5) * 2 * np.pi / 26
            for eps in (-1e-9, 0.0, 1e-9):
                border.append(at(7.0 * np.array([np.cos(ph + eps), np.sin(ph + eps), 0.0])))
        out.append((own, own_q, np.stack(border[:N_MAX]), np.array([3] * N_MAX)))
        out.append((own, own_q, np.stack(border[N_MAX - 9:][:N_MAX]), np.array([4] * len(border[N_MAX - 9:][:N_MAX]))))
    return out


def main():
    from oracle import refshim
    refshim.install()
    sys.path.insert(0, refshim.REFERENCE_SRC)
    from core.dataclasses.message_context import MessageContext
    from core.entities.entity_type import EntityType
    from core.entities.quadcopters.components.sensors.lidar import LIDAR

    rng = np.random.RandomState(20260)
    sc = scenes(rng)
    S = len(sc)
    own_pos = np.zeros((S, 3)); own_quat = np.zeros((S, 4)); n_ent = np.zeros(S, dtype=np.int32)
    ent_pos = np.zeros((S, N_MAX, 3)); ent_type = np.zeros((S, N_MAX), dtype=np.int32)
    spheres = np.zeros((S, 2, 13, 26), dtype=np.float32)
    for s, (own, q, ents, types) in enumerate(sc):
        lidar = LIDAR(parent_id=0, client_id=0, radius=40, resolution=16)          # 13 x 26 sectors
        assert (lidar.n_theta_points, lidar.n_phi_points) == (13, 26)
        lidar.buffer_inertial_data({"position": own.copy(), "quaternion": q.copy(), "publisher_type": EntityType.LOYALWINGMAN},
                                   MessageContext(publisher_id=0, step=0, entity_type=EntityType.LOYALWINGMAN))
        for i, (p, t) in enumerate(zip(ents, types)):
            et = EntityType(int(t))
            lidar.buffer_inertial_data({"position": p.copy(), "quaternion": np.array([0.0, 0, 0, 1]), "publisher_type": et},
                                       MessageContext(publisher_id=i + 1, step=0, entity_type=et))
        lidar.update_data()
        sph = lidar.read_data()["lidar"]
        assert sph.shape == (2, 13, 26) and sph.dtype == np.float32
        own_pos[s], own_quat[s], n_ent[s] = own, q, len(ents)
        ent_pos[s, :len(ents)], ent_type[s, :len(ents)] = ents, types
        spheres[s] = sph
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "classic_lidar.npz")
    np.savez_compressed(path, own_pos=own_pos, own_quat=own_quat, n_ent=n_ent, ent_pos=ent_pos, ent_type=ent_type, sphere=spheres)
    print(f"{S} scenes, {int((spheres[:, 0] < 1).sum())} marked cells -> {path}")


if __name__ == "__main__":
    main()
