"""CPU oracle of the opt-in ray-cast LiDAR (dc_lidar_raycast).

TEST INFRASTRUCTURE, not the product.  The reference has NO ray cast (both of its sensor classes
project known entity centres, fused_lidar.py:143-150 docstring), so this
restates OUR definition, anchored on the reference at the two ends it does pin:
  * ray directions are the cell centres of LidarMath.radian_from_index (lidar_math.py:103-105) mapped
    through LidarMath.spherical_to_cartesian (:16-22) and rotated body -> world with the observer's
    quaternion (pybullet rotateVector convention, xyzw);
  * with every bounding radius -> 0 the sphere equals FusedLIDAR.update_data's projection sphere
    (env_oracle.lidar_project, pinned by tests/golden), because every entity always claims the cell that
    contains its centre at its centre distance.
Parity: unpinned by the reference (nothing to pin against); property-pinned as stated above.
"""
from __future__ import annotations

import numpy as np

from . import dynamics as dy
from .env_oracle import N_PHI, N_THETA, lidar_project


def lidar_raycast(own_pos, own_quat, ent_pos, ent_radius, ent_type, ent_id, max_range=40.0):
    """Returns (sphere f32 (3,13,26), ids i32 (13,26), margin f64 (13,26)).

    margin = smallest |miss distance - radius| (metres) over the entities of the cell's ray: cells whose
    margin is tiny may legitimately flip between float32 (kernel) and float64 (here) arithmetic.
    """
    own_pos32 = np.asarray(own_pos, dtype=np.float32)
    own_q32 = np.asarray(own_quat, dtype=np.float32)
    ent_pos32 = np.asarray(ent_pos, dtype=np.float32).reshape(-1, 3)
    n = len(ent_pos32)
    sphere = np.ones((3, N_THETA, N_PHI), dtype=np.float32)
    ids = np.full((N_THETA, N_PHI), -1, dtype=np.int32)
    margin = np.full((N_THETA, N_PHI), np.inf)
    # centre projections: the reference-pinned arithmetic, one entity at a time so ties do not hide anyone
    cen_cell = np.full(n, -1); cen_rn = np.ones(n)
    for k in range(n):
        s, i = lidar_project(own_pos32, own_q32.astype(np.float64), ent_pos32[k:k + 1], [ent_type[k]], [k], "fused", max_range)
        hit = np.argwhere(i >= 0)
        if len(hit):
            cen_cell[k] = hit[0][0] * N_PHI + hit[0][1]
            cen_rn[k] = float(s[0, hit[0][0], hit[0][1]])
    R = dy.rot_from_quat(own_q32.astype(np.float64))
    rel = ent_pos32.astype(np.float64) - own_pos32.astype(np.float64)
    for ti in range(N_THETA):
        for pj in range(N_PHI):
            th = (ti + 0.5) / N_THETA * np.pi
            ph = -np.pi + (pj + 0.5) / N_PHI * 2 * np.pi
            d = R @ np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
            best, best_k = 1.0, -1
            for k in range(n):
                tca = float(rel[k] @ d)
                d2 = float(rel[k] @ rel[k]) - tca * tca
                r2 = float(ent_radius[k]) ** 2
                dn = 2.0
                if tca > 0:
                    margin[ti, pj] = min(margin[ti, pj], abs(np.sqrt(max(d2, 0.0)) - float(ent_radius[k])))
                    if d2 <= r2:
                        dn = min(max((tca - np.sqrt(r2 - d2)) / max_range, 0.0), 1.0)
                if cen_cell[k] == ti * N_PHI + pj:
                    dn = min(dn, cen_rn[k])
                if dn < best:
                    best, best_k = dn, k
            if best_k >= 0:
                sphere[0, ti, pj] = best
                sphere[1, ti, pj] = ent_type[best_k] / 5
                sphere[2, ti, pj] = 0.1
                ids[ti, pj] = ent_id[best_k]
    return sphere, ids, margin
