"""the oracle's classic LiDAR flavour against recordings of the reference's OWN `LIDAR` class driven through its
sensor interface (oracle/make_golden_classic_lidar.py: 181 scenes incl. ties, the radius cull, the +-z / -x wrap-arounds and
directions on the rounding borders of a cell).  Cells and float32 values must be identical."""
import os

import numpy as np

from oracle.env_oracle import lidar_project

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "classic_lidar.npz")


def test_classic_flavour_matches_the_reference_class():
    with np.load(GOLDEN) as z:
        g = {k: z[k] for k in z.files}
    S = len(g["n_ent"])
    assert S >= 150
    marked = ties = culled = 0
    for s in range(S):
        n = int(g["n_ent"][s])
        pos, types = g["ent_pos"][s, :n], g["ent_type"][s, :n]
        sph, ids = lidar_project(g["own_pos"][s], g["own_quat"][s], list(pos), list(types), list(range(1, n + 1)), "classic", 40.0)
        want = g["sphere"][s]
        assert np.array_equal(sph < 1, want < 1), f"scene {s}: marked cells differ"
        assert np.array_equal(sph, want), f"scene {s}: values differ by {np.abs(sph - want).max()}"
        m = int((want[0] < 1).sum())
        marked += m
        r = np.linalg.norm(pos - g["own_pos"][s], axis=1)
        culled += int(((r <= 0) | (r >= 40.0)).sum())
        ties += int(((r > 0) & (r < 40.0)).sum()) - m
    assert marked > 800 and ties > 20 and culled > 10, (marked, ties, culled)     # the corner cases are really in there
