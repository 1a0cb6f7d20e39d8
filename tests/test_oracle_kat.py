"""Known-answer vectors the reference's own code produced (recorded by
importing /root/reference/src through pybullet/PyFlyt stubs) and the asserts its unit tests
carry -- replayed against the oracle's function-level restatements."""
import math

import numpy as np

from oracle import dynamics as dy
from oracle.env_oracle import EnvOracle, PRESETS, Stage03Config, calculate_rounds, lidar_project, N_THETA, N_PHI

I4 = np.array([0.0, 0.0, 0.0, 1.0])


def test_lidarspec_grid():
    # core/dataclasses/angle_grid.py:25-44 with resolution 16
    side = math.sqrt(1 / 16)
    assert side == 0.25
    assert (math.ceil(math.pi / side), math.ceil(2 * math.pi / side)) == (N_THETA, N_PHI) == (13, 26)


def test_fused_projection_known_cells():
    # KAT (2) of SURVEY 8c: observer at origin, identity quaternion, radius 40
    cases = [((1, 0, 0), 0.025, (6, 13)), ((0, 2, 0), 0.05, (6, 19)), ((0, 0, 3), 0.075, (0, 13)),
             ((-1, -1e-3, -1), 0.035355, (9, 0))]
    for p, rn, cell in cases:
        sph, ids = lidar_project(np.zeros(3), I4, [np.array(p, float)], [1], [5], "fused", 40.0)
        assert ids[cell] == 5 and (ids >= 0).sum() == 1
        assert abs(float(sph[0][cell]) - rn) < 1e-6
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([1.0, 0, 0]), np.array([2.0, 0, 0])], [1, 1], [3, 4], "fused", 40.0)
    assert ids[6, 13] == 3                                   # (2,0,0) loses to (1,0,0)
    assert np.allclose(sph[:, 6, 13], [0.025, 0.2, 0.1])     # LM flag 1/5, normalised age 0.1
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([1.0, 0, 0])], [3], [9], "fused", 40.0)
    assert np.isclose(sph[1, 6, 13], 0.6)                    # LW flag 3/5


def test_fused_tie_first_wins_and_far_clips():
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([10.0, 0, 0]), np.array([10.0, 0, 0])], [1, 3], [1, 2], "fused", 40.0)
    assert ids[6, 13] == 1                                   # strict '<': first entity keeps the cell
    # ...but the cell holds float32 and the challenger is float64 (lidar_math.py:296-304): 0.025 rounds UP
    # in float32, so an exactly equidistant second entity does take the cell.  Reference behaviour, kept.
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([1.0, 0, 0]), np.array([1.0, 0, 0])], [1, 3], [1, 2], "fused", 40.0)
    assert ids[6, 13] == 2
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([50.0, 0, 0])], [1], [1], "fused", 40.0)
    assert (ids == -1).all() and (sph == 1).all()            # r > R clips to 1.0 and never wins


def test_classic_index_rule():
    # KAT (3): theta=pi/2 -> row 6 (round(6.5) = 6, banker's), phi=0 -> col 13, phi=pi -> col 0, theta=pi -> row 0
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([1.0, 0, 0])], [1], [1], "classic", 40.0)
    assert ids[6, 13] == 1
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([-1.0, 0, 0])], [1], [1], "classic", 40.0)
    assert ids[6, 0] == 1
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([0, 0, -1.0])], [1], [1], "classic", 40.0)
    assert ids[0, 13] == 1                                   # theta = pi wraps to row 0
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([1.0, 0, 0]), np.array([1.0, 0, 0])], [1, 3], [1, 2], "classic", 40.0)
    assert ids[6, 13] == 2 and sph.shape == (2, 13, 26)      # non-strict: last wins ties
    sph, ids = lidar_project(np.zeros(3), I4, [np.array([40.0, 0, 0])], [1], [1], "classic", 40.0)
    assert (ids == -1).all()                                 # culled unless 0 < r < radius


def test_lidar_math_reframe_expectation():
    # sensors/components/math_test.py:11-69: neighbour at origin sees a target at r=0.5*20 m, theta=pi/2,
    # phi=0; own drone at (1,0,0), R=20 -> r ~= 0.45, theta ~= pi/2, phi ~= 0
    target_world = np.array([10.0, 0.0, 0.0])
    sph, ids = lidar_project(np.array([1.0, 0, 0]), I4, [target_world], [1], [1], "fused", 20.0)
    assert abs(float(sph[0, 6, 13]) - 0.45) < 0.01


def test_spherical_cartesian_kat():
    # apps/threatengage_runner/stage02/auxiliary/test_lidar.py:12-20
    r, th, ph = 1.0, np.pi / 4, np.pi / 4
    c = np.array([r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)])
    assert np.allclose(c, [0.5, 0.5, math.sqrt(2) / 2])


def test_calculate_rounds():
    assert calculate_rounds(1, 20) == 6 and calculate_rounds(2, 20) == 9


def test_generate_positions_formula():
    # KAT (5): np.random.seed(0); generate_positions(2, 6) in the reference.  The first four uniforms of
    # that Mersenne stream are fixed numbers; feed them through the oracle's formula.
    u = np.array([0.5488135039273248, 0.7151893663724195, 0.6027633760716439, 0.5448831829968969])
    thetas = np.pi * u[:2]
    min_phi = np.arccos(4 / 6)
    phis = min_phi + (np.pi / 2 - min_phi) * u[2:]
    got = np.column_stack((6 * np.sin(phis) * np.cos(thetas), 6 * np.sin(phis) * np.sin(thetas), 6 * np.cos(phis)))
    want = np.array([[-0.87827368, 5.68220358, 1.71499207], [-3.54909453, 4.42459714, 1.95623826]])
    assert np.allclose(got, want, atol=1e-8)


def test_cone_known_answers():
    # KAT (6) + navigators/legacy/geometry_utils_test.py:6-26
    apex, base = np.array([0, 0, 10.0]), np.array([0, 0, 1.0])
    assert EnvOracle._inside_cone(np.array([0, 1.0, 5.0]), apex, base, 60)
    assert not EnvOracle._inside_cone(np.array([0, 5.0, 5.0]), apex, base, 60)
    ang = np.degrees(np.arccos(np.dot([0, 1, -5], [0, 0, -9]) / (np.linalg.norm([0, 1, -5]) * 9)))
    assert abs(ang - 11.3099) < 1e-3


def _one_env(cfg):
    orc = EnvOracle(cfg, 1, seed=1)
    orc.reset()
    return orc


def test_lm_fsm_known_answer():
    # KAT (7) / navigators/tests/test_lm_navigator.py:141-184: LWs (0,5,5),(0,1,5), LM (0,0,10),
    # building (0,0,1) -> CollideWithWingman, and the command of that step is still Wait's [0,0,0,0.4]
    cfg = Stage03Config(n_lw=2, n_lm=1, lm_nav="full", building=(0.0, 0.0, 1.0), noise_ratio=0.0)
    orc = _one_env(cfg)
    for slot, p in ((0, (0, 5, 5)), (1, (0, 1, 5)), (2, (0, 0, 10))):
        orc.imu["position"][0, slot] = p
        orc.armed[0, slot] = True
    orc._offsets(0)
    orc.nav[0] = 0
    orc._navigate(0)
    assert orc.nav[0, 2] == 1
    assert np.allclose(orc.setpoint[0, 2], 0.0)
    # second scenario: LWs (0,5,5),(0,10,5) -> CollideWithBuilding
    orc.imu["position"][0, 1] = (0, 10, 5)
    orc._offsets(0)
    orc.nav[0] = 0
    orc._navigate(0)
    assert orc.nav[0, 2] == 2


def test_gun_known_answers():
    # KAT (8): fresh -> [1,0,1], cooldown 60; one shot of 20 -> [0.95,1,0]; step 30 -> [0.95,.5,0]; step 60 -> [0.95,0,1]
    orc = _one_env(PRESETS["exp02_vFinal"])
    assert list(orc._gun_state(0, 0)) == [1.0, 0.0, 1.0]
    orc.ammo[0, 0] -= 1
    orc.last_fired[0, 0] = 0
    assert list(orc._gun_state(0, 0)) == [0.95, 1.0, 0.0]
    orc.step_count[0] = 30
    assert list(orc._gun_state(0, 0)) == [0.95, 0.5, 0.0]
    orc.step_count[0] = 60
    assert list(orc._gun_state(0, 0)) == [0.95, 0.0, 1.0]


def test_lw_behaviour_tree_leaves():
    # navigators/tests/test_lw_navigator.py:140-212: gun available -> chase; unavailable (has ammo) ->
    # formation; 0 ammo -> sacrifice (= chase nearest)
    cfg = Stage03Config(n_lw=2, n_lm=1, noise_ratio=0.0)
    orc = _one_env(cfg)
    orc.imu["position"][0, 0] = (0, 0, 1); orc.imu["position"][0, 1] = (1, 0, 1); orc.imu["position"][0, 2] = (5, 0, 1)
    orc.formation[0, 1] = (1, -3, 1)
    orc.armed[0] = True
    orc._offsets(0)
    orc._navigate(0)
    assert np.allclose(orc.setpoint[0, 1], [0.6, 0, 0, 0])           # ChaseThreat
    orc.last_fired[0, 1] = 0; orc.step_count[0] = 10                   # reloading, ammo left
    orc._navigate(0)
    assert np.allclose(orc.setpoint[0, 1], [0, -0.6, 0, 0])          # MoveToFormation
    orc.ammo[0, 1] = 0
    orc._navigate(0)
    assert np.allclose(orc.setpoint[0, 1], [0.6, 0, 0, 0])           # SacrificeAttack


def test_quaternion_conventions():
    e = np.array([0.3, -0.2, 1.1])
    q = dy.quat_from_euler(e)
    assert np.allclose(dy.euler_from_quat(q), e)
    # rotateVector(q, v) with yaw 90 deg maps x -> y
    q = dy.quat_from_euler(np.array([0, 0, np.pi / 2]))
    assert np.allclose(dy.rotate_vector(q, np.array([1.0, 0, 0])), [0, 1, 0], atol=1e-12)


def test_command_to_setpoint():
    # quadcopter.py:379-396
    assert np.allclose(dy.command_to_setpoint(np.array([3.0, 0, 4.0, 0.5])), [0.3, 0, 0, 0.4])
    assert np.allclose(dy.command_to_setpoint(np.array([0.0, 0, 0, 0.4])), [0, 0, 0, 0])
