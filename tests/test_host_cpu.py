"""CPU-side checks: the C-ABI library loads and exports what include/dronechase_b200.h declares,
struct layouts agree between Python and C, and the product's config mirrors the oracle's."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_exports():
    import __graft_entry__ as g
    g.build()
    from dronechase_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dronechase_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|size_t|uint64_t|const char\*)\s+(dc_\w+)\(", hdr, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} not exported by libdronechase_b200.so"


def test_struct_layout_matches_header():
    """sizeof/offsetof of dc_config and dc_buffers as the C compiler sees them."""
    from dronechase_b200 import _lib
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "dronechase_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(dc_config), offsetof(dc_config, seed), offsetof(dc_config, dome_radius),
               offsetof(dc_config, building), offsetof(dc_config, quad), sizeof(dc_buffers),
               offsetof(dc_config, respawn_r_min), offsetof(dc_config, support_munition),
               offsetof(dc_config, initial_invaders), offsetof(dc_config, max_rounds), offsetof(dc_buffers, obs_mask));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    got = [C.sizeof(_lib.dc_config), _lib.dc_config.seed.offset, _lib.dc_config.dome_radius.offset,
           _lib.dc_config.building.offset, _lib.dc_config.quad.offset, C.sizeof(_lib.dc_buffers),
           _lib.dc_config.respawn_r_min.offset, _lib.dc_config.support_munition.offset,
           _lib.dc_config.initial_invaders.offset, _lib.dc_config.max_rounds.offset, _lib.dc_buffers.obs_mask.offset]
    assert [int(v) for v in out] == got


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dronechase_b200 import BatchedThreatEngageEnv, DroneChaseError
    with pytest.raises(DroneChaseError):
        BatchedThreatEngageEnv("exp02_vFinal", n_envs=2)
    # and the C ABI itself refuses, it does not fall back
    from dronechase_b200 import _lib
    cfg = _lib.dc_config(); cfg.abi_version = _lib.DC_ABI_VERSION; cfg.n_envs = 1; cfg.n_lw = 1; cfg.n_lm = 6; cfg.initial_round = 1; cfg.substeps = 16
    sim = C.c_void_p()
    assert _lib.lib().dc_create(C.byref(cfg), 0, C.byref(sim)) == -4
    assert b"no CPU fallback" in _lib.lib().dc_last_error()


def test_config_mirrors_oracle():
    from dronechase_b200.config import PRESETS, preset, quad_param_vector, calculate_rounds, CF2X
    from oracle import dynamics as dy
    from oracle.env_oracle import PRESETS as OP
    assert np.array_equal(quad_param_vector(CF2X), dy.QuadParams().flat())
    assert len(quad_param_vector(CF2X)) == 88
    assert calculate_rounds(1, 20) == 6 and calculate_rounds(2, 20) == 9
    for name in ("exp02_vFinal", "exp03_vFinal", "exp04_vFinal", "exp02_v2_full", "swarm"):
        p, o = preset(name), OP[name]
        for f in ("n_lw", "n_lm", "munition", "born_radius", "lw_spawn_radius", "explosion_range", "shoot_range",
                  "step_increment", "max_step", "initial_round", "cooldown_steps", "fire_probability", "lm_speed",
                  "bt_speed", "lm_nav", "ally_mode", "ally_stop_mag", "reward", "vel_bonus", "fixed_lw_spawn", "lidar"):
            assert getattr(p, f) == getattr(o, f), (name, f)
        assert tuple(p.building) == tuple(o.building) and p.substeps == o.substeps == 16
    from oracle.level5_oracle import LEVEL5_C1
    p = preset("level5_c1")
    for f in ("n_lw", "n_lm", "munition", "born_radius", "lw_spawn_radius", "explosion_range", "shoot_range", "step_increment",
              "max_step", "initial_round", "cooldown_steps", "fire_probability", "lm_speed", "bt_speed", "lm_nav", "ally_mode",
              "initial_invaders", "invaders_per_round", "max_rounds"):
        assert getattr(p, f) == getattr(LEVEL5_C1, f), ("level5_c1", f)
    assert p.family == "level5" and 2 * p.dome_radius == LEVEL5_C1.lidar_radius


def test_compat_module_paths_resolve():
    """The reference's import paths exist under compat/ and point at the facades (no GPU needed to import)."""
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        for mod, cls in (("threatengage.environments.level4.exp02_vFinal_environment", "Exp02vFinalEnvironment"),
                         ("threatengage.environments.level4.exp03_vFinal_environment", "Exp03vFinalEnvironment"),
                         ("threatengage.environments.level4.exp04_vFinal_environment", "Exp04vFinalEnvironment"),
                         ("threatengage.environments.level3.pyflyt_level3_environment_v2", "PyflytL3EnviromentV2"),
                         ("threatengage.environments.level2.pyflyt_level2_environment_modified_v2", "PyflytL2EnviromentModifiedV2"),
                         ("threatsense.level5.level5_c1_fusion_environment", "Level5C1FusionEnvironment"),
                         ("threatsense.level5.level5_fusion_environment", "Level5FusionEnvironment"),
                         ("threatsense.level5.level5_dumb_multiobs", "Level5DumbMultiObs"),
                         ("threatsense.level5.level5_eval_2bt_environment", "Level52BTEvaluationEnvironment"),
                         ("core.rl_framework.utils.io_data", "IOData")):
            m = importlib.import_module(mod)
            assert hasattr(m, cls)
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for k in [k for k in sys.modules if k.split(".")[0] in ("threatengage", "threatsense", "core")]:
            del sys.modules[k]


def test_host_scatter_sphere_matches_numpy():
    """dc_host_scatter_sphere (host-only helper of the VecEnv adapter) against a plain numpy rebuild."""
    from dronechase_b200 import _lib
    L = _lib.lib()
    rng = np.random.RandomState(0)
    for C_, n_lw, D in ((3, 1, 7), (2, 2, 12)):
        E = 257
        def draw():
            h = np.full((E, D, 2), -1, dtype=np.int32)
            for e in range(E):
                k = rng.randint(0, D)
                slots = rng.choice(np.arange(1, D), size=min(k, D - 1), replace=False)
                cells = rng.choice(338, size=len(slots), replace=False)      # winners hold distinct cells
                h[e, slots, 0] = cells
                h[e, slots, 1] = rng.uniform(0.01, 0.9, len(slots)).astype(np.float32).view(np.int32)
            return h
        def dense_of(h):
            out = np.ones((E, C_, 338), dtype=np.float32)
            for e in range(E):
                for d in range(D):
                    c = h[e, d, 0]
                    if c >= 0:
                        out[e, 0, c] = h[e, d, 1:2].view(np.float32)[0]
                        out[e, 1, c] = np.float32(0.6 if d < n_lw else 0.2)
                        if C_ == 3:
                            out[e, 2, c] = np.float32(0.1)
            return out
        prev = np.full((E, D, 2), -1, dtype=np.int32)
        dense = np.ones((E, C_, 13, 26), dtype=np.float32)
        def evolve(h):
            # what a real step looks like: most holders keep their cell (new distance), one cell changes holder,
            # one holder moves to a free cell, one disappears
            n = h.copy()
            for e in range(E):
                held = np.flatnonzero(n[e, :, 0] >= 0)
                n[e, held, 1] = rng.uniform(0.01, 0.9, len(held)).astype(np.float32).view(np.int32)
                free_slots = [d for d in range(1, D) if n[e, d, 0] < 0]
                if len(held) and free_slots and rng.rand() < 0.5:
                    a = rng.choice(held); b2 = rng.choice(free_slots)
                    n[e, b2] = n[e, a]; n[e, a] = (-1, np.float32(1.0).view(np.int32))
                held = np.flatnonzero(n[e, :, 0] >= 0)
                if len(held) and rng.rand() < 0.5:
                    free_cells = np.setdiff1d(np.arange(338), n[e, held, 0])
                    n[e, rng.choice(held), 0] = rng.choice(free_cells)
                held = np.flatnonzero(n[e, :, 0] >= 0)
                if len(held) and rng.rand() < 0.3:
                    n[e, rng.choice(held)] = (-1, np.float32(1.0).view(np.int32))
            return n
        for it in range(7):
            cur = draw() if it in (0, 4) else evolve(prev)
            rc = L.dc_host_scatter_sphere(dense.ctypes.data, prev.ctypes.data, cur.ctypes.data, E, D, n_lw, C_, 1 + it)
            assert rc == 0
            assert np.array_equal(dense.reshape(E, C_, 338), dense_of(cur)), (C_, it)
            prev = cur


def test_builtin_quad_table_is_the_python_default():
    """The cf2x table folded into the float32 dynamics (csrc/quad_dynamics.cuh CF2X_FLAT) is config.CF2X word for word: an
    edit on either side would silently send every preset down the run-time-constants instantiation."""
    import copy
    import numpy as np
    from dronechase_b200 import _lib
    from dronechase_b200.config import CF2X, quad_param_vector
    L = _lib.lib()
    q = np.ascontiguousarray(quad_param_vector(CF2X))
    assert L.dc_quad_is_builtin(q.ctypes.data_as(C.c_void_p)) == 1
    q2 = np.ascontiguousarray(quad_param_vector(CF2X, noise_ratio=0.0, ground_z=-1e9))      # run-time words: still the folded model
    assert L.dc_quad_is_builtin(q2.ctypes.data_as(C.c_void_p)) == 1
    other = copy.deepcopy(CF2X); other["mass"] = 0.030
    q3 = np.ascontiguousarray(quad_param_vector(other))
    assert L.dc_quad_is_builtin(q3.ctypes.data_as(C.c_void_p)) == 0
    q4 = np.ascontiguousarray(quad_param_vector(CF2X, gyro_term=True))
    assert L.dc_quad_is_builtin(q4.ctypes.data_as(C.c_void_p)) == 0


def test_host_apply_pairs():
    """dc_host_apply_pairs (the host half of the change-list sphere transfer, include/dronechase_b200.h): dense[index] = value."""
    import numpy as np
    from dronechase_b200 import _lib
    L = _lib.lib()
    rng = np.random.RandomState(0)
    dense = np.ones(1 << 20, dtype=np.float32)
    for n, threads in ((0, 1), (100, 1), (50000, 4)):
        idx = rng.choice(dense.size, n, replace=False).astype(np.int32)
        val = rng.rand(n).astype(np.float32)
        pairs = np.empty((n, 2), dtype=np.int32); pairs[:, 0] = idx; pairs[:, 1] = val.view(np.int32)
        want = dense.copy(); want[idx] = val
        assert L.dc_host_apply_pairs(dense.ctypes.data_as(C.c_void_p), pairs.ctypes.data_as(C.c_void_p), n, threads) == 0
        assert np.array_equal(dense, want)
