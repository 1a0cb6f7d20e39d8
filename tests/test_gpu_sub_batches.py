"""dc_config.sub_batches: the env batch split into independent sub-batches on internal streams must give the
same bits as one batch (envs are independent, Philox streams are keyed by the global env index)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(name, E, K, steps, with_hits=False):
    from dronechase_b200 import BatchedThreatEngageEnv
    env = BatchedThreatEngageEnv(name, n_envs=E, seed=11, device=0, sub_batches=K, with_ids=True, with_terminal_obs=True,
                                 with_hits=with_hits)
    env.reset()
    rng = np.random.RandomState(3)
    out = []
    for t in range(steps):
        a = np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(0, 1, (E, 1))], axis=1).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        if t % 7 == 0 or t == steps - 1:
            out.append({**{k: v.cpu().numpy().copy() for k, v in obs.items()}, "reward": rew.cpu().numpy().copy(),
                        "done": done.cpu().numpy().copy(), "info": info.cpu().numpy().copy(),
                        "ids": env.lidar_ids.cpu().numpy().copy()})
        if t % 20 == 0:                                   # masked resets across the sub-batch borders
            robs = env.reset(torch.arange(E) % 3 == (t // 20) % 3)
            out.append({k: v.cpu().numpy().copy() for k, v in robs.items()})
    st = env.get_state()
    stats = env.stats.cpu().numpy().copy()
    env.close()
    return out, st, stats


@pytest.mark.parametrize("name,E,steps", [("exp02_vFinal", 200, 260), ("exp03_vFinal", 330, 120), ("level5_c1", 200, 90)])
def test_sub_batches_bit_identical(name, E, steps):
    ref, st_ref, stats_ref = _run(name, E, 1, steps)
    got, st_got, stats_got = _run(name, E, 3, steps)
    for a, b in zip(ref, got):
        for k in a:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    for k in st_ref:
        np.testing.assert_array_equal(st_ref[k], st_got[k], err_msg=k)
    np.testing.assert_allclose(stats_ref, stats_got, rtol=1e-12)      # float64 atomics: order differs


def test_sub_batches_masked_reset_and_set_state():
    from dronechase_b200 import BatchedThreatEngageEnv
    E = 256
    a = BatchedThreatEngageEnv("exp02_vFinal", n_envs=E, seed=5, device=0, sub_batches=1)
    b = BatchedThreatEngageEnv("exp02_vFinal", n_envs=E, seed=5, device=0, sub_batches=4)
    a.reset(); b.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(40):
        act = torch.rand(E, 4, device="cuda", generator=g); act[:, :3] = act[:, :3] * 2 - 1
        a.step(act); b.step(act)
    mask = (torch.arange(E, device="cuda") % 5 == 0)
    a.reset(mask); b.reset(mask)
    for k in a.obs:
        assert torch.equal(a.obs[k], b.obs[k]), k
    # state written through the split path reads back identically and steps identically
    st = a.get_state()
    b.set_state(st)
    st2 = b.get_state()
    for k in st:
        np.testing.assert_array_equal(st[k], st2[k], err_msg=k)
    for t in range(10):
        act = torch.rand(E, 4, device="cuda", generator=g); act[:, :3] = act[:, :3] * 2 - 1
        a.step(act); b.step(act)
    assert torch.equal(a.reward, b.reward) and torch.equal(a.done, b.done) and torch.equal(a.info, b.info)
    for k in a.obs:
        assert torch.equal(a.obs[k], b.obs[k]), k
    a.close(); b.close()


@pytest.mark.parametrize("name,K", [("exp02_vFinal", 4), ("level5_c1", 2)])
def test_graph_stepping_is_bit_identical(name, K):
    """step_graph (two captured graphs replayed alternately, dc_note_graph_replay) == step, also across a masked reset
    and when the two ways of stepping are mixed."""
    from dronechase_b200 import BatchedThreatEngageEnv
    E = 512
    a = BatchedThreatEngageEnv(name, n_envs=E, seed=9, device=0, sub_batches=1)
    b = BatchedThreatEngageEnv(name, n_envs=E, seed=9, device=0, sub_batches=K)
    a.reset(); b.reset()
    b.capture_step_graphs()
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    for t in range(61):
        act = torch.rand(E, 4, device="cuda", generator=g); act[:, :3] = act[:, :3] * 2 - 1
        a.step(act)
        if t % 10 == 7:
            b.step(act)                      # a plain step in between keeps the graph pair aligned
        else:
            b.step_graph(act)
        if t == 30:
            mask = torch.arange(E, device="cuda") % 4 == 1
            a.reset(mask); b.reset(mask)
        assert torch.equal(a.reward, b.reward) and torch.equal(a.done, b.done) and torch.equal(a.info, b.info), t
        for k in a.obs:
            assert torch.equal(a.obs[k], b.obs[k]), (t, k)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
    a.close(); b.close()


@pytest.mark.parametrize("name,kw,attr", [("level5_fusion", {"with_student": True, "with_hits": True}, "student_obs"),
                                          ("level5_dumb_multiobs", {"with_hits": True}, "multi_obs")])
def test_sub_batches_bit_identical_student_and_multi_observer(name, kw, attr):
    """The student stack (dc_buffers.student_*) and the multi-observer tensors (dc_buffers.mo_*) are per-env rows like
    every other output: three sub-batches give the bits of one batch, hit lists included."""
    from dronechase_b200 import BatchedThreatEngageEnv
    E, steps = 90, 70
    envs = [BatchedThreatEngageEnv(name, n_envs=E, seed=13, device=0, sub_batches=k, **kw) for k in (1, 3)]
    for e in envs:
        e.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    marked = 0
    for t in range(steps):
        act = torch.rand(E, 4, device="cuda", generator=g); act[:, :3] = act[:, :3] * 2 - 1
        for e in envs:
            e.step(act)
        a, b = (getattr(e, attr) for e in envs)
        for k in a:
            assert torch.equal(a[k], b[k]), f"step {t}: {k}"
        ha, hb = ((e.student_hits if attr == "student_obs" else e.multi_hits) for e in envs)
        assert torch.equal(ha, hb), f"step {t}: hit lists"
        marked += int((a["stacked_spheres"] < 1).sum())
        if t == 40:
            m = torch.arange(E, device="cuda") % 4 == 0
            for e in envs:
                e.reset(m)
    assert marked > 1000
    # the hit list is a complete sparse description of the dense tensor (dc_scatter_stack rebuilds it bit for bit)
    import ctypes as C
    from dronechase_b200 import _lib
    e = envs[1]
    hits = e.student_hits if attr == "student_obs" else e.multi_hits
    dense = getattr(e, attr)["stacked_spheres"]
    rows = hits.numel() // (hits.shape[-2] * 2)
    out = torch.empty_like(dense)
    _lib.check(_lib.lib().dc_scatter_stack(C.c_void_p(hits.data_ptr()), None, rows, e.cfg.n_drones, C.c_void_p(out.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "dc_scatter_stack")
    assert torch.equal(out, dense)
    for e in envs:
        e.close()
