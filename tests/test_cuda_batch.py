"""n_batch > 1: several independent rollouts in one handle (BASELINE config 4) must equal the same rollouts run in
separate handles -- states, per-rollout wrenches, adjoints and per-rollout primitive gradients."""
import numpy as np
import pytest

import scenes
from harness import sim_cfg, rel_l2

pytestmark = pytest.mark.gpu


def make(n, B, max_steps, tab, sort_every=3):
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    m = Mesh(sdf=dict(sdf=tab["sdf"], normal=tab["normal"], position=(tab["lower"], tab["upper"]), dx=tab["dx"]), cfg=dict(friction=0.5),
             max_timesteps=max_steps)
    prims = Primitives(primitives=[m], max_timesteps=max_steps)
    sim = MPMSimulator(sim_cfg(n, max_steps=max_steps), prims, env_dt=1e-3, sort_every=sort_every, n_batch=B)
    prims.initialize()
    return sim, prims


def test_batched_rollouts_match_separate_handles():
    n, B, steps = 2500, 3, 7
    center = np.array([0.5, 0.3, 0.5])
    tab = scenes.sphere_table()
    rng = np.random.default_rng(11)
    states = [scenes.contact_rollout_state(n, np.random.default_rng(100 + b), center, speed=0.5 + 0.5 * b) for b in range(B)]
    poses = [np.concatenate([center + [0.004 * b, 0, 0], [1, 0, 0, 0], [0, 0.1 * (b + 1), 0], [0, 0, 0.2 * b]]) for b in range(B)]
    seeds = [rng.normal(size=(n, 3)) for _ in range(B)]
    ext = [rng.normal(size=6) * 1e-3 for _ in range(B)]

    # one batched handle
    sim, prims = make(n, B, steps + 2, tab)
    views = [prims.view(b) for b in range(B)]
    for b in range(B):
        views[b][0].set_all_states(0, poses[b], f_end=steps + 2)
    sim.reset(np.vstack(states))
    for b in range(B):
        views[b][0].clear_ext_f()
    sim.step(0, steps)
    got_state = sim.get_state(steps).reshape(B, n, 24)
    got_extf = [views[b][0].get_ext_f() for b in range(B)]
    sim.clear_all_gradients()
    sim.add_x_grad(steps, np.vstack(seeds))
    for f in range(steps - 1, -1, -1):
        for b in range(B):
            views[b][0].set_ext_f_grad(ext[b])
        sim.substep_grad(f)
    got_adj = sim.get_state_grad(0).reshape(B, n, 24)
    got_pg = [views[b][0].get_all_states_grad(0, f_end=steps) for b in range(B)]
    assert sim.counters()["resorts"] >= 2

    # the same rollouts, one handle each
    for b in range(B):
        s1, p1 = make(n, 1, steps + 2, tab)
        p1[0].set_all_states(0, poses[b], f_end=steps + 2)
        s1.reset(states[b])
        p1[0].clear_ext_f()
        s1.step(0, steps)
        ref = s1.get_state(steps)
        assert rel_l2(got_state[b][:, :3], ref[:, :3]) <= 1e-6
        assert rel_l2(got_state[b][:, 3:], ref[:, 3:]) <= 2e-5
        fe = p1[0].get_ext_f()
        assert np.abs(fe).max() > 0 and rel_l2(got_extf[b], fe) <= 1e-4
        s1.clear_all_gradients()
        s1.add_x_grad(steps, seeds[b])
        for f in range(steps - 1, -1, -1):
            s1.substep_grad(f, ext_f_grad=[ext[b]])
        assert rel_l2(got_adj[b], s1.get_state_grad(0)) <= 1e-4
        pg = p1[0].get_all_states_grad(0, f_end=steps)
        assert np.abs(pg).max() > 0 and rel_l2(got_pg[b], pg) <= 1e-3
    # rollouts differ from each other (the batches really are independent problems)
    assert rel_l2(got_state[0], got_state[1]) > 1e-3


def test_batch_broadcast_reset_and_permutation():
    n, B = 1000, 4
    sim, prims = make(n, B, 6, scenes.sphere_table())
    st = scenes.blob_state(n, np.random.default_rng(3))
    sim.reset(st)                                   # (n, 24) is broadcast to every rollout
    got = sim.get_state(0).reshape(B, n, 24)
    for b in range(B):
        assert np.array_equal(got[b], st)
    perm = sim.permutation(0).reshape(B, n)
    for b in range(B):                              # sorting never mixes rollouts
        assert perm[b].min() >= b * n and perm[b].max() < (b + 1) * n
    keys = sim.sort_keys(0)
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
