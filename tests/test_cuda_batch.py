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


def test_batched_env_matches_single_rollout_envs():
    """Two rollouts with different actions in one handle + bulk coupling == two separate TaichiEnv episodes."""
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    from softmac_b200.engine.taichi_env import TaichiEnv
    from softmac_b200.engine.batched_env import BatchedTaichiEnv
    from softmac_b200.engine.rigid_simulator import RigidSimulator
    from softmac_b200.engine.losses import PointwiseLoss
    from softmac_b200.config import CfgNode
    n, B, env_steps, substeps = 2000, 2, 4, 5
    n_grid, dt = 32, 2e-4
    max_steps = env_steps * substeps + substeps + 2
    rng = np.random.default_rng(5)
    x = ((rng.random((n, 3)) * 2 - 1) * 0.05 + np.array([0.5, 0.3, 0.5])).astype(np.float32).astype(np.float64)
    tab = scenes.sphere_table(radius=0.06, dx=0.01, margin=0.04)
    bodies = [dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 - 0.108, 0.3, 0.5), mass=1.0, gravity=False),
              dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 + 0.108, 0.3, 0.5), mass=1.0, gravity=False)]
    rcfg = CfgNode(gravity=(0., 0., 0.), init_state=(0., 0., 0.4, -0.4), bodies=bodies)
    target = x + np.array([0.0, 0.01, 0.0])
    actions = np.stack([np.tile([40.0, -40.0], (env_steps, 1)), np.tile([10.0, -70.0], (env_steps, 1))])     # (B, steps, 2)

    def build(nb):
        ms = [Mesh(sdf=dict(sdf=tab["sdf"], normal=tab["normal"], position=(tab["lower"], tab["upper"]), dx=tab["dx"]), cfg=dict(friction=0.3),
                   max_timesteps=max_steps) for _ in range(2)]
        prims = Primitives(primitives=ms, max_timesteps=max_steps)
        sim = MPMSimulator(sim_cfg(n, n_grid=n_grid, max_steps=max_steps, dt=dt), prims, env_dt=dt * substeps, n_batch=nb)
        return sim, prims

    # batched
    sim, prims = build(B)
    env = BatchedTaichiEnv(sim, prims, lambda b, views: RigidSimulator(rcfg, views, substeps=substeps, env_dt=dt * substeps), x)
    sim.clear_all_gradients()
    for k in range(env_steps):
        env.step(actions[:, k])
    f_end = env_steps * substeps
    xs = sim.get_x(f_end).reshape(B, n, 3)
    sim.add_x_grad(f_end, (xs - target).reshape(B * n, 3))
    gb = env.backward()                                                     # (B, steps, 2)
    rb = env.rigid_states()
    assert env.vec is not None                                              # prismatic joints: vectorised stand-in

    # one env per rollout
    for b in range(B):
        s1, p1 = build(1)
        rigid = RigidSimulator(rcfg, p1, substeps=substeps, env_dt=dt * substeps)
        e1 = TaichiEnv(s1, p1, rigid, x, loss=PointwiseLoss(s1, target))
        s1.clear_all_gradients()
        for k in range(env_steps):
            e1.step(actions[b, k])
        e1.compute_loss(f_end)
        g1 = e1.backward()
        assert rel_l2(xs[b], s1.get_x(f_end)) <= 1e-6
        assert rel_l2(rb[b], rigid.states[-1]) <= 1e-6
        assert np.abs(g1).max() > 0 and rel_l2(gb[b], g1) <= 1e-3, (gb[b], g1)


@pytest.mark.parametrize("joints", ["prismatic", "free", "revolute"])
def test_device_resident_rigid_coupling_matches_host_bridge(joints):
    """smx_rigid_linear_* (the rigid stand-in on the GPU, no host round trip per env step) == the host bridges: same particle
    states, rigid states, action gradients and adjoint of the initial rigid state.  "prismatic": two fingers + a fixed body whose
    wrench is ignored (enable_external_force = False, rigid_simulator.py:96), checked against LinearBatchedRigid; "free": two
    free-floating bodies with spin (demo_pour's joints: the pose map is nonlinear), checked against one Python bridge per rollout;
    "revolute": a hinged slab swinging into the material about a vertical axis plus a fixed body (the door scene's joint,
    config/demo_door_config.py:31-56), closed-form pose Jacobian on both sides."""
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    from softmac_b200.engine.batched_env import BatchedTaichiEnv
    from softmac_b200.engine.rigid_simulator import RigidSimulator
    from softmac_b200.config import CfgNode
    n, B, env_steps, substeps = 2000, 3, 5, 5
    n_grid, dt = 32, 2e-4
    max_steps = env_steps * substeps + substeps + 2
    rng = np.random.default_rng(6)
    x = ((rng.random((n, 3)) * 2 - 1) * 0.05 + np.array([0.5, 0.3, 0.5])).astype(np.float32).astype(np.float64)
    tab = scenes.sphere_table(radius=0.06, dx=0.01, margin=0.04)
    if joints == "prismatic":
        bodies = [dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 - 0.108, 0.3, 0.5), mass=1.0, gravity=False),
                  dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 + 0.108, 0.3, 0.5), mass=1.5, gravity=False),
                  dict(joint="fixed", origin=(0.5, 0.3 + 0.105, 0.5))]
        init = (0., 0., 0.4, -0.4)
        actions = np.stack([np.tile([40.0, -40.0], (env_steps, 1)), np.tile([10.0, -70.0], (env_steps, 1)), np.tile([0.0, 0.0], (env_steps, 1))])
        enable = [True, True, False]
    elif joints == "revolute":
        tab = scenes.box_table(half=(0.10, 0.04, 0.03), dx=0.01, margin=0.04)
        # the slab lies along x beside the particle cloud (its +z face just short of the cloud's boundary z = 0.45 for x in [0.45, 0.50]: no
        # particle starts inside it, which would eject it at sdf / dt) and swings into the cloud about the vertical hinge at its centre:
        # z' = -x sin(theta), so omega < 0 moves the free end towards +z
        bodies = [dict(joint="revolute", axis=(0, 1, 0), origin=(0.40, 0.3, 0.42), inertia=2e-3, gravity=False),
                  dict(joint="fixed", origin=(0.5, 0.3 + 0.12, 0.5))]
        init = (0.002, -4.0)
        actions = np.stack([np.tile([-0.5], (env_steps, 1)), np.tile([0.8], (env_steps, 1)), np.tile([0.0], (env_steps, 1))])
        enable = [True, False]
    else:
        bodies = [dict(joint="free", origin=(0.5 - 0.106, 0.3, 0.5), quat=(0.9238795, 0.0, 0.3826834, 0.0), mass=1.0, inertia=0.02, gravity=True),
                  dict(joint="free", origin=(0.5 + 0.106, 0.3, 0.5), mass=2.0, inertia=0.05, gravity=False)]
        init = (0.1, -0.2, 0.05, 0.0, 0.0, 0.0) + (0.0,) * 6 + (3.0, 1.0, -2.0, 0.4, 0.0, 0.0) + (0.0, 0.5, 0.0, -0.4, 0.0, 0.0)
        a0 = np.array([0.02, 0.0, 0.05, 30.0, 5.0, 0.0, 0.0, 0.01, 0.0, -40.0, 0.0, 3.0])
        actions = np.stack([np.tile(a0 * sc, (env_steps, 1)) for sc in (1.0, -0.5, 0.0)])
        enable = [True, False]
    rcfg = CfgNode(gravity=(0., -9.8, 0.), init_state=init, bodies=bodies)
    target = x + np.array([0.0, 0.01, 0.0])

    def run(device_rigid):
        ms = [Mesh(sdf=dict(sdf=tab["sdf"], normal=tab["normal"], position=(tab["lower"], tab["upper"]), dx=tab["dx"]),
                   cfg=dict(friction=0.3, enable_external_force=enable[i]), max_timesteps=max_steps) for i in range(len(bodies))]
        prims = Primitives(primitives=ms, max_timesteps=max_steps)
        sim = MPMSimulator(sim_cfg(n, n_grid=n_grid, max_steps=max_steps, dt=dt), prims, env_dt=dt * substeps, n_batch=B)
        env = BatchedTaichiEnv(sim, prims, lambda b, views: RigidSimulator(rcfg, views, substeps=substeps, env_dt=dt * substeps), x,
                               device_rigid=device_rigid)
        out = []
        for it in range(2):                                                 # the second episode checks reset()
            env.reset()
            sim.clear_all_gradients()
            for k in range(env_steps):
                env.step(actions[:, k])
            f_end = env_steps * substeps
            xs = sim.get_x(f_end).reshape(B, n, 3)
            sim.add_x_grad(f_end, (xs - target).reshape(B * n, 3))
            g = env.backward()
            sg = env.dev.state_grad if env.dev else (env.vec.state_grad if env.vec else np.stack([r.state_grad for r in env.rigid]))
            out.append((xs, g, env.rigid_states(), np.array(sg)))
        assert rel_l2(out[0][1], out[1][1]) <= 1e-4                       # episodes differ only by the order of the float atomics
        return out[1]

    xh, gh, rh, sh = run(False)
    xd, gd, rd, sd = run(True)
    assert gh.shape == gd.shape == (B, env_steps, actions.shape[2]) and np.abs(gh[:2]).max() > 0
    assert rel_l2(xd, xh) <= 1e-6
    assert rel_l2(rd, rh) <= 1e-7
    assert rel_l2(gd, gh) <= 1e-4, (gd, gh)
    assert rel_l2(sd, sh) <= 1e-4, (sd, sh)


def test_device_chamfer_loss_matches_host_restatement():
    """smx_chamfer_loss vs the numpy restatement of loss_grip.py:45-68 (value and seed), single and batched handles."""
    from softmac_b200.engine import MPMSimulator
    from softmac_b200.engine.losses import ChamferLoss, DeviceChamferLoss
    rng = np.random.default_rng(31)
    n = 3000
    st = scenes.blob_state(n, rng)
    target = st[:, :3] * np.array([0.8, 1.1, 1.0]) + 0.01 * rng.normal(size=(n, 3))
    sim = MPMSimulator(sim_cfg(n, max_steps=4), (), env_dt=1e-3)
    sim.reset(st)
    sim.substep(0)
    host = ChamferLoss(sim, target, weight=0.7)
    lh = host.compute_loss(1)["loss"]
    gh = sim.get_state_grad(1)[:, :3]
    sim.clear_all_gradients()
    ld = DeviceChamferLoss(sim, target, weight=0.7).compute_loss(1)["loss"]
    gd = sim.get_state_grad(1)[:, :3]
    assert abs(ld - lh) <= 1e-5 * abs(lh)
    assert np.abs(gh).max() > 0 and rel_l2(gd, gh) <= 1e-5
    # batched: two rollouts with different states, loss = sum over rollouts, seeds per rollout
    st2 = np.vstack([st, scenes.blob_state(n, np.random.default_rng(32))])
    simb = MPMSimulator(sim_cfg(n, max_steps=4), (), env_dt=1e-3, n_batch=2)
    simb.reset(st2)
    lb = DeviceChamferLoss(simb, target, weight=1.0).compute_loss(0)["loss"]
    gb = simb.get_state_grad(0)[:, :3].reshape(2, n, 3)
    ref_l, ref_g = 0.0, []
    for b in range(2):
        s1 = MPMSimulator(sim_cfg(n, max_steps=4), (), env_dt=1e-3)
        s1.reset(st2[b * n:(b + 1) * n])
        ref_l += ChamferLoss(s1, target).compute_loss(0)["loss"]
        ref_g.append(s1.get_state_grad(0)[:, :3])
    assert abs(lb - ref_l) <= 1e-5 * abs(ref_l)
    assert rel_l2(gb[0], ref_g[0]) <= 1e-5 and rel_l2(gb[1], ref_g[1]) <= 1e-5


def test_cuda_graph_replay_of_env_steps_is_identical():
    """smx_step_graph / smx_step_grad_graph (every env step's substeps captured, patched into the handle's executable graph and launched
    as ONE graph) == the ordinary launch sequence: same states and action gradients over an episode with re-sorts, forecast contact and the
    device-resident rigid coupling; the calls that must take the ordinary path (first after reset, re-sort inside, first of the backward
    pass) do so, the others are graph launches."""
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    from softmac_b200.engine.batched_env import BatchedTaichiEnv
    from softmac_b200.engine.rigid_simulator import RigidSimulator
    from softmac_b200.config import CfgNode
    n, B, env_steps, substeps = 2000, 2, 12, 5
    n_grid, dt = 32, 2e-4
    max_steps = env_steps * substeps + substeps + 2
    rng = np.random.default_rng(8)
    x = ((rng.random((n, 3)) * 2 - 1) * 0.05 + np.array([0.5, 0.3, 0.5])).astype(np.float32).astype(np.float64)
    tab = scenes.sphere_table(radius=0.06, dx=0.01, margin=0.04)
    bodies = [dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 - 0.115, 0.3, 0.5), mass=1.0, gravity=False),
              dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 + 0.115, 0.3, 0.5), mass=1.5, gravity=False)]
    rcfg = CfgNode(gravity=(0., 0., 0.), init_state=(0., 0., 0.3, -0.3), bodies=bodies)
    actions = np.stack([np.tile([20.0, -20.0], (env_steps, 1)), np.tile([5.0, -30.0], (env_steps, 1))])
    target = x + np.array([0.0, 0.01, 0.0])

    def run(graphs):
        ms = [Mesh(sdf=dict(sdf=tab["sdf"], normal=tab["normal"], position=(tab["lower"], tab["upper"]), dx=tab["dx"]), cfg=dict(friction=0.3), max_timesteps=max_steps)
              for _ in bodies]
        prims = Primitives(primitives=ms, max_timesteps=max_steps)
        sim = MPMSimulator(sim_cfg(n, n_grid=n_grid, max_steps=max_steps, dt=dt), prims, env_dt=dt * substeps, n_batch=B, sort_every=15)
        sim.use_graphs = graphs
        env = BatchedTaichiEnv(sim, prims, lambda b, views: RigidSimulator(rcfg, views, substeps=substeps, env_dt=dt * substeps), x, device_rigid=True)
        for it in range(2):
            env.reset()
            sim.clear_all_gradients()
            for k in range(env_steps):
                env.step(actions[:, k])
            f_end = env_steps * substeps
            xs = sim.get_x(f_end).reshape(B, n, 3)
            sim.add_x_grad(f_end, (xs - target).reshape(B * n, 3))
            g = env.backward()
        return xs, g, env.rigid_states(), sim.graph_status(), sim.counters()

    xp, gp, rp, sp, cp = run(False)
    xg, gg, rg, sg, cg = run(True)
    assert sp["graph_calls"] == 0 and cp["resorts"] >= 3
    assert sg["graph_calls"] >= env_steps and sg["plain_calls"] >= 4, sg         # both episodes, forward and backward
    assert np.abs(gp).max() > 0
    assert rel_l2(xg, xp) <= 1e-6 and rel_l2(rg, rp) <= 1e-7
    assert rel_l2(gg, gp) <= 1e-4, (gg, gp)
