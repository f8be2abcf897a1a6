"""CUDA path vs the f64 oracle on identical (fp32-rounded) inputs, through the C ABI (libsoftmac_b200.so).

Tolerances (BASELINE.json north_star): per-substep relative L2 <= 1e-4 on x / v / F / C; gradients are held to
relative L2 <= 2e-3 and cosine >= 0.9999 per substep, cosine >= 0.999 over a multi-substep rollout.
Quantities whose oracle norm is ~0 (e.g. C after a rigid translation) use an absolute floor stated in the test.
"""
import numpy as np
import pytest

import scenes
from harness import Pair, rel_l2, cosine, prim_states_for, sim_cfg

pytestmark = pytest.mark.gpu

COLS = dict(x=slice(0, 3), v=slice(3, 6), F=slice(6, 15), C=slice(15, 24))
TOL = 1e-4


def assert_state_close(got, ref, tol=TOL, floors=None):
    floors = floors or {}
    for k, sl in COLS.items():
        e = rel_l2(got[:, sl], ref[:, sl], floor=floors.get(k, 0.0))
        assert e <= tol, f"{k}: rel L2 {e:.3e} > {tol}"


def make_pair(rng, n=4000, P=2, center=(0.5, 0.3, 0.5), **kw):
    tabs = [scenes.sphere_table() for _ in range(P)]
    params = [(0.4 + 0.3 * i, 666.) for i in range(P)]
    pair = Pair(n, tables=tabs, prim_params=params, **kw)
    st = scenes.blob_state(n, rng, center=center)
    prs = prim_states_for(rng, P, center)
    for i, s13 in enumerate(prs):
        pair.set_prim_state(i, 0, pair.cfg.max_steps, s13)
    pair.reset(st)
    pair.clear_ext_f()
    return pair, st


@pytest.mark.parametrize("ptype,material_model", [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1)])
def test_forward_substep_mixed_contact(ptype, material_model):
    rng = np.random.default_rng(100 + 3 * material_model + ptype)
    pair, st = make_pair(rng, ptype=ptype, material_model=material_model)
    pair.substep(0)
    ref, got = pair.orc.get_frame(1), pair.gpu.get_state(1)
    assert_state_close(got, ref)
    for i in range(pair.P):
        fo, fg = pair.orc.get_ext_f(i), pair.prims[i].get_ext_f()
        assert np.abs(fo).max() > 0, "scene must exercise contact"
        assert rel_l2(fg, fo) <= 1e-3, (fg, fo)


@pytest.mark.parametrize("collision_type", [0, 1])
def test_forward_substep_other_contact_models(collision_type):
    rng = np.random.default_rng(110 + collision_type)
    pair, st = make_pair(rng, collision_type=collision_type)
    pair.substep(0)
    assert_state_close(pair.gpu.get_state(1), pair.orc.get_frame(1))
    for i in range(pair.P):
        fo, fg = pair.orc.get_ext_f(i), pair.prims[i].get_ext_f()
        assert np.abs(fo).max() > 0
        assert rel_l2(fg, fo) <= 1e-3


def test_grid_matches_oracle():
    rng = np.random.default_rng(120)
    pair, st = make_pair(rng, P=0)
    pair.substep(0)
    gvin, gm, gvout = pair.orc.get_grid()
    a, b = pair.gpu.get_grid()
    assert rel_l2(a[:, 3], gm) <= 1e-5
    assert rel_l2(a[:, :3], gvin) <= 1e-4
    assert rel_l2(b[:, :3], gvout) <= 1e-4
    assert np.array_equal(b[:, 3] > 0, gm > 1e-10) or np.mean((b[:, 3] > 0) != (gm > 1e-10)) < 1e-4


def run_backward(pair, rng, f, seed_scale=1.0):
    n, P = pair.n, pair.P
    cot = rng.normal(size=(n, 24)) * seed_scale
    cot = cot.astype(np.float32).astype(np.float64)
    ext = [rng.normal(size=6).astype(np.float32).astype(np.float64) for _ in range(P)]
    pair.orc.clear_grads(); pair.gpu.clear_all_gradients()
    pair.orc.add_frame_grad(f + 1, cot); pair.gpu.add_state_grad(f + 1, cot)
    for i in range(P):
        pair.orc.set_ext_f_grad(i, ext[i])
    pair.orc.substep_grad(f)
    pair.gpu.substep_grad(f, ext_f_grad=ext)
    return pair.orc.get_frame_grad(f), pair.gpu.get_state_grad(f)


@pytest.mark.parametrize("ptype,material_model", [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1)])
def test_backward_substep_mixed_contact(ptype, material_model):
    rng = np.random.default_rng(200 + 3 * material_model + ptype)
    pair, st = make_pair(rng, ptype=ptype, material_model=material_model)
    pair.substep(0)
    go, gg = run_backward(pair, rng, 0)
    for k, sl in COLS.items():
        e, c = rel_l2(gg[:, sl], go[:, sl]), cosine(gg[:, sl], go[:, sl])
        assert e <= 2e-3 and c >= 0.9999, f"adjoint {k}: rel L2 {e:.3e}, cos {c:.6f}"
    for i in range(pair.P):
        po, pg = pair.orc.get_primitive_state_grad(i, 0), pair.prims[i].get_all_states_grad(0)
        assert np.abs(po).max() > 0
        assert rel_l2(pg, po) <= 5e-3 and cosine(pg, po) >= 0.9999, (pg, po)


def test_von_mises_return_mapping_forward_backward_and_rollout():
    """soft_cloth's plastic flow rule (soft_cloth/engine/mpm_simulator.py:172-189, :232) behind smx_set_plasticity: one substep forward and
    adjoint with yielding and non-yielding particles in the same blob, then a fused 10-substep rollout (smx_step / smx_step_grad)."""
    rng = np.random.default_rng(230)
    pair, st = make_pair(rng, max_steps=14)
    # yield stress at the median of the strain norm of frame 0 (numpy evaluation of the rule on F_tmp): half of the blob yields
    st = np.asarray(st, dtype=np.float32).astype(np.float64)
    F0, C0 = st[:, 6:15].reshape(-1, 3, 3), st[:, 15:24].reshape(-1, 3, 3)
    sg = np.linalg.svd((np.eye(3)[None] + pair.cfg.dt * C0) @ F0, compute_uv=False)
    eps = np.log(np.maximum(sg, 0.05)); eh = eps - eps.mean(1, keepdims=True)
    nrm = np.sqrt((eh * eh).sum(1) + 1e-8)
    mu = pair.cfg.E / (2 * (1 + pair.cfg.nu))
    ys = 2 * mu * float(np.median(nrm))
    pair.orc.set_plasticity(1, ys); pair.gpu.set_plasticity("von_mises", ys)
    # particles whose norm is within fp32 resolution of the threshold may take the other branch on the device
    yields, near = nrm - ys / (2 * mu) > 0, np.abs(nrm - ys / (2 * mu)) < 1e-6
    assert 0.3 < yields.mean() < 0.7 and near.mean() < 0.01
    pair.substep(0)
    ref, got = pair.orc.get_frame(1), pair.gpu.get_state(1)
    assert_state_close(got, ref)
    yy, nn = yields & ~near, ~yields & ~near
    assert rel_l2(got[yy, 6:15], ref[yy, 6:15]) <= 1e-5 and rel_l2(got[nn, 6:15], ref[nn, 6:15]) <= 1e-6
    go, gg = run_backward(pair, rng, 0)
    for k, sl in COLS.items():
        e, c = rel_l2(gg[:, sl], go[:, sl]), cosine(gg[:, sl], go[:, sl])
        assert e <= 2e-3 and c >= 0.9999, f"adjoint {k}: rel L2 {e:.3e}, cos {c:.6f}"
    for i in range(pair.P):
        po, pg = pair.orc.get_primitive_state_grad(i, 0), pair.prims[i].get_all_states_grad(0)
        assert rel_l2(pg, po) <= 5e-3 and cosine(pg, po) >= 0.9999, (pg, po)
    # rollout through the fused kernels (smx_step / smx_step_grad) on the scene of the rollout test below
    T, n, center = 10, 4000, np.array([0.5, 0.3, 0.5])
    pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=T + 2, sort_every=4, substeps=4)
    pair.set_prim_state(0, 0, T + 2, np.concatenate([center, scenes.random_quat(rng), [0.0, 0.2, 0.0], 0.5 * rng.normal(size=3)]))
    st = scenes.contact_rollout_state(n, rng, center)
    sg = np.linalg.svd((np.eye(3)[None] + pair.cfg.dt * st[:, 15:24].reshape(-1, 3, 3)) @ st[:, 6:15].reshape(-1, 3, 3), compute_uv=False)
    eps = np.log(sg); eh = eps - eps.mean(1, keepdims=True)
    ys = 2 * mu * float(np.median(np.sqrt((eh * eh).sum(1) + 1e-8)))
    pair.orc.set_plasticity(1, ys); pair.gpu.set_plasticity("von_mises", ys)
    pair.reset(st); pair.clear_ext_f()
    for f in range(T):
        pair.orc.substep(f)
    pair.gpu.step(0, T)
    ref, got = pair.orc.get_frame(T), pair.gpu.get_state(T)
    assert np.isfinite(ref).all()
    assert rel_l2(got[:, 0:3], ref[:, 0:3]) <= 1e-5 and rel_l2(got[:, 3:6], ref[:, 3:6]) <= 2e-3 and rel_l2(got[:, 6:15], ref[:, 6:15]) <= 1e-4
    cot = np.zeros((n, 24)); cot[:, :6] = rng.normal(size=(n, 6))
    cot = cot.astype(np.float32).astype(np.float64)
    pair.orc.clear_grads(); pair.gpu.clear_all_gradients()
    pair.orc.add_frame_grad(T, cot); pair.gpu.add_state_grad(T, cot)
    for f in range(T - 1, -1, -1):
        pair.orc.substep_grad(f)
    pair.gpu.step_grad(T, T)             # adjoint of substeps T-1 ... 0
    go, gg = pair.orc.get_frame_grad(0), pair.gpu.get_state_grad(0)
    assert np.abs(go[:, 6:15]).max() > 0 and cosine(gg, go) >= 0.999, cosine(gg, go)


@pytest.mark.parametrize("collision_type", [0, 1])
def test_backward_substep_other_contact_models(collision_type):
    rng = np.random.default_rng(210 + collision_type)
    pair, st = make_pair(rng, collision_type=collision_type)
    pair.substep(0)
    go, gg = run_backward(pair, rng, 0)
    for k, sl in COLS.items():
        e, c = rel_l2(gg[:, sl], go[:, sl]), cosine(gg[:, sl], go[:, sl])
        assert e <= 2e-3 and c >= 0.9999, f"adjoint {k}: rel L2 {e:.3e}, cos {c:.6f}"
    for i in range(pair.P):
        po, pg = pair.orc.get_primitive_state_grad(i, 0), pair.prims[i].get_all_states_grad(0)
        assert rel_l2(pg, po) <= 5e-3, (pg, po)


def test_near_isotropic_plastic_adjoint():
    """F = I + O(1e-4): the regime where a naive fp32 evaluation of backward_svd loses the gradient
    (hard part 2).  The deviation-form kernels must still match the f64 oracle."""
    rng = np.random.default_rng(220)
    n = 4000
    pair = Pair(n, ptype=0)
    st = scenes.blob_state(n, rng, Fdev=1e-4, Cdev=0.5)
    pair.reset(st)
    pair.substep(0)
    assert_state_close(pair.gpu.get_state(1), pair.orc.get_frame(1))
    go, gg = run_backward(pair, rng, 0)
    for k, sl in COLS.items():
        e, c = rel_l2(gg[:, sl], go[:, sl]), cosine(gg[:, sl], go[:, sl])
        assert e <= 2e-3 and c >= 0.9999, f"adjoint {k}: rel L2 {e:.3e}, cos {c:.6f}"


def test_identity_F_first_substep():
    """reset(x (n,3)): F = I exactly, C = 0 -> all singular values equal (the clamp branch of backward_svd)."""
    rng = np.random.default_rng(221)
    n = 3000
    pair = Pair(n, ptype=0)
    x = (rng.random((n, 3)) * 0.1 + np.array([0.45, 0.2, 0.45])).astype(np.float32).astype(np.float64)
    st = np.zeros((n, 24)); st[:, :3] = x; st[:, 6] = st[:, 10] = st[:, 14] = 1
    pair.orc.set_frame(0, st)
    pair.gpu.reset(x)
    pair.substep(0)
    # v = g*dt everywhere, so C = 0 analytically (|C| ~ 1e-12 in f64): measure the C error against the natural
    # scale of a velocity gradient, |v| / dx, instead of against ~0
    ref = pair.orc.get_frame(1)
    assert_state_close(pair.gpu.get_state(1), ref, floors=dict(C=np.linalg.norm(ref[:, 3:6]) * 32))
    go, gg = run_backward(pair, rng, 0)
    for k, sl in COLS.items():
        e = rel_l2(gg[:, sl], go[:, sl])
        assert e <= 2e-3, f"adjoint {k}: rel L2 {e:.3e}"


def test_control_action_gradient():
    rng = np.random.default_rng(230)
    n = 3000
    pair = Pair(n, n_control=2, ptype=1, gravity=(0., 0., 0.), ground_friction=0.)
    st = scenes.blob_state(n, rng)
    idx = rng.integers(-1, 2, size=n)
    action = rng.normal(size=(2, 3)) * 50
    action = action.astype(np.float32).astype(np.float64)
    pair.reset(st)
    pair.orc.set_control_idx(idx); pair.gpu.set_control_idx(idx)
    pair.orc.set_action(action)
    pair.orc.substep(0); pair.gpu.substep(0, action)
    assert_state_close(pair.gpu.get_state(1), pair.orc.get_frame(1))
    cot = rng.normal(size=(n, 24)).astype(np.float32).astype(np.float64)
    pair.orc.clear_grads(); pair.gpu.clear_all_gradients()
    pair.orc.add_frame_grad(1, cot); pair.gpu.add_state_grad(1, cot)
    pair.orc.set_action(action)
    pair.orc.substep_grad(0)
    ga = pair.gpu.substep_grad(0, action)
    assert ga.shape == action.shape
    assert rel_l2(ga, pair.orc.get_action_grad()) <= 1e-3
    assert pair.gpu.substep_grad(0) is None      # mpm_simulator.py:376-377


def test_rollout_with_resort_forward_and_backward():
    """12 substeps with a re-sort every 4: the adjoint must be carried back across the re-orderings.
    Seeds on x at three frames (as GripLoss does, loss_grip.py:117-140) and wrench seeds every substep."""
    rng = np.random.default_rng(240)
    n, steps = 6000, 12
    center = np.array([0.5, 0.3, 0.5])
    pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=steps + 2, sort_every=4, substeps=4)
    s13 = np.concatenate([center, scenes.random_quat(rng), [0.0, 0.2, 0.0], 0.5 * rng.normal(size=3)])
    pair.set_prim_state(0, 0, steps + 2, s13)
    pair.reset(scenes.contact_rollout_state(n, rng, center))
    pair.clear_ext_f()
    for f in range(steps):
        pair.substep(f)
        ref, got = pair.orc.get_frame(f + 1), pair.gpu.get_state(f + 1)
        # both trajectories start from the same frame 0 and drift apart by fp32 rounding only
        assert rel_l2(got[:, :3], ref[:, :3]) <= 1e-5
        assert rel_l2(got[:, 3:6], ref[:, 3:6]) <= 2e-3
    fo, fg = pair.orc.get_ext_f(0), pair.prims[0].get_ext_f()
    assert np.abs(fo[:3]).max() > 0 and rel_l2(fg, fo) <= 5e-3, (fg, fo)
    assert pair.gpu.counters()["resorts"] >= 3
    pair.orc.clear_grads(); pair.gpu.clear_all_gradients()
    for f in (steps, steps - 5, 3):
        g = rng.normal(size=(n, 3)).astype(np.float32).astype(np.float64)
        g24 = np.zeros((n, 24)); g24[:, :3] = g
        pair.orc.add_frame_grad(f, g24); pair.gpu.add_x_grad(f, g)
    ext = [rng.normal(size=6) * 1e-3 for _ in range(pair.P)]
    for f in range(steps - 1, -1, -1):
        for i in range(pair.P):
            pair.orc.set_ext_f_grad(i, ext[i])
        pair.orc.substep_grad(f)
        pair.gpu.substep_grad(f, ext_f_grad=ext)
    go, gg = pair.orc.get_frame_grad(0), pair.gpu.get_state_grad(0)
    c = cosine(gg, go)
    assert c >= 0.999, f"rollout adjoint cosine {c}"
    for i in range(pair.P):
        po = sum(pair.orc.get_primitive_state_grad(i, f) for f in range(steps))
        pg = pair.prims[i].get_all_states_grad(0, f_end=steps)
        assert np.abs(po).max() > 0 and cosine(pg, po) >= 0.999, (pg, po)
    assert pair.gpu.counters()["clamped"] == 0 and pair.gpu.counters()["left_active_region"] == 0


def test_sort_contract_bit_exact():
    """key = block-major linearised base cell, stable ascending sort, ties by previous order."""
    rng = np.random.default_rng(250)
    n, ng = 20000, 32
    pair = Pair(n, n_grid=ng, sort_every=2)
    st = scenes.blob_state(n, rng, width=0.4, vel=3.0)
    pair.reset(st)
    x = st[:, :3].astype(np.float32)

    def keys_of(x32):
        b = (x32 * np.float32(ng) - np.float32(0.5)).astype(np.int32)       # trunc toward zero, like .cast(int)
        b = np.clip(b, 0, ng - 3)
        nb = ng // 4
        i, j, k = b[:, 0], b[:, 1], b[:, 2]
        return ((((i >> 2) * nb + (j >> 2)) * nb + (k >> 2)) * 64 + (((i & 3) << 4) | ((j & 3) << 2) | (k & 3))).astype(np.uint32)

    perm0 = np.argsort(keys_of(x), kind="stable").astype(np.uint32)
    assert np.array_equal(pair.gpu.permutation(0), perm0)
    assert np.array_equal(pair.gpu.sort_keys(0), keys_of(x)[perm0])
    pair.gpu.substep(0); pair.gpu.substep(1)        # re-sort happens after substep 1
    x2 = pair.gpu.get_x(2).astype(np.float32)
    idx = np.argsort(keys_of(x2[perm0]), kind="stable")
    assert np.array_equal(pair.gpu.permutation(2), perm0[idx])
    k2 = pair.gpu.sort_keys(2)
    assert np.all(np.diff(k2.astype(np.int64)) >= 0)


@pytest.mark.parametrize("flags", [1, 2, 4])
def test_dense_and_unsorted_modes_agree(flags):
    """SMX_FLAG_DENSE_GRID / SMX_FLAG_NO_SORT / SMX_FLAG_DIRECT_RED change the traversal, not the arithmetic
    (up to fp32 summation order)."""
    center = np.array([0.5, 0.3, 0.5])

    def run(fl):
        rng = np.random.default_rng(260)
        pair = Pair(5000, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=8, sort_every=2, flags=fl)
        s13 = np.concatenate([center, [1, 0, 0, 0], [0.0, 0.2, 0.0], [0, 0, 0]])
        pair.prims[0].set_all_states(0, s13, f_end=8)
        pair.gpu.reset(scenes.contact_rollout_state(5000, rng, center))
        for f in range(6):
            pair.gpu.substep(f)
        return pair.gpu.get_state(6), pair.prims[0].get_ext_f()

    (a, fa), (b, fb) = run(0), run(flags)
    assert np.abs(fa).max() > 0
    assert_state_close(a, b, tol=2e-5)
    assert rel_l2(fb, fa) <= 1e-4


def test_grid_checkpoint_and_recompute_adjoints_agree():
    """The adjoint either restores the per-substep grid checkpoint (default) or re-runs P2G + grid update
    (SMX_FLAG_NO_GRID_CKPT, the reference's strategy, mpm_simulator.py:351-359): same gradients."""
    center = np.array([0.5, 0.3, 0.5])

    def run(fl):
        rng = np.random.default_rng(265)
        n, steps = 5000, 6
        pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=8, sort_every=3, flags=fl)
        s13 = np.concatenate([center, [1, 0, 0, 0], [0.0, 0.2, 0.0], [0, 0, 0.3]])
        pair.prims[0].set_all_states(0, s13, f_end=8)
        pair.gpu.reset(scenes.contact_rollout_state(n, rng, center))
        for f in range(steps):
            pair.gpu.substep(f)
        pair.gpu.clear_all_gradients()
        pair.gpu.add_x_grad(steps, rng.normal(size=(n, 3)))
        ext = [rng.normal(size=6) * 1e-3]
        for f in range(steps - 1, -1, -1):
            pair.gpu.substep_grad(f, ext_f_grad=ext)
        return pair.gpu.get_state_grad(0), pair.prims[0].get_all_states_grad(0, f_end=steps)

    (ga, pa), (gb, pb) = run(0), run(8)
    assert np.abs(ga).max() > 0 and np.abs(pa).max() > 0
    assert rel_l2(ga, gb) <= 1e-5 and rel_l2(pa, pb) <= 1e-4


def test_grid_checkpoint_overflow_invalidates_only_the_substeps_that_did_not_fit():
    """An exploding blob: the active region outgrows the checkpoint arena (sized from the first substep) halfway through the rollout.
    The substeps whose record did not fit recompute their grid in the adjoint, the earlier ones keep restoring theirs (fewer
    launches than recomputing everything), the gradient equals the all-recompute run, and after the next reset the arena is large
    enough for every substep."""
    from softmac_b200.engine import MPMSimulator
    n, steps = 6000, 16
    rng = np.random.default_rng(77)
    st = scenes.blob_state(n, rng, center=(0.5, 0.5, 0.5), width=0.08, vel=0.0, Fdev=0.0, Cdev=0.0)
    st[:, 3:6] = 40.0 * (st[:, :3] - 0.5) / 0.04                       # radial burst: +-8 cells in 16 substeps
    st = st.astype(np.float32).astype(np.float64)
    seed = rng.normal(size=(n, 3))

    def run(fl, episodes=1):
        sim = MPMSimulator(sim_cfg(n, n_grid=64, max_steps=steps + 2, gravity=(0., 0., 0.)), (), env_dt=1e-3, sort_every=2, flags=fl)
        out = []
        for _ in range(episodes):
            sim.reset(st)
            for f in range(steps):
                sim.substep(f)
            blocks = sim.counters()["active_blocks"]
            sim.clear_all_gradients(); sim.add_x_grad(steps, seed)
            l0 = sim.launch_count()
            for f in range(steps - 1, -1, -1):
                sim.substep_grad(f)
            out.append((sim.get_state_grad(0), sim.launch_count() - l0, blocks))
        return out

    (g_rec, l_rec, blocks), (g_rec2, l_rec2, _) = run(0, episodes=2)
    (g_all, l_all, _), = run(8)
    assert blocks > 170                                                  # started from <= 64 active blocks: the first arena held <= 160
    assert np.abs(g_all).max() > 0 and rel_l2(g_rec, g_all) <= 1e-5 and rel_l2(g_rec2, g_all) <= 1e-5
    assert l_rec2 < l_rec < l_all                                        # partial fallback < full recomputation; none after the arena grew


@pytest.mark.parametrize("flags", [64, 128, 64 + 128])
@pytest.mark.parametrize("ptype", [0, 1])
def test_svd_record_and_tma_staging_do_not_change_the_adjoint(flags, ptype):
    """The adjoint normally takes U, V, sigma - 1, J - 1 from the SVD record written by the forward P2G and runs the persistent
    TMA-staged P2G-adjoint kernel; SMX_FLAG_NO_SVD_REC (64) repeats the Jacobi SVD, SMX_FLAG_NO_TMA (128) reads the planes
    straight from HBM.  Same gradients (plastic state stressed enough that the sigma clip is active)."""
    def run(fl):
        rng = np.random.default_rng(266)
        n, steps = 6000, 6
        pair = Pair(n, max_steps=8, sort_every=3, flags=fl, ptype=ptype)
        st = scenes.blob_state(n, rng)
        st[:, 6:15] += 0.01 * rng.normal(size=(n, 9))
        st[:, 15:24] = rng.normal(size=(n, 9))
        pair.gpu.reset(np.asarray(st, dtype=np.float32).astype(np.float64))
        for f in range(steps):
            pair.gpu.substep(f)
        pair.gpu.clear_all_gradients()
        pair.gpu.add_state_grad(steps, rng.normal(size=(n, 24)))
        for f in range(steps - 1, -1, -1):
            pair.gpu.substep_grad(f)
        return pair.gpu.get_state(steps), pair.gpu.get_state_grad(0)

    (sa, ga), (sb, gb) = run(0), run(flags)
    assert np.abs(ga).max() > 0
    assert_state_close(sa, sb, tol=2e-5)         # forward is identical up to the order of the L2 reductions
    assert rel_l2(gb, ga) <= 2e-5 and cosine(gb, ga) >= 1 - 1e-9


def test_get_grad_is_the_xv_part_of_the_state_adjoint():
    """MPMSimulator.get_grad(f) -> (x.grad[f], v.grad[f]) (mpm_simulator.py:561-574), after a backward pass and for a frame that
    only holds its loss seed."""
    rng = np.random.default_rng(267)
    n, steps = 3000, 4
    pair = Pair(n, max_steps=8, sort_every=2)
    pair.gpu.reset(scenes.blob_state(n, rng))
    for f in range(steps):
        pair.gpu.substep(f)
    pair.gpu.clear_all_gradients()
    pair.gpu.add_state_grad(steps, rng.normal(size=(n, 24)))
    xs, vs = pair.gpu.get_grad(steps)                  # no backward step yet: just the seed
    g = pair.gpu.get_state_grad(steps)
    assert np.array_equal(xs, g[:, :3]) and np.array_equal(vs, g[:, 3:6]) and np.abs(xs).max() > 0
    for f in range(steps - 1, -1, -1):
        pair.gpu.substep_grad(f)
    xg, vg = pair.gpu.get_grad(0)
    g = pair.gpu.get_state_grad(0)
    assert np.array_equal(xg, g[:, :3]) and np.array_equal(vg, g[:, 3:6]) and np.abs(vg).max() > 0


def test_api_quirks_and_errors():
    from softmac_b200._capi import SmxError
    rng = np.random.default_rng(270)
    n = 500
    pair = Pair(n)
    st = scenes.blob_state(n, rng)
    pair.gpu.reset(st)
    got = pair.gpu.get_state(0)
    assert got.shape == (n, 24) and got.dtype == np.float64
    assert np.array_equal(got, st)                                     # fp32-representable input round-trips exactly
    x = st[:, :3] + 0.01
    pair.gpu.set_x(0, x); assert np.allclose(pair.gpu.get_x(0), x, atol=1e-7)
    pair.gpu.set_v(0, st[:, 3:6] * 2); assert np.allclose(pair.gpu.get_v(0), st[:, 3:6] * 2, atol=1e-6)
    pair.gpu.set_state(3, [st[:, :3], st[:, 3:6], st[:, 6:15].reshape(n, 3, 3), st[:, 15:].reshape(n, 3, 3)])
    assert np.array_equal(pair.gpu.get_state(3), st)
    pair.gpu.copyframe(3, 5); assert np.array_equal(pair.gpu.get_state(5), st)
    pair.gpu.reset(st[:, :3])                                          # (n,3): v = 0, F = I, C = 0
    s0 = pair.gpu.get_state(0)
    assert np.array_equal(s0[:, :3], st[:, :3]) and np.all(s0[:, 3:6] == 0) and np.all(s0[:, 15:] == 0)
    assert np.array_equal(s0[:, 6:15], np.tile(np.eye(3).ravel(), (n, 1)))
    assert pair.gpu.cur == 0
    with pytest.raises(SmxError):
        pair.gpu.substep(pair.cfg.max_steps - 1)
    with pytest.raises(SmxError):
        pair.gpu.get_state(6)               # never written
    with pytest.raises(SmxError):
        pair.gpu.substep_grad(4)            # never run forward


def test_empty_particle_set():
    pair = Pair(0)
    pair.gpu.reset(np.zeros((0, 3)))
    pair.gpu.substep(0)
    assert pair.gpu.get_state(1).shape == (0, 24)


def test_properties_at_full_size():
    """BASELINE config 3 sizes (1M particles, 128^3): size-independent properties instead of an oracle run.
    mass conservation, momentum change = gravity impulse, and a zero seed gives a zero adjoint."""
    from softmac_b200.engine import MPMSimulator
    from harness import sim_cfg
    n = 1_000_000
    cfg = sim_cfg(n, n_grid=128, max_steps=6, ground_friction=20., dt=1e-4)
    sim = MPMSimulator(cfg, env_dt=5e-4)
    st = scenes.cube_state(n)
    sim.reset(st)
    keys = sim.sort_keys(0)
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(sim.permutation(0)), np.arange(n, dtype=np.uint32))
    sim.substep(0)
    g_in, g_out = sim.get_grid()
    p_mass = (1 / 128 * 0.5) ** 2
    assert abs(g_in[:, 3].astype(np.float64).sum() / (n * p_mass) - 1) < 1e-5
    s1 = sim.get_state(1)
    assert np.allclose(s1[:, 3:6].mean(0), [0, -9.8 * 1e-4, 0], atol=1e-7)     # free fall, away from walls
    sim.substep(1); sim.substep(2); sim.substep(3)
    sim.clear_all_gradients()
    for f in range(3, -1, -1):
        sim.substep_grad(f)
    assert np.all(sim.get_state_grad(0) == 0)
    assert sim.counters()["clamped"] == 0 and sim.counters()["left_active_region"] == 0
    # the adjoint is linear in its seed: seed 2 g gives twice the gradient of seed g (up to the order of the L2 reductions)
    g = np.ascontiguousarray(st[:, :3] - st[:, :3].mean(0))
    grads = []
    for scale in (1.0, 2.0):
        sim.clear_all_gradients()
        sim.add_x_grad(4, scale * g)
        for f in range(3, -1, -1):
            sim.substep_grad(f)
        grads.append(sim.get_state_grad(0))
    assert np.abs(grads[0]).max() > 0 and np.all(np.isfinite(grads[0]))
    assert rel_l2(2.0 * grads[0], grads[1]) <= 1e-5


def test_properties_at_the_largest_baseline_size():
    """BASELINE config 5 sizes (8M particles, 256^3) in one handle: sort contract, mass conservation, free-fall momentum and the
    linearity of the adjoint, i.e. the same size-independent properties as at 1M."""
    from softmac_b200.engine import MPMSimulator
    from harness import sim_cfg
    n = 8_000_000
    cfg = sim_cfg(n, n_grid=256, max_steps=5, ground_friction=20., dt=5e-5)
    sim = MPMSimulator(cfg, env_dt=2.5e-4)
    st = scenes.cube_state(n)
    sim.reset(st)
    keys = sim.sort_keys(0)
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
    sim.step(0, 3)
    g_in, g_out = sim.get_grid()
    p_mass = (1 / 256 * 0.5) ** 2
    assert abs(g_in[:, 3].astype(np.float64).sum() / (n * p_mass) - 1) < 1e-5
    v3 = sim.get_v(3)
    assert np.allclose(v3.mean(0), [0, -9.8 * 3 * 5e-5, 0], atol=2e-7)          # three substeps of free fall
    g = np.ascontiguousarray(st[:, :3] - st[:, :3].mean(0))
    grads = []
    for scale in (1.0, -0.5):
        sim.clear_all_gradients()
        sim.add_x_grad(3, scale * g)
        sim.step_grad(3, 3)
        grads.append(sim.get_grad(0))
    assert np.abs(grads[0][0]).max() > 0 and np.all(np.isfinite(grads[0][1]))
    assert rel_l2(-0.5 * grads[0][0], grads[1][0]) <= 1e-5 and rel_l2(-0.5 * grads[0][1], grads[1][1]) <= 1e-5
    assert sim.counters()["clamped"] == 0 and sim.counters()["left_active_region"] == 0


def test_velocity_control_forward_kinematics_and_action_grad():
    """rigid_velocity_control: Primitive.set_action writes v, w of `substeps` frames, forward_kinematics integrates the pose
    inside every substep (mpm_simulator.py:329-331, primitive_base.py:280-304) and get_action_grad collects the adjoint."""
    rng = np.random.default_rng(280)
    n, substeps = 3000, 4
    center = np.array([0.5, 0.3, 0.5])
    pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=2 * substeps + 2, sort_every=3, substeps=substeps, vctrl=True)
    s13 = np.concatenate([center, scenes.random_quat(rng), np.zeros(6)])
    pair.set_prim_state(0, 0, 1, s13)
    a6 = np.array([0.4, -0.3, 0.2, 0.0, 0.15, 0.0]).astype(np.float32).astype(np.float64)       # [w(3), v(3)]
    pair.orc.set_primitive_action(0, 0, substeps, a6)
    pair.prims[0].set_action(0, substeps, a6)
    pair.reset(scenes.contact_rollout_state(n, rng, center))
    pair.clear_ext_f()
    for f in range(substeps):
        pair.substep(f)
    po, pg = pair.orc.get_primitive_state(0, substeps), pair.prims[0].get_all_states(substeps)
    assert rel_l2(pg[:7], po[:7]) <= 1e-6                                                   # integrated pose
    assert_state_close(pair.gpu.get_state(substeps), pair.orc.get_frame(substeps), tol=2e-4)
    pair.orc.clear_grads(); pair.gpu.clear_all_gradients()
    g = rng.normal(size=(n, 3)).astype(np.float32).astype(np.float64)
    g24 = np.zeros((n, 24)); g24[:, :3] = g
    pair.orc.add_frame_grad(substeps, g24); pair.gpu.add_x_grad(substeps, g)
    for f in range(substeps - 1, -1, -1):
        pair.orc.substep_grad(f); pair.gpu.substep_grad(f)
    ao, ag = pair.orc.get_primitive_action_grad(0, 0, substeps), pair.prims[0].get_action_grad(0, substeps)
    assert np.abs(ao).max() > 0 and cosine(ag, ao) >= 0.999 and rel_l2(ag, ao) <= 2e-2, (ag, ao)


def test_copy_mode_and_mid_run_set_state():
    """TaichiEnv._is_copy (taichi_env.py:106-115): run `substeps` substeps, copy the last frame to frame 0, repeat; and a
    set_state in the middle of a run re-bins the particles without disturbing the result."""
    rng = np.random.default_rng(290)
    n, substeps = 4000, 3
    pair = Pair(n, max_steps=substeps + 2, sort_every=2, substeps=substeps)
    st = scenes.blob_state(n, rng, Fdev=0.003, vel=0.3, Cdev=0.5)
    pair.reset(st)
    ref = mo_rollout(pair.orc, st, 3 * substeps)
    for rep in range(3):
        for f in range(substeps):
            pair.gpu.substep(f)
        pair.gpu.copyframe(substeps, 0)
    got = pair.gpu.get_state(0)
    assert rel_l2(got[:, :3], ref[:, :3]) <= 1e-6 and rel_l2(got[:, 3:6], ref[:, 3:6]) <= 1e-4
    # set_state of the same values at frame 0, then continue: identical to continuing directly
    a = pair.gpu.get_state(0)
    pair.gpu.substep(0)
    cont = pair.gpu.get_state(1)
    pair.gpu.set_state(0, [a[:, :3], a[:, 3:6], a[:, 6:15].reshape(n, 3, 3), a[:, 15:].reshape(n, 3, 3)])
    pair.gpu.substep(0)
    assert rel_l2(pair.gpu.get_state(1), cont) <= 1e-6


def mo_rollout(orc, st, steps):
    """oracle rollout that recycles two frames (the oracle keeps only max_steps frames)."""
    orc.set_frame(0, st)
    for f in range(steps):
        orc.substep(0)
        orc.set_frame(0, orc.get_frame(1))
    return orc.get_frame(0)


def test_fused_step_matches_substep_loop():
    """smx_step fuses the G2P of substep f into the P2G of substep f+1 (G2P2G); frames, wrench and adjoints must equal the
    plain substep loop (SMX_FLAG_NO_FUSION) -- up to the fp32 summation order of the scatter."""
    center = np.array([0.5, 0.3, 0.5])
    steps = 9

    def run(flags, use_step):
        rng = np.random.default_rng(300)
        pair = Pair(5000, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=steps + 2, sort_every=4, flags=flags, n_control=1)
        s13 = np.concatenate([center, [1, 0, 0, 0], [0.0, 0.2, 0.0], [0, 0, 0.3]])
        pair.prims[0].set_all_states(0, s13, f_end=steps + 2)
        pair.gpu.reset(scenes.contact_rollout_state(5000, rng, center))
        pair.gpu.set_control_idx(np.zeros(5000, dtype=np.int32))
        pair.gpu.set_action(np.array([[3.0, -2.0, 1.0]]))
        pair.prims[0].clear_ext_f()
        if use_step:
            pair.gpu.step(0, steps)
        else:
            for f in range(steps):
                pair.gpu.substep(f)
        frames = [pair.gpu.get_state(f) for f in (1, 4, 5, steps)]
        fe = pair.prims[0].get_ext_f()
        pair.gpu.clear_all_gradients()
        pair.gpu.add_x_grad(steps, rng.normal(size=(5000, 3)))
        pair.gpu.add_state_grad(6, 0.3 * rng.normal(size=(5000, 24)))   # a seed inside the call: that boundary is not fused
        pair.prims[0].set_ext_f_grad(1e-3 * rng.normal(size=6))
        l0 = pair.gpu.launch_count()
        if use_step:
            pair.gpu.step_grad(steps, steps)        # P2G adjoint (f) + G2P adjoint (f-1) share a launch inside an ordering
        else:
            for f in range(steps - 1, -1, -1):
                pair.gpu.substep_grad(f)
        nl = pair.gpu.launch_count() - l0
        return frames, fe, pair.gpu.get_state_grad(0), pair.prims[0].get_all_states_grad(0, f_end=steps), pair.gpu.get_action_grad(), nl

    (fa, ea, ga, pa, aa, la), (fb, eb, gb, pb, ab, lb) = run(0, True), run(32, False)
    for a, b in zip(fa, fb):
        assert_state_close(a, b, tol=2e-5)
    assert np.abs(ea).max() > 0 and rel_l2(ea, eb) <= 1e-4
    assert rel_l2(ga, gb) <= 1e-4
    assert np.abs(pa).max() > 0 and rel_l2(pa, pb) <= 1e-3 and np.abs(aa).max() > 0 and rel_l2(aa, ab) <= 1e-4
    assert la <= lb - 4, (la, lb)                   # at least four of the eight boundaries were fused


def test_repeated_backward_passes_on_one_forward_are_identical():
    """Nothing a backward pass leaves behind may leak into the next one: the grids restored from the per-substep records, the adjoint grids
    re-zeroed block by block (gg_mix only where the contact adjoint scattered: the block flags of k_contact_grad) and the flags themselves.
    Two fused passes (smx_step_grad) and one substep-wise pass over the same forward rollout, with forecast contact and re-sorts."""
    center, steps, n = np.array([0.5, 0.3, 0.5]), 10, 5000
    rng = np.random.default_rng(310)
    pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=steps + 2, sort_every=4)
    pair.prims[0].set_all_states(0, np.concatenate([center, [1, 0, 0, 0], [0.0, 0.2, 0.0], [0, 0, 0.3]]), f_end=steps + 2)
    pair.gpu.reset(scenes.contact_rollout_state(n, rng, center))
    pair.prims[0].clear_ext_f()
    pair.gpu.step(0, steps)
    seed, ext = rng.normal(size=(n, 3)), 1e-3 * rng.normal(size=6)

    def backward(fused):
        pair.gpu.clear_all_gradients()
        pair.gpu.add_x_grad(steps, seed)
        pair.prims[0].set_ext_f_grad(ext)
        if fused:
            pair.gpu.step_grad(steps, steps)
        else:
            for f in range(steps - 1, -1, -1):
                pair.gpu.substep_grad(f)
        return pair.gpu.get_state_grad(0), pair.prims[0].get_all_states_grad(0, f_end=steps)

    g1, p1 = backward(True)
    g2, p2 = backward(True)
    g3, p3 = backward(False)
    g4, p4 = backward(True)
    assert np.abs(g1).max() > 0 and np.abs(p1).max() > 0
    # the scatter's L2 reductions arrive in any order: equal up to fp32 summation order
    assert rel_l2(g2, g1) <= 1e-5 and rel_l2(p2, p1) <= 1e-4
    assert rel_l2(g3, g1) <= 1e-4 and rel_l2(p3, p1) <= 1e-3
    assert rel_l2(g4, g1) <= 1e-5 and rel_l2(p4, p1) <= 1e-4


def test_copy_mode_resorts_by_age_not_by_frame_index():
    """TaichiEnv.set_copy(True) (taichi_env.py:106-115): step `substeps` substeps from frame 0, copyframe(cur, 0), cur = 0 -- frame
    indices never reach sort_every = max(substeps, 4) when substeps < 4 (demo_pour has 1).  The re-sort is triggered by the number
    of substeps since the last binning, so a blob flying several cells stays inside its active-block list and equals the oracle."""
    rng = np.random.default_rng(910)
    n, substeps, env_steps = 3000, 2, 40
    pair = Pair(n, max_steps=substeps + 2, substeps=substeps, gravity=(0., 0., 0.), n_grid=32)      # default sort_every = 4
    st = scenes.blob_state(n, rng, center=(0.3, 0.5, 0.5), width=0.10, vel=0.05, Fdev=0.002, Cdev=0.2)
    st[:, 3] += 8.0                                     # 8 m/s along x: 80 substeps * 2e-4 * 8 = 0.128 = 4.1 cells of the 32^3 grid
    pair.reset(st)
    for k in range(env_steps):
        for f in range(substeps):
            pair.substep(f)
        last = pair.orc.get_frame(substeps)
        pair.orc.set_frame(0, last)                     # the oracle has no orderings: copyframe is a plain copy
        pair.gpu.copyframe(substeps, 0)
    c = pair.gpu.counters()
    assert c["resorts"] >= env_steps * substeps // 4 - 1, c
    assert c["left_active_region"] == 0 and c["clamped"] == 0, c
    ref, got = pair.orc.get_frame(0), pair.gpu.get_state(0)
    assert np.abs(ref[:, 0] - st[:, 0]).min() > 0.1    # the blob did move four cells
    assert rel_l2(got[:, :3], ref[:, :3]) <= 1e-5
    assert rel_l2(got[:, 3:6], ref[:, 3:6]) <= 1e-3
    assert rel_l2(got[:, 6:15], ref[:, 6:15]) <= 1e-4


def test_primitive_reset_clears_adjoints_and_action_buffer():
    """Primitive.reset() (primitive_base.py:236-246, 267-275, 321-326) zeroes the state series, their adjoints, the action buffer and
    the wrench; a Primitive built with the default max_timesteps (2048) on a simulator with fewer frames must not raise."""
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    t = scenes.sphere_table()
    m = Mesh(sdf=dict(sdf=t["sdf"], normal=t["normal"], position=(t["lower"], t["upper"]), dx=t["dx"]), cfg=dict(friction=0.5), rigid_velocity_control=True)
    prims = Primitives(primitives=[m], rigid_velocity_control=True)
    sim = MPMSimulator(sim_cfg(64, max_steps=16), prims, env_dt=1e-3, rigid_velocity_control=True)
    m.set_all_states(0, np.arange(1.0, 14.0), f_end=16)
    m.add_all_states_grad(3, np.ones(13))
    m.set_action(1, 5, np.arange(6.0))
    assert np.abs(m.get_all_states_grad(0, f_end=16)).sum() > 0
    m.reset()                                           # default max_timesteps 2048 > max_steps 16
    assert np.abs(m.get_all_states(5)).max() == 0 and np.abs(m.get_all_states_grad(0, f_end=16)).max() == 0
    assert np.abs(m.get_action_grad(1, 5)).max() == 0 and np.abs(m.get_ext_f()).max() == 0
    del sim


def test_fp32_host_and_device_entry_points_equal_the_f64_calls():
    """reset / get_state / add_x_grad / add_state_grad / get_grad / get_state_grad with float32 host arrays (smx_*_f32, buffers pinned with
    MPMSimulator.pin) and with device arrays (__cuda_array_interface__: torch CUDA tensors -> smx_*_dev) give exactly what the float64
    calls give: the float64 calls convert to the same fp32 rows."""
    import torch
    rng = np.random.default_rng(77)
    n, S = 3000, 4
    center = np.array([0.5, 0.3, 0.5])
    st = scenes.contact_rollout_state(n, rng, center, speed=1.0)        # fp32-representable; rests just outside the sphere, moving into it
    seed3 = rng.normal(size=(n, 3)).astype(np.float32)
    seed24 = rng.normal(size=(n, 24)).astype(np.float32)

    def run(mode):
        pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=S + 2, sort_every=2)
        sim = pair.gpu
        pair.prims[0].set_all_states(0, np.concatenate([center, [1, 0, 0, 0], [0, 0.3, 0], [0, 0, 0]]), f_end=S + 2)
        if mode == "f64":
            sim.reset(st)
        elif mode == "f32":
            st32 = st.astype(np.float32)
            sim.pin(st32)
            sim.reset(st32)
            sim.unpin(st32)
        else:
            sim.reset(torch.from_numpy(st.astype(np.float32)).cuda())
        sim.step(0, S)
        sim.clear_all_gradients()
        f0 = sim.get_state(0)
        assert np.array_equal(f0, st), mode                 # the uploaded rows are exactly the caller's
        if mode == "f64":
            sim.add_x_grad(S, seed3.astype(np.float64)); sim.add_state_grad(S - 1, seed24.astype(np.float64))
            sim.step_grad(S, S)
            return sim.get_state(S), sim.get_grad(0), sim.get_state_grad(0)
        if mode == "f32":
            sim.add_x_grad(S, seed3); sim.add_state_grad(S - 1, seed24)
            sim.step_grad(S, S)
            xg, vg = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
            sim.pin(xg, vg)
            out = sim.get_state(S, dtype=np.float32), sim.get_grad(0, out=(xg, vg)), sim.get_state_grad(0, dtype=np.float32)
            sim.unpin(xg, vg)
            return out
        sim.add_x_grad(S, seed3); sim.add_state_grad(S - 1, torch.from_numpy(seed24).cuda())
        sim.step_grad(S, S)
        a, b = torch.empty((n, 24), device="cuda"), torch.empty((n, 24), device="cuda")
        sim.get_state(S, out=a); sim.get_state_grad(0, out=b)
        sim.synchronize()
        return a.cpu().numpy(), sim.get_grad(0, dtype=np.float32), b.cpu().numpy()

    ref_state, (ref_xg, ref_vg), ref_adj = run("f64")
    assert np.abs(ref_xg).max() > 0
    for mode in ("f32", "dev"):
        state, (xg, vg), adj = run(mode)
        # the same kernels on the same inputs: runs differ only by the order of the float reductions in the scatters
        assert state.dtype == np.float32 and rel_l2(state, ref_state) <= 1e-6, (mode, rel_l2(state, ref_state))
        assert xg.dtype == np.float32 and rel_l2(xg, ref_xg) <= 1e-4 and rel_l2(vg, ref_vg) <= 1e-4, mode
        assert rel_l2(adj, ref_adj) <= 1e-4, (mode, rel_l2(adj, ref_adj))
        assert np.array_equal(adj[:, :3].astype(np.float32), xg) and np.array_equal(adj[:, 3:6].astype(np.float32), vg), mode
    with pytest.raises(TypeError):
        Pair(10).gpu.reset(torch.zeros((10, 24), device="cuda", dtype=torch.float64))
