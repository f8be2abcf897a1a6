"""The oracle pins itself: every hand-written adjoint in oracle/mpm_oracle.c is checked against central
finite differences of the oracle's own forward in float64 (the reference has
no tests, so this is the only available pin for the adjoint)."""
import numpy as np
import pytest
from oracle import mpm_oracle as mo
import scenes


def make_sim(rng, n=160, n_grid=32, collision_type=2, ptype=0, material_model=0, n_prim=1, n_control=0,
             gravity=(0., -9.8, 0.), ground_friction=20., substeps=5, center=(0.5, 0.3, 0.5), vctrl=False,
             frame=0, max_steps=4):
    sim = mo.OracleSim(n, n_grid=n_grid, max_steps=max_steps, dt=2e-4, E=3e3, nu=0.2, gravity=gravity,
                       ground_friction=ground_friction, material_model=material_model, ptype=ptype,
                       collision_type=collision_type, substeps=substeps, n_control=n_control,
                       rigid_velocity_control=vctrl)
    tab = scenes.sphere_table()
    prim_states = []
    for i in range(n_prim):
        sim.add_primitive(tab["sdf"], tab["normal"], tab["lower"], tab["upper"], tab["dx"],
                          friction=0.4 + 0.3 * i, softness=666., enabled=True)
        pos = np.asarray(center) + np.array([0.09 * (1 - 2 * i), -0.03, 0.02 * i])
        s13 = np.concatenate([pos, scenes.random_quat(rng) * 1.07, 0.3 * rng.normal(size=3), 2.0 * rng.normal(size=3)])
        prim_states.append(s13)
    st = scenes.blob_state(n, rng, center=center, fp32=False)
    return sim, st, prim_states


def run_forward(sim, st, prim_states, action, f=0):
    for i, s13 in enumerate(prim_states):
        sim.clear_ext_f(i)
        sim.set_primitive_state(i, f, s13)
    sim.set_frame(f, st)
    if action is not None:
        sim.set_action(action)
    sim.substep(f)
    out = [sim.get_frame(f + 1).ravel()]
    for i in range(len(prim_states)):
        out.append(sim.get_ext_f(i))
        if sim_vctrl(sim):
            out.append(sim.get_primitive_state(i, f + 1)[:7])
    return np.concatenate(out)


def sim_vctrl(sim):
    return getattr(sim, "_vctrl", False)


def check_vjp(rng, sim, st, prim_states, action=None, f=0, eps=1e-7, rtol=2e-5, n_dirs=3):
    n, P = sim.n, len(prim_states)
    y0 = run_forward(sim, st, prim_states, action, f)
    cot = rng.normal(size=y0.shape)
    # adjoint
    sim.clear_grads()
    off = n * 24
    sim.add_frame_grad(f + 1, cot[:off].reshape(n, 24))
    for i in range(P):
        sim.set_ext_f_grad(i, cot[off:off + 6]); off += 6
        if sim_vctrl(sim):
            g13 = np.zeros(13); g13[:7] = cot[off:off + 7]; off += 7
            sim.add_primitive_state_grad(i, f + 1, g13)
    if action is not None:
        sim.set_action(action)
    sim.substep_grad(f)
    g_st = sim.get_frame_grad(f)
    g_pr = [sim.get_primitive_state_grad(i, f) for i in range(P)]
    g_act = sim.get_action_grad() if action is not None else None
    for _ in range(n_dirs):
        d_st = rng.normal(size=st.shape)
        d_pr = [rng.normal(size=13) for _ in range(P)]
        d_act = rng.normal(size=action.shape) if action is not None else None
        yp = run_forward(sim, st + eps * d_st, [s + eps * d for s, d in zip(prim_states, d_pr)],
                         None if action is None else action + eps * d_act, f)
        ym = run_forward(sim, st - eps * d_st, [s - eps * d for s, d in zip(prim_states, d_pr)],
                         None if action is None else action - eps * d_act, f)
        fd = cot @ (yp - ym) / (2 * eps)
        an = (g_st * d_st).sum() + sum((g * d).sum() for g, d in zip(g_pr, d_pr))
        if action is not None:
            an += (g_act * d_act).sum()
        assert abs(fd - an) <= rtol * max(abs(fd), abs(an), 1e-3), (fd, an)
    return g_st, g_pr


@pytest.mark.parametrize("ptype", [0, 1, 2])
def test_vjp_mixed_contact_corotated(ptype):
    rng = np.random.default_rng(10 + ptype)
    sim, st, prs = make_sim(rng, ptype=ptype, n_prim=2)
    g_st, g_pr = check_vjp(rng, sim, st, prs)
    assert np.abs(g_pr[0]).max() > 0  # contact was active: primitive received gradient


def test_vjp_von_mises_return_mapping():
    """soft_cloth's plastic material (soft_cloth/engine/mpm_simulator.py:172-189, :232): the adjoint through the log-strain return mapping,
    with yielding and non-yielding particles in the same blob, and the forward against a direct numpy evaluation of the formula."""
    rng = np.random.default_rng(77)
    sim, st, prs = make_sim(rng, ptype=0, n_prim=1)
    mu = 3e3 / (2 * 1.2)
    F0, C0 = st[:, 6:15].reshape(-1, 3, 3), st[:, 15:24].reshape(-1, 3, 3)
    Ftmp = (np.eye(3)[None] + 2e-4 * C0) @ F0
    U, sg, Vt = np.linalg.svd(Ftmp)
    eps = np.log(np.maximum(sg, 0.05))
    eh = eps - eps.mean(1, keepdims=True)
    nrm = np.sqrt((eh * eh).sum(1) + 1e-8)
    yield_stress = 2 * mu * float(np.median(nrm))                 # half of the blob yields
    sim.set_plasticity(1, yield_stress)
    check_vjp(rng, sim, st, prs)
    # forward: F[f+1] of every particle from numpy's SVD of F_tmp (U diag(.) V^T is invariant to the SVD's sign / ordering freedom)
    run_forward(sim, st, prs, None)
    F1 = sim.get_frame(1)[:, 6:15].reshape(-1, 3, 3)
    dg = nrm - yield_stress / (2 * mu)
    yields = dg > 0
    assert 0.3 < yields.mean() < 0.7
    Fy = U @ (np.exp(eps - (dg / nrm)[:, None] * eh)[:, :, None] * Vt)
    want = np.where(yields[:, None, None], Fy, Ftmp)
    assert np.abs(F1 - want).max() <= 1e-12


def test_contact_is_exercised():
    rng = np.random.default_rng(3)
    sim, st, prs = make_sim(rng, n_prim=2)
    run_forward(sim, st, prs, None)
    assert np.abs(sim.get_ext_f(0)).max() > 0 and np.abs(sim.get_ext_f(1)).max() > 0


@pytest.mark.parametrize("ptype", [1, 2])
def test_vjp_neohookean(ptype):
    rng = np.random.default_rng(20 + ptype)
    sim, st, prs = make_sim(rng, ptype=ptype, material_model=1, n_prim=1)
    check_vjp(rng, sim, st, prs)


def test_vjp_grid_contact():
    rng = np.random.default_rng(30)
    sim, st, prs = make_sim(rng, collision_type=0, n_prim=2)
    g_st, g_pr = check_vjp(rng, sim, st, prs)
    assert np.abs(g_pr[0]).max() > 0


def test_vjp_particle_contact():
    rng = np.random.default_rng(31)
    sim, st, prs = make_sim(rng, collision_type=1, n_prim=2)
    g_st, g_pr = check_vjp(rng, sim, st, prs)
    assert np.abs(g_pr[0]).max() > 0


def test_vjp_control_action_and_walls():
    rng = np.random.default_rng(32)
    # cloud pushed into the floor / wall corner so the boundary condition masks are exercised
    sim, st, prs = make_sim(rng, n_prim=0, n_control=2, center=(0.09, 0.09, 0.5), ground_friction=0.)
    sim.set_control_idx(rng.integers(-1, 2, size=sim.n))
    action = rng.normal(size=(2, 3)) * 50
    check_vjp(rng, sim, st, prs, action=action)
    sim2, st2, prs2 = make_sim(rng, n_prim=0, center=(0.5, 0.08, 0.5), ground_friction=20.)
    check_vjp(rng, sim2, st2, prs2)


def test_vjp_later_frame_life():
    rng = np.random.default_rng(33)
    sim, st, prs = make_sim(rng, n_prim=1, max_steps=6)
    check_vjp(rng, sim, st, prs, f=3)   # life = 1/(5 - 3) = 0.5


def test_vjp_velocity_control_fk():
    rng = np.random.default_rng(34)
    sim, st, prs = make_sim(rng, n_prim=1, vctrl=True)
    sim._vctrl = True
    check_vjp(rng, sim, st, prs)


def test_svd_grad_formula_matches_fd():
    """mpm_simulator.py:140-157 against finite differences of U S^a V^T style functions."""
    rng = np.random.default_rng(5)
    F = np.eye(3) + 0.3 * rng.normal(size=(3, 3))
    gu, gs, gv = rng.normal(size=(3, 3)), np.diag(rng.normal(size=3)), rng.normal(size=(3, 3))

    def phi(F):
        U, S, V = mo.svd3(F)
        # a basis-invariant scalar of (U, S, V): sum(G1 * U S^2 V^T) + sum(G2 * U V^T)
        return (gu * (U @ S @ S @ V.T)).sum() + (gv * (U @ V.T)).sum()

    U, S, V = mo.svd3(F)
    # adjoints of U, S, V for phi
    GU = gu @ V @ S @ S + gv @ V
    GV = gu.T @ U @ S @ S + gv.T @ U
    GS = np.diag(np.diag(U.T @ gu @ V) * 2 * np.diag(S))
    an = mo.backward_svd(GU, GS, GV, U, S, V)
    eps = 1e-6
    fd = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            E = np.zeros((3, 3)); E[i, j] = eps
            fd[i, j] = (phi(F + E) - phi(F - E)) / (2 * eps)
    assert np.allclose(an, fd, rtol=1e-6, atol=1e-8)


def test_mass_and_momentum_conservation():
    rng = np.random.default_rng(6)
    sim, st, prs = make_sim(rng, n_prim=0, gravity=(0., 0., 0.), ground_friction=0.)
    st[:, 15:] = 0  # C = 0 so that grid momentum equals particle momentum exactly
    st[:, 6:15] = np.eye(3).ravel()
    run_forward(sim, st, prs, None)
    gvin, gm, gvout = sim.get_grid()
    p_mass = (1.0 / 32 * 0.5) ** 2
    assert np.isclose(gm.sum(), sim.n * p_mass, rtol=1e-12)
    assert np.allclose(gvin.sum(0), p_mass * st[:, 3:6].sum(0), rtol=1e-9, atol=1e-14)
