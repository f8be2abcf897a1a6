"""Two independent restatements of the hot path must agree: oracle/mpm_oracle.c (C, hand-written adjoints) against
oracle/mpm_torch_oracle.py (PyTorch float64, transliterated separately from the reference sources, adjoints from torch.autograd with
Taichi's sub-gradient conventions made explicit) -- forward AND adjoint, for every code-path row of SURVEY.md 8a's coverage table.
Neither has been compared with Taichi output (not installable here): parity to the reference stays unpinned, but a misreading of the
reference now has to be made twice, in two languages and two derivations of the adjoint, to go unnoticed."""
import numpy as np
import pytest

import scenes
from oracle import mpm_oracle as mo
from oracle.mpm_torch_oracle import TorchOracle

# (name, kwargs): the rows of SURVEY.md 8a "code-path coverage by config"
ROWS = [
    ("grip: corotated plastic, forecast contact, sticky floor", dict(ptype=0, material_model=0, collision_type=2, n_prim=2)),
    ("pour: corotated liquid, forecast contact, free-slip floor, substeps 1", dict(ptype=2, material_model=0, collision_type=2, n_prim=2, ground_friction=0., substeps=1)),
    ("door: corotated elastic, forecast contact, particle-force control, no gravity", dict(ptype=1, material_model=0, collision_type=2, n_prim=1, n_control=2, gravity=(0., 0., 0.), ground_friction=0.)),
    ("pour_vel: liquid, particle (penalty) contact, velocity control", dict(ptype=2, material_model=0, collision_type=1, n_prim=1, vctrl=True, ground_friction=0.)),
    ("neo-Hookean elastic", dict(ptype=1, material_model=1, collision_type=2, n_prim=1)),
    ("neo-Hookean liquid", dict(ptype=2, material_model=1, collision_type=2, n_prim=1)),
    ("grid contact (collision_type 0)", dict(ptype=0, material_model=0, collision_type=0, n_prim=2)),
    ("blob on the sticky floor and against a wall, later frame (life = 1/2)", dict(ptype=0, material_model=0, collision_type=2, n_prim=1, center=(0.085, 0.085, 0.5), frame=3, max_steps=6)),
    ("soft_cloth: corotated plastic with the von Mises return mapping (yield 50: about half of the particles yield)", dict(ptype=0, material_model=0, collision_type=2, n_prim=1, yield_stress=50.)),
    ("free-slip walls and ceiling", dict(ptype=1, material_model=0, collision_type=2, n_prim=0, center=(0.88, 0.88, 0.88), ground_friction=0., gravity=(3., 9.8, 2.))),
]


def build(rng, n=500, n_grid=32, collision_type=2, ptype=0, material_model=0, n_prim=1, n_control=0, gravity=(0., -9.8, 0.),
          ground_friction=20., substeps=5, center=(0.5, 0.3, 0.5), vctrl=False, frame=0, max_steps=4, yield_stress=None):
    kw = dict(n_grid=n_grid, dt=2e-4, E=3e3, nu=0.2, gravity=gravity, ground_friction=ground_friction, material_model=material_model,
              ptype=ptype, collision_type=collision_type, substeps=substeps, n_control=n_control, rigid_velocity_control=vctrl)
    c = mo.OracleSim(n, max_steps=max_steps, **kw)
    t = TorchOracle(**kw)
    if yield_stress is not None:                    # soft_cloth/engine/mpm_simulator.py:232
        c.set_plasticity(1, yield_stress); t.set_plasticity(1, yield_stress)
    tab = scenes.sphere_table()
    s13s = []
    for i in range(n_prim):
        fr = 0.4 + 0.3 * i
        c.add_primitive(tab["sdf"], tab["normal"], tab["lower"], tab["upper"], tab["dx"], friction=fr, softness=666., enabled=True)
        t.add_primitive(tab["sdf"], tab["normal"], tab["lower"], tab["upper"], tab["dx"], fr, 666.)
        pos = np.asarray(center) + np.array([0.09 * (1 - 2 * i), -0.03, 0.02 * i])
        s13s.append(np.concatenate([pos, scenes.random_quat(rng) * 1.07, 0.3 * rng.normal(size=3), 2.0 * rng.normal(size=3)]))
    st = scenes.blob_state(n, rng, center=center, fp32=False)
    if ptype == 2 and material_model == 0:          # liquid states are multiples of I after the first substep; start from a generic F anyway
        pass
    return c, t, st, s13s


@pytest.mark.parametrize("name,kw", ROWS, ids=[r[0].split(":")[0] for r in ROWS])
def test_c_oracle_equals_torch_oracle_forward_and_adjoint(name, kw):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    kw = dict(kw)
    f = kw.get("frame", 0)
    c, t, st, s13s = build(rng, **kw)
    n, P = len(st), len(s13s)
    nc = kw.get("n_control", 0)
    action = rng.normal(size=(nc, 3)) * 50.0 if nc else None
    ctrl = rng.integers(-1, nc, size=n).astype(np.int32) if nc else None
    vctrl = kw.get("vctrl", False)
    # ---- C oracle: forward -----------------------------------------------------------------------------------------
    for i, s in enumerate(s13s):
        c.clear_ext_f(i); c.set_primitive_state(i, f, s)
    c.set_frame(f, st)
    if nc:
        c.set_control_idx(ctrl); c.set_action(action)
    c.substep(f)
    out_c = c.get_frame(f + 1)
    wr_c = [c.get_ext_f(i) for i in range(P)]
    # ---- cotangents ---------------------------------------------------------------------------------------------------
    cot = rng.normal(size=(n, 24))
    gext = [rng.normal(size=6) * 1e-2 for _ in range(P)]
    gpose = [np.concatenate([rng.normal(size=7), np.zeros(6)]) for _ in range(P)] if vctrl else None
    # ---- torch oracle: forward + autograd -------------------------------------------------------------------------------
    out_t, wr_t, nxt_t, g_st_t, g_pr_t, g_act_t = t.substep_with_vjp(f, st, s13s, cot, gext, action, ctrl, gpose)
    scale = lambda a: max(np.abs(a).max(), 1e-30)
    for k, sl in dict(x=slice(0, 3), v=slice(3, 6), F=slice(6, 15), C=slice(15, 24)).items():
        assert np.abs(out_c[:, sl] - out_t[:, sl]).max() <= 1e-10 * scale(out_t[:, sl]), (name, k)
    for i in range(P):
        if kw.get("collision_type", 2) != 2 or True:
            assert np.abs(wr_c[i] - wr_t[i]).max() <= 1e-9 * max(scale(wr_t[i]), 1e-6), (name, "wrench", wr_c[i], wr_t[i])
    if P and kw.get("center") is None:
        assert max(np.abs(w).max() for w in wr_t) > 0, "the scene must exercise contact"
    if vctrl:
        for i in range(P):
            assert np.abs(c.get_primitive_state(i, f + 1)[:7] - nxt_t[i]).max() <= 1e-12
    # ---- C oracle: adjoint ------------------------------------------------------------------------------------------------
    c.clear_grads()
    c.add_frame_grad(f + 1, cot)
    for i in range(P):
        c.set_ext_f_grad(i, gext[i])
        if vctrl:
            c.add_primitive_state_grad(i, f + 1, gpose[i])
    if nc:
        c.set_action(action)
    c.substep_grad(f)
    g_st_c = c.get_frame_grad(f)
    assert np.abs(g_st_t).max() > 0
    assert np.abs(g_st_c - g_st_t).max() <= 1e-8 * scale(g_st_t), (name, "state adjoint", np.abs(g_st_c - g_st_t).max(), scale(g_st_t))
    for i in range(P):
        g_c = c.get_primitive_state_grad(i, f)
        assert np.abs(g_c - g_pr_t[i]).max() <= 1e-8 * max(scale(g_pr_t[i]), 1e-9), (name, "primitive adjoint", g_c, g_pr_t[i])
    if nc:
        g_c = c.get_action_grad()
        assert np.abs(g_act_t).max() > 0 and np.abs(g_c - g_act_t).max() <= 1e-9 * scale(g_act_t), (name, "action adjoint")
