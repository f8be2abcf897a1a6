"""Golden fixture tests/golden/grip_palm_contact.npz (inputs: the reference's grip initial state and palm SDF
pickle; outputs: f64 oracle -- regression pins, see tests/golden/make_fixtures.py).
CPU: the oracle still reproduces the committed vectors.  GPU: the CUDA path matches them through the C ABI."""
import os

import numpy as np
import pytest

from harness import rel_l2, cosine

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grip_palm_contact.npz"))


def table():
    return dict(sdf=G["sdf"].astype(np.float64), normal=G["normal"].astype(np.float64), lower=G["lower"].astype(np.float64),
                upper=G["upper"].astype(np.float64), dx=float(G["sdf_dx"]))


def test_oracle_reproduces_golden():
    from oracle import mpm_oracle as mo
    st, steps, t = G["state0"].astype(np.float64), int(G["steps"]), table()
    sim = mo.OracleSim(len(st), n_grid=64, max_steps=steps + 1, dt=2e-4, E=3e3, nu=0.2, gravity=(0., -9.8, 0.), ground_friction=20.,
                       material_model=0, ptype=0, collision_type=2, substeps=5)
    sim.add_primitive(t["sdf"], t["normal"], t["lower"], t["upper"], t["dx"], friction=0.001, softness=666.)
    for f in range(steps + 1):
        sim.set_primitive_state(0, f, G["prim_state"])
    sim.set_frame(0, st)
    for f in range(steps):
        sim.substep(f)
    assert rel_l2(sim.get_frame(1), G["state_1"]) < 1e-12
    assert rel_l2(sim.get_frame(steps), G["state_final"]) < 1e-11
    assert rel_l2(sim.get_ext_f(0), G["ext_f"]) < 1e-9
    g24 = np.zeros_like(st); g24[:, :3] = G["seed_x"]
    sim.add_frame_grad(steps, g24)
    for f in range(steps - 1, -1, -1):
        sim.set_ext_f_grad(0, G["ext_seed"])
        sim.substep_grad(f)
    assert rel_l2(sim.get_frame_grad(0), G["adj0"]) < 1e-9
    pg = np.stack([sim.get_primitive_state_grad(0, f) for f in range(steps)])
    assert rel_l2(pg, G["prim_grad"]) < 1e-8


@pytest.mark.gpu
def test_cuda_matches_golden():
    from harness import Pair
    st, steps, t = G["state0"].astype(np.float64), int(G["steps"]), table()
    pair = Pair(len(st), tables=[t], prim_params=[(0.001, 666.)], n_grid=64, max_steps=steps + 1, substeps=5, sort_every=5)
    pair.prims[0].set_all_states(0, G["prim_state"], f_end=steps + 1)
    pair.gpu.reset(st)
    pair.prims[0].clear_ext_f()
    pair.gpu.substep(0)
    s1 = pair.gpu.get_state(1)
    for k, sl in dict(x=slice(0, 3), v=slice(3, 6), F=slice(6, 15), C=slice(15, 24)).items():
        e = rel_l2(s1[:, sl], G["state_1"][:, sl])
        assert e <= 1e-4, f"substep 0, {k}: rel L2 {e:.3e}"
    for f in range(1, steps):
        pair.gpu.substep(f)
    sf = pair.gpu.get_state(steps)
    assert rel_l2(sf[:, :3], G["state_final"][:, :3]) <= 1e-6
    assert rel_l2(sf[:, 3:6], G["state_final"][:, 3:6]) <= 1e-3
    assert rel_l2(sf[:, 6:15], G["state_final"][:, 6:15]) <= 1e-5
    assert rel_l2(pair.prims[0].get_ext_f(), G["ext_f"]) <= 2e-3
    pair.gpu.clear_all_gradients()
    pair.gpu.add_x_grad(steps, G["seed_x"])
    for f in range(steps - 1, -1, -1):
        pair.gpu.substep_grad(f, ext_f_grad=[G["ext_seed"]])
    adj0 = pair.gpu.get_state_grad(0)
    assert cosine(adj0, G["adj0"]) >= 0.9999 and rel_l2(adj0, G["adj0"]) <= 5e-3
    pg = np.stack([pair.prims[0].get_all_states_grad(f) for f in range(steps)])
    assert cosine(pg, G["prim_grad"]) >= 0.999, (pg, G["prim_grad"])
