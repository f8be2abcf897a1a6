"""CPU checks behind tools/bench_demo.py and the device-resident rigid coupling:
  * the demo fixtures (reference initial states, Chamfer targets, URDF / OBJ assets copied as data) load and have the shapes the demo
    configs state (softmac/config/demo_grip_config.py, demo_pour_config.py);
  * the stand-in integrator is exactly affine in (state, action, wrench) for fixed, prismatic and free joints -- the property
    smx_rigid_linear_* relies on (softmac_b200/csrc/smx_rigid.cuh) -- and the matrices DeviceLinearRigid would upload reproduce it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_demo_scenes_load():
    import bench_demo as bd
    from softmac_b200.engine.primitive.primitives import Primitives
    from softmac_b200.engine.primitive.sdf_builder import load_obj
    grip, pour = bd.scene("grip"), bd.scene("pour")
    assert grip["state"].shape == (10000, 24) and grip["target"].shape == (10000, 3) and grip["substeps"] == 5
    assert pour["state"].shape == (5000, 24) and pour["target"].shape == (5000, 3) and pour["substeps"] == 1
    assert abs(pour["state"][:, 1].min() - (0.19998057 + 0.04)) < 1e-6            # SHAPES offset (0, 0.04, 0), demo_pour_config.py:34
    assert len(pour["rigid_init"]) == 24 and len(grip["rigid_init"]) == 4
    n_meshes = []
    for sc in (grip, pour):
        for c in sc["prims"]:
            paths, _ = Primitives.load_info_from_urdf(c["urdf_path"])
            n_meshes.append(len(paths))
            for p in paths:
                V, F = load_obj(p)
                assert len(V) > 0 and len(F) > 0 and F.max() < len(V)
    assert n_meshes == [3, 1, 1]                                                   # palm + two fingers, glass, bowl
    joints = [[b["joint"] for b in bd.rigid_bodies(sc)] for sc in (grip, pour)]
    assert joints == [["fixed", "prismatic", "prismatic"], ["free", "free"]]
    a = bd.actions_for("pour", 3000)
    assert a.shape == (3000, 12) and np.allclose(a[0, 3:6], [0, 0.9, 0]) and np.allclose(a[700, :3], [0, 0, 0.05]) and not a[:, 6:].any()
    assert np.allclose(bd.actions_for("grip", 400)[7], [0.3, -0.3])


class _Prim:
    enable_external_force = True

    def set_all_states(self, *a, **k):
        pass


def test_standin_integrator_is_affine_for_every_joint_type():
    from softmac_b200.engine.rigid_simulator import RigidSimulator
    from softmac_b200.config import CfgNode
    bodies = [dict(joint="free", origin=(0.4, 0.3, 0.5), quat=(0.9238795, 0, 0.3826834, 0), mass=1.3, inertia=0.02, gravity=True),
              dict(joint="prismatic", axis=(0.6, 0.8, 0.0), origin=(0.6, 0.3, 0.5), mass=2.0, gravity=True),
              dict(joint="fixed", origin=(0.5, 0.5, 0.5))]
    r = RigidSimulator(CfgNode(gravity=(0, -9.8, 0), init_state=(), bodies=bodies), [_Prim(), _Prim(), _Prim()], substeps=5, env_dt=1e-3)
    sd, ad, nw = r.state_dim, r.action_dim, 18
    assert (sd, ad) == (14, 7)
    s0, a0, w0 = np.zeros(sd), np.zeros(ad), np.zeros(nw)
    c = r._advance(s0, a0, w0)                      # what DeviceLinearRigid.__init__ computes (eps = 1: exact for an affine map)
    As = r._jac(lambda x: r._advance(x, a0, w0), s0, eps=1.0).T
    Aa = r._jac(lambda x: r._advance(s0, x, w0), a0, eps=1.0).T
    Aw = r._jac(lambda x: r._advance(s0, a0, x), w0, eps=1.0).T
    rng = np.random.default_rng(0)
    for _ in range(5):
        s, a, w = rng.normal(size=sd), rng.normal(size=ad), rng.normal(size=nw)
        assert np.abs(r._advance(s, a, w) - (s @ As + a @ Aa + w @ Aw + c)).max() <= 1e-14
    assert np.abs(Aw[12:]).max() == 0               # the fixed body ignores its wrench
    assert c[7 + 4] < 0                             # gravity enters through the constant term (free body, y velocity)
