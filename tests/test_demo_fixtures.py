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
              dict(joint="fixed", origin=(0.5, 0.5, 0.5)),
              dict(joint="revolute", axis=(0.0, 0.6, 0.8), origin=(0.25, 0.0, 0.3), quat=(0.9659258, 0.2588190, 0, 0), inertia=7.8e-6)]
    r = RigidSimulator(CfgNode(gravity=(0, -9.8, 0), init_state=(), bodies=bodies), [_Prim(), _Prim(), _Prim(), _Prim()], substeps=5, env_dt=1e-3)
    sd, ad, nw = r.state_dim, r.action_dim, 24
    assert (sd, ad) == (16, 8)
    s0, a0, w0 = np.zeros(sd), np.zeros(ad), np.zeros(nw)
    c = r._advance(s0, a0, w0)                      # what DeviceLinearRigid.__init__ computes (eps = 1: exact for an affine map)
    As = r._jac(lambda x: r._advance(x, a0, w0), s0, eps=1.0).T
    Aa = r._jac(lambda x: r._advance(s0, x, w0), a0, eps=1.0).T
    Aw = r._jac(lambda x: r._advance(s0, a0, x), w0, eps=1.0).T
    rng = np.random.default_rng(0)
    for _ in range(5):
        s, a, w = rng.normal(size=sd), rng.normal(size=ad), rng.normal(size=nw)
        ref = r._advance(s, a, w)
        assert (np.abs(ref - (s @ As + a @ Aa + w @ Aw + c)) / np.maximum(1.0, np.abs(ref))).max() <= 1e-14
    assert np.abs(Aw[12:18]).max() == 0             # the fixed body ignores its wrench
    assert c[8 + 4] < 0                             # gravity enters through the constant term (free body, y velocity)
    assert np.abs(Aw[18:21]).max() == 0 and np.abs(Aw[21:24, 8 + 7]).max() > 0      # a hinge only feels the torque about its axis
    # closed-form pose Jacobians (fixed / prismatic / revolute) against central differences of the pose map
    for i in (1, 2, 3):
        for _ in range(3):
            s = rng.normal(size=sd)
            Ja, Jn = r._pose_jac(s, i), r._jac(lambda x: r._pose(x, i), s)
            assert np.abs(Ja - Jn).max() <= 1e-8, (i, np.abs(Ja - Jn).max())
    # the hinged body turns about its (body-frame) axis: R = R0 Rot(axis, theta), unit quaternion, twist (0, axis * omega)
    s = np.zeros(sd); s[7] = 0.7; s[8 + 7] = -2.0
    p = r._pose(s, 3)
    assert np.allclose(p[:3], (0.25, 0.0, 0.3)) and np.isclose(np.linalg.norm(p[3:7]), 1.0)
    assert np.allclose(p[7:10], 0) and np.allclose(p[10:13], -2.0 * np.array([0.0, 0.6, 0.8]))


def test_door_urdf_maps_to_a_revolute_body(tmp_path):
    """A door-like URDF (one revolute joint about y, inertia tensor on the child link; structure of assets/door/door.urdf) ->
    one revolute Body with the moment of inertia about the hinge axis."""
    from softmac_b200.engine.rigid_simulator import bodies_from_urdf, Body
    urdf = """<?xml version="1.0" ?><robot name="d"><link name="world"/>
      <joint name="j" type="revolute"><parent link="world"/><child link="leaf"/><origin xyz="0.25 0.0 0.3" rpy="0 0 0"/><axis xyz="0 1 0"/>
        <limit lower="-3.14" upper="3.14" effort="0" velocity="6.5"/></joint>
      <link name="leaf"><inertial><mass value="0.0125"/><inertia ixx="2.8e-06" ixy="0" ixz="0" iyy="7.9e-06" iyz="0" izz="1.1e-05"/></inertial>
        <collision><geometry><mesh filename="leaf.obj"/></geometry></collision></link></robot>"""
    path = tmp_path / "d.urdf"
    path.write_text(urdf)
    specs = bodies_from_urdf(str(path))
    assert len(specs) == 1 and specs[0]["joint"] == "revolute" and specs[0]["axis"] == (0.0, 1.0, 0.0)
    assert np.allclose(specs[0]["origin"], (0.25, 0.0, 0.3)) and np.isclose(specs[0]["inertia"], 7.9e-06)
    assert Body(**specs[0]).ndof == 1
