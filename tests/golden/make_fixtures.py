"""Generates tests/golden/*.npz.  Run HERE (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_fixtures.py

Inputs come from the reference's own data fixtures:
  softmac/envs/grip/grip_mpm_init_state.npy            (10000, 24) f64  -- initial plasticine state of demo_grip
  softmac/assets/gripper/68956732...                   cached SDF pickle of the gripper palm (mesh.py:148-163)
Outputs ("golden vectors") come from the f64 oracle, because the reference itself (Taichi 1.4.1) cannot be imported
in this image -- they are REGRESSION pins of the restatement, not outputs of the reference (parity unpinned).

grip_palm_contact.npz : 2500 grip particles (fp32-rounded), the palm SDF (fp32) pressed into the top of the blob,
                        5 substeps (substeps = 5 => life = 1/5 .. 1), demo_grip material (plastic, corotated, E 3e3,
                        nu 0.2, gravity -9.8, sticky floor, mixed contact, softness 666, friction 0.001), then the adjoint
                        of a dense seed x_bar[5] = x[5] - mean and a wrench seed.
reference_rest_states.npz : the reference's own simulated states, copied as data (fp32): the full demo_grip initial state
                        (10000 x 24: plasticine that the REFERENCE simulator let settle on the sticky floor, residual rms velocity
                        1.3e-3 m/s, all singular values of F inside the plastic clip [0.998, 1.003]) and the demo_pour initial state
                        (5000 x 24 liquid at rest in the glass).  These are the only outputs of the reference simulator that ship
                        with the repository; tests/test_reference_states.py uses them as a (weak) pin: a state the reference
                        produced as a rest state must be a rest state of the oracle and of the CUDA path under the demo's
                        material parameters, and must stop being one when the parameters are wrong.
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/softmac"


def reference_states():
    grip = np.load(os.path.join(REF, "envs/grip/grip_mpm_init_state.npy")).astype(np.float32)
    pour = np.load(os.path.join(REF, "envs/pour/pour_mpm_init_state_corotated.npy")).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_rest_states.npz"), grip=grip, pour=pour)
    print("reference_rest_states.npz", grip.shape, pour.shape)


def main():
    from oracle import mpm_oracle as mo
    reference_states()
    st = np.load(os.path.join(REF, "envs/grip/grip_mpm_init_state.npy"))
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(len(st), 2500, replace=False))
    st = st[sel].astype(np.float32).astype(np.float64)
    with open(os.path.join(REF, "assets/gripper/68956732a79bf09d8703ab990a2e2319bf5492c792294e9a86632db03b5ac4d5"), "rb") as f:
        blob = pickle.load(f)["sdf"]
    sdf = blob["sdf"].astype(np.float32)
    nrm = blob["normal"].astype(np.float32)
    lower, upper = [np.asarray(p, dtype=np.float32) for p in blob["position"]]
    dx = float(blob["dx"][0])
    # palm box is 0.6 x 0.3 x 0.16 (half extents 0.3, 0.15, 0.08); rotate it so that its thin axis (z) points up (+y) and press
    # its bottom face 4 mm into the top of the plasticine blob, moving down at 0.3 m/s with a small spin
    ymax = st[:, 1].max()
    c, s_ = np.cos(-np.pi / 4), np.sin(-np.pi / 4)          # rotation by -90 deg about x: (w, x, y, z)
    quat = np.array([c, s_, 0.0, 0.0])
    pos = np.array([0.5, ymax + 0.08 - 0.004, 0.5])
    s13 = np.concatenate([pos, quat, [0.0, -0.3, 0.0], [0.0, 0.0, 0.5]]).astype(np.float32).astype(np.float64)
    steps = 5
    sim = mo.OracleSim(len(st), n_grid=64, max_steps=steps + 1, dt=2e-4, E=3e3, nu=0.2, gravity=(0., -9.8, 0.), ground_friction=20.,
                       material_model=0, ptype=0, collision_type=2, substeps=5)
    sim.add_primitive(sdf.astype(np.float64), nrm.astype(np.float64), lower.astype(np.float64), upper.astype(np.float64), dx,
                      friction=0.001, softness=666.)
    for f in range(steps + 1):
        sim.set_primitive_state(0, f, s13)
    sim.set_frame(0, st)
    frames = [st]
    for f in range(steps):
        sim.substep(f)
        frames.append(sim.get_frame(f + 1))
    ext_f = sim.get_ext_f(0)
    assert np.abs(ext_f).max() > 0, "fixture must exercise contact"
    seed = frames[-1][:, :3] - frames[-1][:, :3].mean(0)
    g24 = np.zeros_like(st); g24[:, :3] = seed
    ext_seed = np.array([1e-3, -2e-3, 5e-4, 1e-4, 2e-4, -1e-4])
    sim.clear_grads()
    sim.add_frame_grad(steps, g24)
    for f in range(steps - 1, -1, -1):
        sim.set_ext_f_grad(0, ext_seed)
        sim.substep_grad(f)
    adj0 = sim.get_frame_grad(0)
    pgrad = np.stack([sim.get_primitive_state_grad(0, f) for f in range(steps)])
    out = os.path.join(HERE, "grip_palm_contact.npz")
    np.savez_compressed(out, state0=st.astype(np.float32), sdf=sdf, normal=nrm, lower=lower, upper=upper, sdf_dx=np.float64(dx),
                        prim_state=s13, steps=np.int32(steps), state_final=frames[-1], state_1=frames[1], ext_f=ext_f,
                        seed_x=seed, ext_seed=ext_seed, adj0=adj0, prim_grad=pgrad, particle_index=sel.astype(np.int32))
    print("wrote", out, os.path.getsize(out) / 1e6, "MB; ext_f", ext_f, "| |adj0|", np.linalg.norm(adj0), "| prim grad", np.abs(pgrad).max())


if __name__ == "__main__":
    main()


def reference_sdf_tables():
    """tests/golden/mesh_sdf_reference_tables.npz: the two SDF caches the REFERENCE ships (output of its own
    Mesh.trimesh2sdf with trimesh 3.21.5): gripper palm and door -- mesh (vertices, faces) + tables.  sdf stored as fp32,
    normals as int8 (both meshes are axis-aligned boxes: every stored normal is +-e_k / (1 + 1e-8))."""
    out = {}
    for name, rel in (("palm", "assets/gripper/68956732a79bf09d8703ab990a2e2319bf5492c792294e9a86632db03b5ac4d5"),
                      ("door", "assets/door/e7ab3378b317f8d1d4de18fa5bfa4d98e79629e714104b720ebcf0470dfc561a")):
        with open(os.path.join(REF, rel), "rb") as f:
            b = pickle.load(f)
        V, Fc = [np.array(a) for a in b["meshes"][0]]
        s = b["sdf"]
        assert np.allclose(np.abs(np.round(s["normal"])), np.abs(s["normal"]) * (1 + 1e-8))
        out[name + "_V"], out[name + "_F"] = V, Fc.astype(np.int32)
        out[name + "_sdf"], out[name + "_normal"] = s["sdf"].astype(np.float32), np.round(s["normal"]).astype(np.int8)
        out[name + "_lower"], out[name + "_upper"] = s["position"]
        out[name + "_dx"], out[name + "_res"] = s["dx"][0], np.array(s["res"])
    np.savez_compressed(os.path.join(HERE, "mesh_sdf_reference_tables.npz"), **out)


if __name__ == "__main__" and "--sdf" in sys.argv:
    reference_sdf_tables()


def demo_targets():
    """tests/golden/demo_targets.npz: the Chamfer target point sets of demo_grip / demo_pour (cfg.ENV.loss.target_path,
    demo_grip_config.py / demo_pour_config.py), copied as data in fp32 -- inputs of tools/bench_demo.py (BASELINE configs 1, 2)."""
    grip = np.load(os.path.join(REF, "envs/grip/grip_mpm_target_position.npy")).astype(np.float32)
    pour = np.load(os.path.join(REF, "envs/pour/pour_mpm_target_position_corotated.npy")).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "demo_targets.npz"), grip=grip, pour=pour)
    print("demo_targets.npz", grip.shape, pour.shape)


if __name__ == "__main__" and "--targets" in sys.argv:
    demo_targets()


def demo_meshes():
    """tests/golden/demo_meshes.npz: the collision meshes and rigid-body parameters of the demo_grip / demo_pour assets as ARRAYS
    (vertices f64, triangles int32; joint type / origin / axis / mass per body read from the URDFs) -- the inputs
    tests/scenes.py:write_demo_assets turns back into OBJ + URDF files in a scratch directory for the URDF-driven builders."""
    import xml.etree.ElementTree as ET
    from oracle.sdf_builder import load_obj
    out = {}
    for name, urdf in (("gripper", "assets/gripper/gripper.urdf"), ("glass", "assets/glass/glass.urdf"), ("bowl", "assets/bowl/bowl.urdf")):
        root = ET.parse(os.path.join(REF, urdf)).getroot()
        joints = {j.find("child").attrib["link"]: j for j in root.findall("joint")}
        links, meshes = [], {}
        for link in root.findall("link"):
            m = link.find("collision/geometry/mesh")
            if m is None:
                continue
            fn = m.attrib["filename"]
            if fn not in meshes:
                V, Fc = load_obj(os.path.join(REF, os.path.dirname(urdf), fn))
                meshes[fn] = len(meshes)
                out[f"{name}_mesh{meshes[fn]}_V"], out[f"{name}_mesh{meshes[fn]}_F"] = V, Fc.astype(np.int32)
            j = joints[link.attrib["name"]]
            org = [float(v) for v in j.find("origin").attrib.get("xyz", "0 0 0").split()]
            ax = [float(v) for v in j.find("axis").attrib["xyz"].split()] if j.find("axis") is not None else [1.0, 0.0, 0.0]
            jt = {"fixed": 0, "prismatic": 1, "floating": 2}[j.attrib["type"]]
            parent = j.find("parent").attrib["link"]
            mass = float(link.find("inertial/mass").attrib["value"])
            rgba = [float(v) for v in link.find("visual/material/color").attrib["rgba"].split()]
            links.append([jt, meshes[fn], 0 if parent == "world" else 1] + org + ax + [mass] + rgba)
        out[f"{name}_links"] = np.array(links)      # rows: joint type, mesh index, parent (0 world / 1 first link), origin xyz, axis xyz, mass, rgba
    np.savez_compressed(os.path.join(HERE, "demo_meshes.npz"), **out)
    print("demo_meshes.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__" and "--meshes" in sys.argv:
    demo_meshes()
