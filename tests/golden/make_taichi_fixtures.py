"""Pins the oracle against the REAL reference: runs the UNMODIFIED softmac/engine/mpm_simulator.py + primitive classes under Taichi
(ti.cpu, f64) on the inputs of tests/golden/grip_palm_contact.npz and writes tests/golden/taichi_grip_palm_contact.npz with the same
output keys.  tests/test_taichi_golden.py then holds the f64 oracle to 1e-9 and the CUDA path to north_star's tolerances against
TAICHI output; while that file is absent those tests skip and parity stays "unpinned" (DESIGN.md section 3).

Needs a machine with the reference's own stack (not installable in the build image: no wheels, no network):
    pip install taichi==1.4.1 trimesh yacs            # requirements.txt of the reference
    PYTHONPATH=/path/to/SoftMAC python tests/golden/make_taichi_fixtures.py

Nothing here copies reference code: it imports it.  The scene is the one make_fixtures.py builds for the oracle: 2500 grip particles,
the gripper-palm SDF table pressed into the blob, demo_grip material, 5 substeps (life = 1/5 .. 1), then the adjoint of a dense seed on
x[5] and of a wrench seed (softmac/engine/mpm_simulator.py:320-378, primitive_base.py:139-181).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import taichi as ti
    import trimesh
    from yacs.config import CfgNode as CN
    ti.init(arch=ti.cpu, default_fp=ti.f64, debug=False)
    from softmac.engine.mpm_simulator import MPMSimulator
    from softmac.engine.primitive.mesh import Mesh

    G = np.load(os.path.join(HERE, "grip_palm_contact.npz"))
    st, steps = G["state0"].astype(np.float64), int(G["steps"])
    n = len(st)
    table = {"sdf": G["sdf"].astype(np.float64), "normal": G["normal"].astype(np.float64),
             "position": (G["lower"].astype(np.float64), G["upper"].astype(np.float64)), "dx": np.ones(3) * float(G["sdf_dx"]),
             "res": np.array(G["sdf"].shape)}
    # Mesh.__init__ (mesh.py:18-43) builds its Taichi tables from preprocess_sdf(mesh_path): hand it the committed table instead of a
    # mesh file (the table IS the reference's own cached pickle of the palm, stored as fp32)
    Mesh.preprocess_sdf = lambda self, mesh_path: (table, [trimesh.creation.box()])
    prim_cfg = CN(); prim_cfg.friction = 0.001; prim_cfg.urdf_path = ""; prim_cfg.enable_external_force = True
    mesh = Mesh("palm", cfg=prim_cfg, max_timesteps=steps + 1, dtype=ti.f64)
    cfg = CN()
    cfg.dim, cfg.quality, cfg.yield_stress, cfg.dtype, cfg.max_steps, cfg.n_particles = 3, 1.0, 30., "float64", steps + 1, n
    cfg.E, cfg.nu, cfg.ground_friction, cfg.gravity, cfg.ptype, cfg.material_model = 3e3, 0.2, 20., (0., -9.8, 0.), 0, 0
    cfg.dt, cfg.n_controllers, cfg.collision_type = 2e-4, 0, 2
    sim = MPMSimulator(cfg, [mesh], env_dt=1e-3)                 # substeps = 5
    sim.initialize()
    mesh.initialize()
    mesh.softness[None] = 666.                                    # Primitives.set_softness (primitives.py:55-56)
    s13 = G["prim_state"].astype(np.float64)
    for f in range(steps + 1):
        mesh.set_all_states(f, s13)
    sim.set_state(0, [st[:, 0:3].copy(), st[:, 3:6].copy(), st[:, 6:15].reshape(n, 3, 3).copy(), st[:, 15:24].reshape(n, 3, 3).copy()])
    mesh.clear_ext_f()
    frames = []
    for f in range(steps):
        sim.substep(f)
        frames.append(sim.get_state(f + 1))
    ext_f = mesh.ext_f.to_numpy().reshape(6)
    # adjoint: x.grad[steps] = seed_x, wrench adjoint = ext_seed, then substep_grad(steps - 1 .. 0)
    ti.ad.clear_all_gradients()
    xg = np.zeros((steps + 1, n, 3)); xg[steps] = G["seed_x"]
    sim.x.grad.from_numpy(xg)
    for f in range(steps - 1, -1, -1):
        sim.substep_grad(f, ext_f_grad=[G["ext_seed"].astype(np.float64)])
    adj0 = np.hstack([sim.x.grad.to_numpy()[0], sim.v.grad.to_numpy()[0], sim.F.grad.to_numpy()[0].reshape(n, 9), sim.C.grad.to_numpy()[0].reshape(n, 9)])
    pgrad = np.stack([mesh.get_all_states_grad(f) for f in range(steps)])
    out = os.path.join(HERE, "taichi_grip_palm_contact.npz")
    np.savez_compressed(out, state_1=frames[0], state_final=frames[-1], ext_f=ext_f, adj0=adj0, prim_grad=pgrad,
                        taichi_version=np.array(ti.__version__), generator=np.array("tests/golden/make_taichi_fixtures.py"))
    print("wrote", out, "| ext_f", ext_f, "| |adj0|", np.linalg.norm(adj0))


if __name__ == "__main__":
    sys.exit(main())
