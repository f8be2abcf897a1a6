"""demo_grip-shaped scene built the way the reference builds it (softmac/engine/taichi_env.py:32-59): primitives from a
URDF (one Mesh per collision mesh, SDF tables built from the OBJ -- here on the GPU), the MPM simulator, a rigid
simulator (stand-in for Jade) configured from the same URDF, a Chamfer loss, and the TaichiEnv loop; contact is
disabled on the palm exactly as demo_grip.py:117 does.  A few optimisation-style iterations must run and produce
finite, non-trivial action gradients."""
import os

import numpy as np
import pytest

from harness import sim_cfg

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_grip_like_episode_from_urdf(tmp_path):
    from softmac_b200.config import CfgNode
    from softmac_b200.engine import MPMSimulator, Primitives
    from softmac_b200.engine.taichi_env import TaichiEnv
    from softmac_b200.engine.rigid_simulator import RigidSimulator, bodies_from_urdf
    from softmac_b200.engine.losses import ChamferLoss
    import scenes
    urdf = scenes.write_demo_assets(str(tmp_path / "assets"))["gripper"]      # palm + two fingers from tests/golden/demo_meshes.npz
    G = np.load(os.path.join(GOLDEN, "grip_palm_contact.npz"))
    x0 = G["state0"].astype(np.float64)                      # 2500 particles of the reference's grip initial state
    n, env_steps, substeps, dt = len(x0), 6, 5, 2e-4
    max_steps = env_steps * substeps + substeps + 2
    gripper = CfgNode(friction=0.001, urdf_path=urdf, enable_external_force=True)
    prims = Primitives([gripper], max_timesteps=max_steps, cache_dir=str(tmp_path))
    assert len(prims) == 3 and prims[1].sdf_res == prims[2].sdf_res          # palm + two fingers (same mesh)
    sim = MPMSimulator(sim_cfg(n, n_grid=64, max_steps=max_steps, dt=dt), prims, env_dt=dt * substeps)
    sim.primitives_contact = [False, True, True]                               # demo_grip.py:117
    bodies = bodies_from_urdf(urdf)
    assert [b["joint"] for b in bodies] == ["fixed", "prismatic", "prismatic"]
    assert np.allclose(bodies[1]["origin"], (0.35, 0.2, 0.5)) and np.allclose(bodies[2]["origin"], (0.65, 0.2, 0.5))
    # start the fingers closer to the plasticine than the URDF rest pose so that 6 env steps reach it
    rcfg = CfgNode(gravity=(0., 0., 0.), init_state=(0.02, -0.02, 3.0, -3.0), bodies=bodies)
    rigid = RigidSimulator(rcfg, prims, substeps=substeps, env_dt=dt * substeps)
    target = x0[:, :3] * np.array([0.7, 1.2, 1.0]) + np.array([0.15, 0.0, 0.0])
    env = TaichiEnv(sim, prims, rigid, x0, loss=ChamferLoss(sim, target), control_mode="rigid")
    actions = np.tile(0.3 * np.array([1.0, -1.0]) * 100, (env_steps, 1))
    losses = []
    for it in range(2):
        env.reset()
        sim.clear_all_gradients()
        for a in actions:
            env.step(a)
        info = env.compute_loss(env_steps * substeps)
        grad = env.backward()
        assert grad.shape == (env_steps, 2) and np.isfinite(grad).all()
        losses.append(info["loss"])
        actions = actions - 1e3 * grad
    fe = [p.get_ext_f() for p in prims]
    assert np.abs(grad).max() > 0                      # the fingers touched the plasticine: the loss depends on the actions
    assert sim.counters()["clamped"] == 0
