"""Builds the same scene in the f64 oracle and in the CUDA simulator (through the C ABI) for parity tests."""
import numpy as np

import scenes


def sim_cfg(n, n_grid=32, max_steps=8, dt=2e-4, E=3e3, nu=0.2, gravity=(0., -9.8, 0.), ground_friction=20.,
            material_model=0, ptype=0, collision_type=2, n_control=0):
    from softmac_b200.config import CfgNode
    return CfgNode(dim=3, quality=n_grid / 64.0, yield_stress=30., dtype="float64", max_steps=max_steps, n_particles=n,
                   E=E, nu=nu, ground_friction=ground_friction, gravity=tuple(gravity), ptype=ptype,
                   material_model=material_model, dt=dt, n_controllers=n_control, collision_type=collision_type)


class Pair:
    """oracle + cuda simulators with identical parameters, primitives and (fp32-rounded) inputs."""

    def __init__(self, n, tables=(), prim_params=(), substeps=5, sort_every=None, flags=0, vctrl=False, yield_stress=None, **kw):
        from oracle import mpm_oracle as mo
        from softmac_b200.engine import MPMSimulator, Primitives, Mesh
        cfg = sim_cfg(n, **kw)
        self.cfg, self.n, self.substeps = cfg, n, substeps
        n_grid = int(128 * cfg.quality * 0.5)
        self.orc = mo.OracleSim(n, n_grid=n_grid, max_steps=cfg.max_steps, dt=cfg.dt, E=cfg.E, nu=cfg.nu, gravity=cfg.gravity,
                                ground_friction=cfg.ground_friction, material_model=cfg.material_model, ptype=cfg.ptype,
                                collision_type=cfg.collision_type, substeps=substeps, n_control=cfg.n_controllers,
                                rigid_velocity_control=vctrl)
        prims = []
        for tab, (fric, soft) in zip(tables, prim_params):
            t32 = {k: (np.asarray(v, dtype=np.float32).astype(np.float64) if k in ("sdf", "normal", "lower", "upper") else v)
                   for k, v in tab.items()}
            self.orc.add_primitive(t32["sdf"], t32["normal"], t32["lower"], t32["upper"], tab["dx"], friction=fric, softness=soft)
            m = Mesh(sdf=dict(sdf=t32["sdf"], normal=t32["normal"], position=(t32["lower"], t32["upper"]), dx=tab["dx"]),
                     cfg=dict(friction=fric), max_timesteps=cfg.max_steps, rigid_velocity_control=vctrl)
            m.softness[None] = soft
            prims.append(m)
        self.prims = Primitives(primitives=prims, max_timesteps=cfg.max_steps, rigid_velocity_control=vctrl)
        self.gpu = MPMSimulator(cfg, self.prims, env_dt=cfg.dt * substeps, rigid_velocity_control=vctrl, sort_every=sort_every, flags=flags)
        self.P = len(prims)
        if yield_stress is not None:        # soft_cloth's flow rule (soft_cloth/engine/mpm_simulator.py:232)
            self.orc.set_plasticity(1, yield_stress)
            self.gpu.set_plasticity("von_mises", yield_stress)

    def set_prim_state(self, i, f0, f1, s13):
        s13 = np.asarray(s13, dtype=np.float32).astype(np.float64)
        for f in range(f0, f1):
            self.orc.set_primitive_state(i, f, s13)
        self.prims[i].set_all_states(f0, s13, f_end=f1)

    def reset(self, st24):
        st24 = np.asarray(st24, dtype=np.float32).astype(np.float64)
        self.orc.set_frame(0, st24)
        self.gpu.reset(st24)

    def clear_ext_f(self):
        for i in range(self.P):
            self.orc.clear_ext_f(i)
            self.prims[i].clear_ext_f()

    def substep(self, f):
        self.orc.substep(f)
        self.gpu.substep(f)


def rel_l2(a, b, floor=0.0):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), floor, 1e-300)


def cosine(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def prim_states_for(rng, P, center):
    out = []
    for i in range(P):
        pos = np.asarray(center) + np.array([0.09 * (1 - 2 * i), -0.03, 0.02 * i])
        out.append(np.concatenate([pos, scenes.random_quat(rng) * 1.07, 0.3 * rng.normal(size=3), 2.0 * rng.normal(size=3)]))
    return out
