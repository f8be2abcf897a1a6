"""Adapters that give the f64 oracle the MPMSimulator / Primitive surface, so the SAME env + rigid stand-in + loss code
can be driven by the oracle and by the CUDA simulator in the coupling tests.  Test infrastructure only."""
import numpy as np
from oracle import mpm_oracle as mo


class _ExtF:
    def __init__(self, p):
        self.p = p

    def to_numpy(self):
        return self.p.sim.get_ext_f(self.p.i)


class OraclePrimitive:
    def __init__(self, sim, i, enable_external_force=True):
        self.sim, self.i, self.enable_external_force = sim, i, enable_external_force
        self.ext_f = _ExtF(self)

    def set_all_states(self, f, state, f_end=None):
        for ff in range(f, (f + 1) if f_end is None else f_end):
            self.sim.set_primitive_state(self.i, ff, state)

    def get_all_states_grad(self, f):
        return self.sim.get_primitive_state_grad(self.i, f)

    def get_all_states(self, f):
        return self.sim.get_primitive_state(self.i, f)

    def add_all_states_grad(self, f, g13):
        self.sim.add_primitive_state_grad(self.i, f, np.asarray(g13, dtype=np.float64))

    def clear_ext_f(self):
        self.sim.clear_ext_f(self.i)

    def set_ext_f_grad(self, g):
        self.sim.set_ext_f_grad(self.i, g)

    def reset(self):
        self.clear_ext_f()


class OraclePrimitives(list):
    def initialize(self):
        pass

    def reset(self):
        for p in self:
            p.reset()


class OracleMPMSimulator:
    def __init__(self, n, n_grid, max_steps, dt, substeps, tables=(), prim_params=(), **kw):
        self.sim = mo.OracleSim(n, n_grid=n_grid, max_steps=max_steps, dt=dt, substeps=substeps, **kw)
        self.n_particles, self.substeps, self.cur, self.dim = n, substeps, 0, 3
        self.n_control, self.dtype = int(kw.get("n_control", 0)), np.float64
        self.primitives = OraclePrimitives()
        for i, (t, (fr, so)) in enumerate(zip(tables, prim_params)):
            self.sim.add_primitive(t["sdf"], t["normal"], t["lower"], t["upper"], t["dx"], friction=fr, softness=so)
            self.primitives.append(OraclePrimitive(self.sim, i))
        self.n_primitive = len(self.primitives)

    def initialize(self):
        pass

    def reset(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.shape[1] == 3:
            st = np.zeros((len(x), 24)); st[:, :3] = x; st[:, 6] = st[:, 10] = st[:, 14] = 1
            x = st
        self.sim.set_frame(0, x)
        self.sim.clear_grads()
        self.cur = 0

    def set_control_idx(self, idx):
        self.sim.set_control_idx(np.asarray(idx))

    def substep(self, s, action=None):
        if action is not None:                                  # mpm_simulator.py:321-322
            self.sim.set_action(np.asarray(action, dtype=np.float64).reshape(self.n_control, 3))
        self.sim.substep(s)

    def substep_grad(self, s, action=None, ext_f_grad=None):
        if action is not None:                                  # set_action zeroes action.grad (mpm_simulator.py:579-586)
            self.sim.set_action(np.asarray(action, dtype=np.float64).reshape(self.n_control, 3))
        if ext_f_grad is not None:
            for i, g in enumerate(ext_f_grad):
                self.sim.set_ext_f_grad(i, g)
        self.sim.substep_grad(s)
        if action is None:
            return None
        return self.sim.get_action_grad().reshape(np.shape(action))

    def get_x(self, f):
        return self.sim.get_frame(f)[:, :3]

    def get_state(self, f):
        return self.sim.get_frame(f)

    def add_x_grad(self, f, g):
        g24 = np.zeros((self.n_particles, 24)); g24[:, :3] = g
        self.sim.add_frame_grad(f, g24)

    def copyframe(self, a, b):
        self.sim.set_frame(b, self.sim.get_frame(a))
