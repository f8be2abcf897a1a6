"""``TaichiEnv`` -- the orchestration loop of ``softmac/engine/taichi_env.py`` (step / step_grad / backward / reset /
compute_loss / set_copy) around injected components.

The reference builds everything from a yacs cfg (URDF + trimesh primitives, Jade, pyrender); those builders are out
of scope, so the components are passed in: any simulator with the ``MPMSimulator``
surface, a ``Primitives`` container, a rigid simulator with the ``RigidSimulator`` surface and an optional loss with
``compute_loss(f)`` / ``seed(f)``.  The control flow -- which is what couples the MPM hot path to the rigid
simulator -- is the reference's (taichi_env.py:93-151); the only change is that the substeps of an env step go down in one
native call (``simulator.step`` / ``step_grad``) when the simulator offers it and no per-substep MPM action is involved.
"""
import numpy as np


class TaichiEnv:
    def __init__(self, simulator, primitives, rigid_simulator, init_particles, loss=None, control_mode="rigid",
                 rigid_velocity_control=False):
        assert control_mode in ("mpm", "rigid")
        self.control_mode = control_mode
        self.rigid_velocity_control = rigid_velocity_control
        self.simulator, self.primitives, self.rigid_simulator, self.loss = simulator, primitives, rigid_simulator, loss
        self.init_particles = np.asarray(init_particles, dtype=np.float64)
        self.n_particles = len(self.init_particles)
        self.substeps = simulator.substeps
        self.use_loss = loss is not None
        self._is_copy = False
        self.action_list = []
        self.initialize()

    def set_copy(self, is_copy: bool):
        self._is_copy = is_copy

    def initialize(self):
        self.primitives.initialize()
        self.simulator.initialize()
        self.rigid_simulator.initialize()
        if self.loss:
            self.loss.initialize()
        self.reset()

    def reset(self):
        self.primitives.reset()
        self.simulator.reset(self.init_particles)
        self.rigid_simulator.reset()
        if self.loss:
            self.loss.reset()
        self.action_list = []

    def step(self, action=None):                                         # taichi_env.py:93-115
        start = 0 if self._is_copy else self.simulator.cur
        self.simulator.cur = start + self.substeps
        mpm_action = action if self.control_mode == "mpm" else None
        rigid_action = action if self.control_mode == "rigid" else None
        self.action_list.append(action)
        if mpm_action is None and hasattr(self.simulator, "step"):
            # the same substeps in ONE native call (smx_step: identical results, the G2P of a substep fused into the next P2G)
            self.simulator.step(start, self.substeps)
        else:
            for s in range(start, self.simulator.cur):
                self.simulator.substep(s, mpm_action)
        self.rigid_simulator.step(start // self.substeps, rigid_action)
        if self._is_copy:
            self.simulator.copyframe(self.simulator.cur, 0)
            self.simulator.cur = 0
            if self.rigid_simulator.n_primitive > 0 and not self.rigid_velocity_control:
                self.rigid_simulator.states = [self.rigid_simulator.states[-1], ]
                self.rigid_simulator.jacob_ds_df = []
                self.rigid_simulator.jacob_ds_ds = []
                self.rigid_simulator.jacob_ds_da = []
                self.rigid_simulator.jacob_external = []

    def step_grad(self, action=None):                                    # taichi_env.py:117-137
        start = self.simulator.cur
        self.simulator.cur = start - self.substeps
        mpm_action = action if self.control_mode == "mpm" else None
        rigid_action = action if self.control_mode == "rigid" else None
        rigid_action_grad, ext_f_grad_list = self.rigid_simulator.step_grad(self.simulator.cur // self.substeps, rigid_action)
        mpm_action_grad = np.zeros(np.shape(action)) if action is not None else None
        if mpm_action is None and hasattr(self.simulator, "step_grad"):
            # substep_grad re-sends the same wrench adjoint before every substep (mpm_simulator.py:344-346): once is enough
            if ext_f_grad_list:
                for i in range(self.simulator.n_primitive):
                    self.simulator.primitives[i].set_ext_f_grad(ext_f_grad_list[i])
            self.simulator.step_grad(start, self.substeps)
        else:
            for s in range(start - 1, self.simulator.cur - 1, -1):
                tmp = self.simulator.substep_grad(s, action=mpm_action, ext_f_grad=ext_f_grad_list if ext_f_grad_list else None)
                if tmp is not None:
                    mpm_action_grad += tmp
        if action is None:
            return None
        return mpm_action_grad if self.control_mode == "mpm" else rigid_action_grad

    def backward(self):                                                  # taichi_env.py:139-151
        if not self.rigid_velocity_control:
            self.rigid_simulator.state_grad = np.zeros(self.rigid_simulator.state_dim)
        total_steps = self.simulator.cur // self.substeps
        action_grad = []
        for s in range(total_steps - 1, -1, -1):
            action_grad = [self.step_grad(self.action_list[s])] + action_grad
        if not self.rigid_velocity_control:
            self.rigid_simulator.state_grad = self.rigid_simulator.state_grad + self.rigid_simulator.get_ext_state_grad(0)
        return np.vstack(action_grad)

    def compute_loss(self, f=None, **kwargs):
        assert self.loss is not None
        if f is None:
            if self._is_copy:                       # taichi_env.py:155-157: rolling mode evaluates frame 0 from a cleared loss
                self.loss.clear() if hasattr(self.loss, "clear") else None
                f = 0
            else:
                f = self.simulator.cur
        return self.loss.compute_loss(f, **kwargs)


def adjust_action_with_ext_force(env, actions):
    """``softmac/utils.py:76-113``: actions found WITHOUT external forces are corrected so that they also cancel what the bodies feel --
    a forward rollout in which, after the substeps of env step t, the averaged contact wrench ``ext_f / substeps`` of every primitive with
    ``enable_external_force`` plus the body's weight is subtracted from action t (layout per body: torque(3), force(3)) before the rigid
    step.  demo_pour starts from ``get_init_actions(choice=0, adjust=True)`` (demo_pour.py:95-110): zeros, adjusted -- the glass is held
    still against gravity and the liquid while it settles.  Leaves the env at the end of the rollout, like the reference; returns the
    adjusted (T, action_dim) array."""
    assert env.control_mode == "rigid" and not env._is_copy
    rigid, sim = env.rigid_simulator, env.simulator
    actions = np.array(actions, dtype=np.float64, copy=True)
    gravity = np.asarray(rigid.gravity, dtype=np.float64)
    for t in range(actions.shape[0]):
        start = sim.cur
        sim.cur = start + env.substeps
        if hasattr(sim, "step"):
            sim.step(start, env.substeps)
        else:
            for s in range(start, sim.cur):
                sim.substep(s)
        for i in range(rigid.n_primitive):
            if not env.primitives[i].enable_external_force:
                continue
            ext_f = np.asarray(env.primitives[i].ext_f.to_numpy(), dtype=np.float32).astype(np.float64) / env.substeps      # FloatTensor in the reference
            force = ext_f[:3] + rigid.bodies[i].mass * gravity
            o = rigid.offsets[i]
            assert rigid.bodies[i].ndof == 6, "adjust_action_with_ext_force: bodies on free joints (6 action components per body)"
            actions[t, o:o + 3] -= ext_f[3:]
            actions[t, o + 3:o + 6] -= force
        rigid.step(start // env.substeps, actions[t])
        env.action_list.append(actions[t])
    return actions
