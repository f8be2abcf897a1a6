"""Lock-step environment for B independent rollouts batched in ONE simulator handle (BASELINE config 4:
"64 independent demo_grip rollouts ... several rollouts per GPU batched in one handle").

The control flow per rollout is exactly ``TaichiEnv.step / step_grad / backward`` (softmac/engine/taichi_env.py:93-151);
what changes is the plumbing: the MPM substeps of all rollouts run in the same kernel launches, and the per-env-step
coupling traffic (wrench out, primitive states in, state adjoints out, wrench adjoints in) is ONE transfer for all
rollouts and primitives (``smx_*_all``) instead of B x P small ones.  Every rollout keeps its own rigid simulator
(the Jade bridge or the stand-in), which talks to ``BufferedPrimitiveView`` objects backed by host mirrors.
"""
import numpy as np

import ctypes as C

from .._capi import lib, check, d_ptr, as_d, SmxRigidLinear


class _ExtF:
    def __init__(self, v):
        self._v = v

    def to_numpy(self):
        return self._v.buf.ext_f[self._v.b, self._v.i].copy()


class BufferedPrimitiveView:
    """Primitive coupling surface of rollout `b`, primitive `i`, backed by a CouplingBuffer."""

    def __init__(self, buf, b, i, prim):
        self.buf, self.b, self.i = buf, b, i
        self.enable_external_force = prim.enable_external_force
        self.friction, self.softness = prim.friction, prim.softness
        self.ext_f = _ExtF(self)

    def clear_ext_f(self):                      # the device-side wrench was cleared for all rollouts right after the pull
        self.buf.ext_f[self.b, self.i] = 0.0

    def set_ext_f_grad(self, g):
        self.buf.ext_f_grad[self.b, self.i] = np.asarray(g, dtype=np.float64)

    def set_all_states(self, f, state, f_end=None):
        self.buf.note_state(self.b, self.i, f, (f + 1) if f_end is None else f_end, np.asarray(state, dtype=np.float64))

    def get_all_states_grad(self, f, f_end=None):
        return self.buf.state_grad_of(self.b, self.i, f, (f + 1) if f_end is None else f_end)

    def reset(self):
        self.clear_ext_f()


class _Views(list):
    def initialize(self):
        pass

    def reset(self):
        for p in self:
            p.reset()


class CouplingBuffer:
    def __init__(self, sim, primitives):
        self.sim, self.h = sim, sim._h
        self.B, self.P = sim.n_batch, len(primitives)
        self.ext_f = np.zeros((self.B, self.P, 6))
        self.ext_f_grad = np.zeros((self.B, self.P, 6))
        self.states = np.zeros((self.B, self.P, 13))
        self.state_range = None
        self.grads = np.zeros((self.B, self.P, 13))
        self.grad_range = None
        self.views = [_Views(BufferedPrimitiveView(self, b, i, primitives[i]) for i in range(self.P)) for b in range(self.B)]

    # wrench ------------------------------------------------------------------------------------------------------
    def pull_ext_f(self):
        check(lib().smx_get_ext_f_all(self.h, d_ptr(self.ext_f)))
        check(lib().smx_clear_ext_f_all(self.h))

    def push_ext_f_grads(self):
        check(lib().smx_set_ext_f_grads_all(self.h, d_ptr(self.ext_f_grad)))

    # primitive states ----------------------------------------------------------------------------------------------
    def note_state(self, b, i, f0, f1, s13):
        self.states[b, i] = s13
        if self.state_range is None:
            self.state_range = [f0, f1]
        else:
            self.state_range = [min(self.state_range[0], f0), max(self.state_range[1], f1)]

    def push_states(self):
        if self.state_range is not None:
            f0, f1 = self.state_range
            f1 = min(f1, self.sim.max_steps)
            if f0 < f1:
                check(lib().smx_set_primitive_states_all(self.h, int(f0), int(f1), d_ptr(self.states)))
            self.state_range = None

    # primitive state adjoints ----------------------------------------------------------------------------------------
    def pull_state_grads(self, f0, f1):
        f1 = min(f1, self.sim.max_steps)
        self.grads[:] = 0
        if f0 < f1:
            check(lib().smx_get_primitive_state_grads_all(self.h, int(f0), int(f1), d_ptr(self.grads)))
        self.grad_range = (f0, f1)

    def state_grad_of(self, b, i, f0, f1):
        """The rigid bridge sums get_all_states_grad(j) over the frames of an env step (rigid_simulator.py:207-208); the sum
        over the pulled range is returned for its first frame and zero for the others."""
        r0, r1 = self.grad_range
        if f0 <= r0 < f1:
            return self.grads[b, i].copy()
        return np.zeros(13)


class LinearBatchedRigid:
    """All B rigid stand-ins advanced at once with numpy matmuls.  Valid when every joint is fixed or prismatic: then
    ``RigidSimulator._advance`` and ``_pose`` are affine in (state, action, wrench), so their exact Jacobians are constant
    matrices, taken once from a prototype by finite differences.  Same coupling semantics as B separate bridges
    (wrench averaged over the env step and truncated to fp32, pose handed over in fp32, rigid_simulator.py:92-93,185)."""

    def __init__(self, proto, B):
        assert all(b.joint in ("fixed", "prismatic") for b in proto.bodies), "LinearBatchedRigid needs fixed / prismatic joints"
        self.p, self.B, self.P = proto, B, proto.n_primitive
        self.substeps, self.sd, self.ad = proto.substeps, proto.state_dim, proto.action_dim
        s0, a0, w0 = np.zeros(self.sd), np.zeros(self.ad), np.zeros(6 * self.P)
        self.c = proto._advance(s0, a0, w0)
        self.As = proto._jac(lambda x: proto._advance(x, a0, w0), s0).T            # s' = s As + a Aa + w Aw + c
        self.Aa = proto._jac(lambda x: proto._advance(s0, x, w0), a0).T if self.ad else np.zeros((0, self.sd))
        self.Aw = proto._jac(lambda x: proto._advance(s0, a0, x), w0).T
        self.pose0 = np.stack([proto._pose(s0, i) for i in range(self.P)])            # (P, 13)
        self.M = np.stack([proto._jac(lambda x, i=i: proto._pose(x, i), s0).T for i in range(self.P)])   # (P, sd, 13)
        self.enable = np.array([proto.primitives[i].enable_external_force for i in range(self.P)], dtype=bool)
        self.fp32 = proto.fp32_bridge
        self.reset()

    def reset(self):
        self.states = [np.tile(self.p.init_state, (self.B, 1))]
        self.masks = []
        self.state_grad = np.zeros((self.B, self.sd))

    def poses(self):
        s = self.states[-1]
        out = np.einsum("bs,psk->bpk", s, self.M) + self.pose0[None]
        return out.astype(np.float32).astype(np.float64) if self.fp32 else out

    def step(self, actions, ext_f):
        """actions (B, ad); ext_f (B, P, 6) accumulated over the env step.  Returns the (B, P, 13) poses of the next step."""
        w = np.asarray(ext_f, dtype=np.float64)
        if self.fp32:
            w = w.astype(np.float32).astype(np.float64)
        w = w / self.substeps
        mask = ((np.abs(w) > 1e-10).any(axis=2) & self.enable[None]).astype(np.float64)          # rigid_simulator.py:96
        w = w * mask[:, :, None]
        self.masks.append(mask)
        a = np.zeros((self.B, self.ad)) if actions is None else np.asarray(actions, dtype=np.float64).reshape(self.B, self.ad)
        self.states.append(self.states[-1] @ self.As + a @ self.Aa + w.reshape(self.B, -1) @ self.Aw + self.c[None])
        return self.poses()

    def step_grad(self, k, pose_grads):
        """pose_grads (B, P, 13): primitive-state adjoints summed over the frames of env step k+1.
        Returns (action grads (B, ad), wrench adjoints (B, P, 6))."""
        self.state_grad = self.state_grad + np.einsum("bpk,psk->bs", pose_grads, self.M) * self.p.ext_grad_scale
        ag = self.state_grad @ self.Aa.T
        gw = (self.state_grad @ self.Aw.T).reshape(self.B, self.P, 6) * self.masks[k][:, :, None] / self.substeps
        self.state_grad = self.state_grad @ self.As.T
        return ag, gw

    def finish(self, pose_grads0):
        self.state_grad = self.state_grad + np.einsum("bpk,psk->bs", pose_grads0, self.M)


class DeviceLinearRigid:
    """The rigid stand-in moved onto the GPU (smx_rigid_linear_*, softmac_b200/csrc/smx_rigid.cuh): fixed, prismatic, revolute and free
    joints; the integrator's constant matrices (``RigidSimulator._advance`` is affine in state, action and wrench) plus the
    closed-form pose map of every body, one small kernel per env step on the simulator's stream, so an episode runs without a
    single host <-> device round trip of the coupling (wrench, poses, state adjoints and wrench adjoints never leave the
    device; rigid_simulator.py:85-220)."""

    JOINT = {"fixed": 0, "prismatic": 1, "free": 2, "revolute": 3}

    def __init__(self, proto, sim, n_batch, max_env_steps):
        self.p, self.sim, self.h, self.K = proto, sim, sim._h, int(max_env_steps)
        self.B, self.sd, self.ad, self.P = int(n_batch), proto.state_dim, proto.action_dim, proto.n_primitive
        s0, a0, w0 = np.zeros(self.sd), np.zeros(self.ad), np.zeros(6 * self.P)
        c = proto._advance(s0, a0, w0)                                               # s' = s As + a Aa + w Aw + c (exactly affine: eps = 1)
        As = proto._jac(lambda x: proto._advance(x, a0, w0), s0, eps=1.0).T
        Aa = proto._jac(lambda x: proto._advance(s0, x, w0), a0, eps=1.0).T if self.ad else np.zeros((0, self.sd))
        Aw = proto._jac(lambda x: proto._advance(s0, a0, x), w0, eps=1.0).T
        body = np.array([np.concatenate([b.origin, b.quat0, b.axis]) for b in proto.bodies])
        joint = np.array([[self.JOINT[b.joint], int(proto.offsets[i])] for i, b in enumerate(proto.bodies)], dtype=np.int32)
        enable = np.array([bool(proto.primitives[i].enable_external_force) for i in range(self.P)], dtype=np.int32)
        keep = [as_d(As), as_d(Aa), as_d(Aw), as_d(c), as_d(body), as_d(proto.init_state)]
        d = SmxRigidLinear()
        d.state_dim, d.action_dim, d.max_env_steps, d.fp32_bridge = self.sd, self.ad, self.K, int(bool(proto.fp32_bridge))
        d.ext_grad_scale = float(proto.ext_grad_scale)
        d.As, d.Aa, d.Aw, d.c, d.body, d.init_state = [d_ptr(a) for a in keep]
        d.joint = np.ascontiguousarray(joint).ctypes.data_as(C.POINTER(C.c_int32))
        d.enable = enable.ctypes.data_as(C.POINTER(C.c_int32))
        check(lib().smx_rigid_linear_create(self.h, C.byref(d)))

    def reset(self):
        check(lib().smx_rigid_linear_reset(self.h))

    def step(self, k, actions):
        if self.ad:
            a = np.zeros((self.B, self.ad)) if actions is None else as_d(actions, (self.B, self.ad))
            check(lib().smx_rigid_linear_set_actions(self.h, int(k), d_ptr(a)))
        check(lib().smx_rigid_linear_step(self.h, int(k)))

    def step_grad(self, k):
        check(lib().smx_rigid_linear_step_grad(self.h, int(k)))

    def finish(self):
        check(lib().smx_rigid_linear_finish(self.h))

    def action_grads(self, k0, k1):
        """(B, k1 - k0, action_dim)"""
        out = np.zeros((k1 - k0, self.B, self.ad))
        if self.ad and k1 > k0:
            check(lib().smx_rigid_linear_get_action_grads(self.h, int(k0), int(k1), d_ptr(out)))
        return np.ascontiguousarray(out.transpose(1, 0, 2))

    def states(self, k):
        out = np.zeros((self.B, self.sd))
        check(lib().smx_rigid_linear_get_states(self.h, int(k), d_ptr(out)))
        return out

    @property
    def state_grad(self):
        out = np.zeros((self.B, self.sd))
        check(lib().smx_rigid_linear_get_state_grad(self.h, d_ptr(out)))
        return out


class BatchedTaichiEnv:
    def __init__(self, simulator, primitives, make_rigid, init_particles, loss=None, vectorize=True, device_rigid=False):
        """make_rigid(b, views) -> a rigid simulator (RigidSimulator surface) for rollout b talking to `views`.
        vectorize: when every joint of the stand-in is fixed / prismatic, advance all rollouts with LinearBatchedRigid
        (numpy matmuls) instead of B Python bridges.
        device_rigid: run the stand-in on the GPU (DeviceLinearRigid; fixed / prismatic / free joints): no host round trip per
        env step; ``step_grad`` then returns None and ``backward`` reads all action gradients once at the end."""
        self.simulator, self.primitives, self.loss = simulator, primitives, loss
        self.B, self.substeps = simulator.n_batch, simulator.substeps
        self.buf = CouplingBuffer(simulator, primitives)
        self.rigid = [make_rigid(0, self.buf.views[0])]
        self.vec = None
        if device_rigid:
            pass                                # one prototype bridge is enough: it only supplies the matrices and body descriptors
        elif vectorize and hasattr(self.rigid[0], "bodies") and all(b.joint in ("fixed", "prismatic") for b in self.rigid[0].bodies):
            self.vec = LinearBatchedRigid(self.rigid[0], self.B)
        else:
            self.rigid += [make_rigid(b, self.buf.views[b]) for b in range(1, self.B)]
        self.init_particles = np.asarray(init_particles, dtype=np.float64)
        self.action_list = []
        self.dev = None
        if device_rigid:
            if not hasattr(self.rigid[0], "bodies"):
                raise ValueError("device_rigid needs the stand-in rigid simulator (bodies on fixed / prismatic / free joints)")
            self.dev = DeviceLinearRigid(self.rigid[0], simulator, self.B, max(simulator.max_steps // max(self.substeps, 1), 1))
        primitives.initialize()
        simulator.initialize()
        self.reset()

    def reset(self):
        self.primitives.reset()
        self.simulator.reset(self.init_particles)
        if self.dev:
            self.dev.reset()
        elif self.vec:
            self.vec.reset()
            self._push_poses(self.vec.poses(), 0, self.substeps)
        else:
            for r in self.rigid:
                r.reset()                  # writes the initial pose of frames [0, substeps) into the buffer
            self.buf.push_states()
        check(lib().smx_clear_ext_f_all(self.simulator._h))
        self.action_list = []

    def step(self, actions):
        """actions: (B, action_dim)."""
        sim = self.simulator
        start = sim.cur
        sim.cur = start + self.substeps
        self.action_list.append(np.asarray(actions, dtype=np.float64))
        sim.step(start, self.substeps)
        k = start // self.substeps
        if self.dev:
            self.dev.step(k, actions)
            return
        self.buf.pull_ext_f()
        if self.vec:
            self._push_poses(self.vec.step(actions, self.buf.ext_f), (k + 1) * self.substeps, (k + 2) * self.substeps)
            return
        for b, r in enumerate(self.rigid):
            r.step(k, actions[b])
        self.buf.push_states()

    def _push_poses(self, poses, f0, f1):
        f1 = min(f1, self.simulator.max_steps)
        if f0 < f1:
            self.buf.states[:] = poses
            check(lib().smx_set_primitive_states_all(self.simulator._h, int(f0), int(f1), d_ptr(self.buf.states)))

    def step_grad(self, actions):
        sim = self.simulator
        start = sim.cur
        sim.cur = start - self.substeps
        k = sim.cur // self.substeps
        if self.dev:
            self.dev.step_grad(k)
            sim.step_grad(start, self.substeps)
            return None
        # rigid.step_grad(k) pulls the primitive-state adjoints of env step k+1 (frames [(k+1) sub, (k+2) sub))
        self.buf.pull_state_grads((k + 1) * self.substeps, (k + 2) * self.substeps)
        if self.vec:
            ag, gw = self.vec.step_grad(k, self.buf.grads)
            self.buf.ext_f_grad[:] = gw
            self.buf.push_ext_f_grads()
            sim.step_grad(start, self.substeps)
            return ag
        grads = []
        for b, r in enumerate(self.rigid):
            ag, ext_list = r.step_grad(k, actions[b])
            grads.append(ag)
            for i, g in enumerate(ext_list):
                self.buf.ext_f_grad[b, i] = g
        self.buf.push_ext_f_grads()
        sim.step_grad(start, self.substeps)
        return np.stack(grads)

    def backward(self):
        total = self.simulator.cur // self.substeps
        if self.vec:
            self.vec.state_grad = np.zeros((self.B, self.vec.sd))
        for r in self.rigid:
            r.state_grad = np.zeros(r.state_dim)
        if self.dev:
            for s in range(total - 1, -1, -1):
                self.step_grad(self.action_list[s])
            self.dev.finish()
            return self.dev.action_grads(0, total)
        out = []
        for s in range(total - 1, -1, -1):
            out = [self.step_grad(self.action_list[s])] + out
        self.buf.pull_state_grads(0, self.substeps)
        if self.vec:
            self.vec.finish(self.buf.grads)
        else:
            for r in self.rigid:
                r.state_grad = r.state_grad + r.get_ext_state_grad(0)
        return np.stack(out, axis=1)            # (B, steps, action_dim)

    def rigid_states(self):
        """(B, state_dim) current rigid states."""
        if self.dev:
            return self.dev.states(len(self.action_list))
        return self.vec.states[-1].copy() if self.vec else np.stack([r.states[-1] for r in self.rigid])
