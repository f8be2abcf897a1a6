"""Host-side mirror of ``softmac/engine/primitive/primitive_base.py:Primitive``.

The Taichi fields of the reference (position/rotation/v/w per frame with grads, ext_f with grad, the SDF
tables) live in device memory owned by libsoftmac_b200.so; this class keeps the reference's method names
and argument meaning and forwards them through the C ABI.  Until the primitive is attached to a simulator
(``MPMSimulator.__init__`` does that, as ``MPMSimulator(cfg, primitives, ...)`` in the reference,
softmac/engine/taichi_env.py:38) state writes are buffered on the host.
"""
import numpy as np

from ..._capi import lib, check, as_d, d_ptr, ip


class _ScalarField:
    """``field[None]`` get/set like a 0-d Taichi field (primitive_base.py:26-27)."""

    def __init__(self, value, on_set):
        self._v, self._on_set = float(value), on_set

    def __getitem__(self, _):
        return self._v

    def __setitem__(self, _, v):
        self._v = float(v)
        self._on_set()


class _ExtF:
    """``primitive.ext_f.to_numpy()`` (rigid_simulator.py:92)."""

    def __init__(self, prim):
        self._p = prim

    def to_numpy(self):
        return self._p.get_ext_f()

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a if dtype is None else a.astype(dtype)


class PrimitiveBatchView:
    """The coupling surface of a primitive for ONE rollout of a batched handle (MPMSimulator(n_batch > 1)): same
    methods as ``Primitive`` (set_all_states / get_all_states_grad / ext_f / clear_ext_f / set_ext_f_grad), addressed to
    batch ``b`` through the ``smx_*_b`` entry points.  A rigid simulator per rollout is handed these views."""

    def __init__(self, prim, batch):
        self._p, self._b = prim, int(batch)
        self.ext_f = _ExtF(self)
        self.enable_external_force = prim.enable_external_force
        self.friction, self.softness = prim.friction, prim.softness

    def get_ext_f(self):
        o = np.zeros(6)
        check(lib().smx_get_ext_f_b(self._p._sim, self._b, self._p._id, d_ptr(o)))
        return o

    def clear_ext_f(self):
        check(lib().smx_clear_ext_f_b(self._p._sim, self._b, self._p._id))

    def set_ext_f_grad(self, g):
        g = as_d(np.asarray(g, dtype=np.float64), (6,))
        check(lib().smx_set_ext_f_grad_b(self._p._sim, self._b, self._p._id, d_ptr(g)))

    def set_all_states(self, f, state, f_end=None):
        s = as_d(np.asarray(state, dtype=np.float64), (13,))
        check(lib().smx_set_primitive_state_b(self._p._sim, self._b, self._p._id, f, (f + 1) if f_end is None else f_end, d_ptr(s)))

    def get_all_states(self, f):
        o = np.zeros(13)
        check(lib().smx_get_primitive_state_b(self._p._sim, self._b, self._p._id, f, d_ptr(o)))
        return o

    def get_state(self, f):
        return self.get_all_states(f)[:7]

    def get_all_states_grad(self, f, f_end=None):
        o = np.zeros(13)
        check(lib().smx_get_primitive_state_grad_b(self._p._sim, self._b, self._p._id, f, (f + 1) if f_end is None else f_end, d_ptr(o)))
        return o

    def add_all_states_grad(self, f, g13):
        g = as_d(np.asarray(g13, dtype=np.float64), (13,))
        check(lib().smx_add_primitive_state_grad_b(self._p._sim, self._b, self._p._id, f, d_ptr(g)))

    def reset(self):
        self.clear_ext_f()


class Primitive:
    state_dim = 7

    def view(self, batch):
        self._need()
        return PrimitiveBatchView(self, batch)

    def __init__(self, cfg=None, dim=3, max_timesteps=2048, dtype="float64", rigid_velocity_control=False, **kwargs):
        defaults = self.default_config()
        if cfg is not None:
            defaults.update({k: cfg[k] for k in cfg})
        defaults.update(kwargs)
        self.cfg = defaults
        self.dim, self.max_timesteps, self.dtype = dim, max_timesteps, dtype
        self.rotation_dim, self.angular_velocity_dim = 4, 3
        self.friction = _ScalarField(self.cfg.get("friction", 0.9), self._push_params)
        self.softness = _ScalarField(0.0, self._push_params)
        self.enable_external_force = self.cfg.get("enable_external_force", True)
        self.rigid_velocity_control = rigid_velocity_control
        self.ext_f = _ExtF(self)
        self._sim, self._id = None, -1
        self._pending = {}      # frame -> state13 written before the primitive was attached
        self.sdf_table = self.normal_table = self.sdf_lower = self.sdf_upper = None
        self.sdf_dx = 0.0

    # -- attachment ---------------------------------------------------------------------------------
    def _attach(self, sim_handle, contact_enabled=True):
        L = lib()
        if self.sdf_table is None:
            pid = check(L.smx_add_primitive(sim_handle, None, None, None, None, None, 1.0, self.friction[None],
                                            self.softness[None], int(contact_enabled)))
        else:
            sdf, nrm = as_d(self.sdf_table), as_d(self.normal_table)
            res = np.ascontiguousarray(sdf.shape, dtype=np.int32)
            lo, up = as_d(self.sdf_lower), as_d(self.sdf_upper)
            pid = check(L.smx_add_primitive(sim_handle, d_ptr(sdf), d_ptr(nrm), res.ctypes.data_as(ip), d_ptr(lo), d_ptr(up),
                                            float(self.sdf_dx), self.friction[None], self.softness[None], int(contact_enabled)))
        self._sim, self._id = sim_handle, pid
        for f, s13 in sorted(self._pending.items()):
            self.set_all_states(f, s13)
        self._pending.clear()
        return pid

    def _push_params(self):
        if self._sim is not None:
            check(lib().smx_set_primitive_params(self._sim, self._id, self.friction[None], self.softness[None]))

    def _need(self):
        if self._sim is None:
            raise RuntimeError("primitive is not attached to an MPMSimulator")

    # -- wrench (primitive_base.py:183-192) ------------------------------------------------------------
    def get_ext_f(self):
        self._need()
        o = np.zeros(6)
        check(lib().smx_get_ext_f(self._sim, self._id, d_ptr(o)))
        return o

    def clear_ext_f(self):
        if self._sim is not None:
            check(lib().smx_clear_ext_f(self._sim, self._id))

    def set_ext_f_grad(self, ext_f_grad):
        self._need()
        g = as_d(np.asarray(ext_f_grad, dtype=np.float64), (6,))
        check(lib().smx_set_ext_f_grad(self._sim, self._id, d_ptr(g)))

    # -- state plumbing (primitive_base.py:204-275) ------------------------------------------------------
    def set_all_states(self, f, state, f_end=None):
        s = as_d(np.asarray(state, dtype=np.float64), (13,))
        if self._sim is None:
            for ff in range(f, (f + 1) if f_end is None else f_end):
                self._pending[ff] = s.copy()
            return
        check(lib().smx_set_primitive_state(self._sim, self._id, f, (f + 1) if f_end is None else f_end, d_ptr(s)))

    def get_all_states(self, f):
        if self._sim is None:
            return self._pending.get(f, np.zeros(13)).copy()
        o = np.zeros(13)
        check(lib().smx_get_primitive_state(self._sim, self._id, f, d_ptr(o)))
        return o

    def get_state(self, f):
        return self.get_all_states(f)[:7]

    def set_state(self, f, state):
        ss = self.get_all_states(f)
        ss[:len(state)] = state
        self.set_all_states(f, ss)

    def get_all_states_grad(self, f, f_end=None):
        self._need()
        o = np.zeros(13)
        check(lib().smx_get_primitive_state_grad(self._sim, self._id, f, (f + 1) if f_end is None else f_end, d_ptr(o)))
        return o

    def add_all_states_grad(self, f, g13):
        """New seam: what Taichi losses do by writing position/rotation/v/w.grad[f] directly."""
        self._need()
        g = as_d(np.asarray(g13, dtype=np.float64), (13,))
        check(lib().smx_add_primitive_state_grad(self._sim, self._id, f, d_ptr(g)))

    def clear_all_states(self):
        if self._sim is None:
            self._pending.clear()
            return
        # state series, their adjoints, the action buffer and its adjoint, over the frames the handle holds (the reference clears
        # [0, max_timesteps) of its own fields, primitive_base.py:236-246; here the series live in the simulator, sized max_steps)
        check(lib().smx_reset_primitive(self._sim, self._id))

    def initialize(self):
        self.friction[None] = self.cfg.get("friction", 0.9)
        self.reset()

    def reset(self):
        self.clear_all_states()
        self.clear_ext_f()

    # -- velocity control (primitive_base.py:277-326) -----------------------------------------------------
    def set_action(self, s, n, action):
        self._need()
        a = as_d(np.asarray(action, dtype=np.float64), (6,))
        check(lib().smx_set_primitive_action(self._sim, self._id, s, n, d_ptr(a)))

    def get_action_grad(self, s, n):
        self._need()
        o = np.zeros(6)
        check(lib().smx_get_primitive_action_grad(self._sim, self._id, s, n, d_ptr(o)))
        return o

    @classmethod
    def default_config(cls):
        from ...config import CfgNode
        return CfgNode(friction=0.9, enable_external_force=True, urdf_path="")
