"""ctypes front end of the float64 CPU oracle (oracle/mpm_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  Nothing under softmac_b200/ may import this module.

PARITY UNPINNED: the reference (Taichi 1.4.1 + Jade) cannot be installed or run in this image and
ships no golden vectors, so the oracle is pinned by source correspondence (file:line citations in the
C file), by finite differences of its own forward, and by conservation properties only.

The class mirrors the slice of ``softmac/engine/mpm_simulator.py:MPMSimulator`` and
``softmac/engine/primitive/primitive_base.py:Primitive`` that the substep path touches.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "mpm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-std=c11", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        L = _lib
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp, C.c_double,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_plasticity.argtypes = [vp, C.c_int, C.c_double]
        L.orc_add_primitive.restype = C.c_int
        L.orc_add_primitive.argtypes = [vp, dp, dp, ip, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_set_primitive_enabled.argtypes = [vp, C.c_int, C.c_int]
        L.orc_set_primitive_params.argtypes = [vp, C.c_int, C.c_double, C.c_double]
        for nm in ("orc_set_frame", "orc_get_frame", "orc_add_frame_grad", "orc_get_frame_grad"):
            getattr(L, nm).argtypes = [vp, C.c_int, dp]
        L.orc_clear_grads.argtypes = [vp]
        for nm in ("orc_set_primitive_state", "orc_get_primitive_state", "orc_get_primitive_state_grad",
                   "orc_add_primitive_state_grad"):
            getattr(L, nm).argtypes = [vp, C.c_int, C.c_int, dp]
        L.orc_get_ext_f.argtypes = [vp, C.c_int, dp]
        L.orc_clear_ext_f.argtypes = [vp, C.c_int]
        L.orc_set_ext_f_grad.argtypes = [vp, C.c_int, dp]
        L.orc_set_action.argtypes = [vp, dp]
        L.orc_get_action_grad.argtypes = [vp, dp]
        L.orc_set_control_idx.argtypes = [vp, ip]
        L.orc_set_primitive_action.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
        L.orc_get_primitive_action_grad.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
        L.orc_substep.argtypes = [vp, C.c_int]
        L.orc_substep_grad.argtypes = [vp, C.c_int]
        L.orc_svd3.argtypes = [dp, dp, dp, dp]
        L.orc_backward_svd.argtypes = [dp] * 7
        L.orc_prim_sdf.restype = C.c_double
        L.orc_prim_sdf.argtypes = [vp, C.c_int, C.c_int, dp]
        L.orc_prim_normal.argtypes = [vp, C.c_int, C.c_int, dp, dp]
        L.orc_get_grid.argtypes = [vp, dp, dp, dp]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _arr(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class OracleSim:
    """f64 restatement of MPMSimulator (+ its primitives) -- see module docstring."""

    def __init__(self, n_particles, n_grid=64, max_steps=8, dt=2e-4, E=3e3, nu=0.2, gravity=(0., -9.8, 0.),
                 ground_friction=20., material_model=0, ptype=0, collision_type=2, substeps=5, n_control=0,
                 rigid_velocity_control=False):
        L = lib()
        g = _arr(gravity)
        self.n, self.n_grid, self.max_steps, self.n_control = n_particles, n_grid, max_steps, n_control
        self.dt, self.substeps = dt, substeps
        self.h = L.orc_create(n_particles, n_grid, max_steps, dt, E, nu, _d(g), ground_friction, material_model,
                              ptype, collision_type, substeps, n_control, int(rigid_velocity_control))
        self.n_primitive = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def set_plasticity(self, mode, yield_stress):
        """mode 1: von Mises return mapping with cfg.yield_stress (soft_cloth/engine/mpm_simulator.py:232) instead of the sigma clip."""
        lib().orc_set_plasticity(self.h, int(mode), float(yield_stress))

    # ---- primitives -------------------------------------------------------------------------
    def add_primitive(self, sdf=None, normal=None, lower=None, upper=None, sdf_dx=0.0, friction=0.9,
                      softness=666., enabled=True):
        L = lib()
        if sdf is None:
            r = L.orc_add_primitive(self.h, None, None, None, None, None, 1.0, friction, softness, int(enabled))
        else:
            sdf = _arr(sdf); normal = _arr(normal)
            res = np.ascontiguousarray(sdf.shape, dtype=np.int32)
            lo, up = _arr(lower), _arr(upper)
            r = L.orc_add_primitive(self.h, _d(sdf), _d(normal), res.ctypes.data_as(C.POINTER(C.c_int)), _d(lo), _d(up),
                                    float(sdf_dx), friction, softness, int(enabled))
        assert r >= 0
        self.n_primitive += 1
        return r

    def set_primitive_enabled(self, i, flag):
        lib().orc_set_primitive_enabled(self.h, i, int(flag))

    def set_primitive_state(self, i, f, s13):
        s = _arr(s13); assert s.size == 13
        lib().orc_set_primitive_state(self.h, i, f, _d(s))

    def get_primitive_state(self, i, f):
        o = np.zeros(13); lib().orc_get_primitive_state(self.h, i, f, _d(o)); return o

    def get_primitive_state_grad(self, i, f):
        o = np.zeros(13); lib().orc_get_primitive_state_grad(self.h, i, f, _d(o)); return o

    def add_primitive_state_grad(self, i, f, g13):
        g = _arr(g13); lib().orc_add_primitive_state_grad(self.h, i, f, _d(g))

    def get_ext_f(self, i):
        o = np.zeros(6); lib().orc_get_ext_f(self.h, i, _d(o)); return o

    def clear_ext_f(self, i):
        lib().orc_clear_ext_f(self.h, i)

    def set_ext_f_grad(self, i, g6):
        g = _arr(g6); lib().orc_set_ext_f_grad(self.h, i, _d(g))

    def set_primitive_action(self, i, s, n, a6):
        a = _arr(a6); lib().orc_set_primitive_action(self.h, i, s, n, _d(a))

    def get_primitive_action_grad(self, i, s, n):
        o = np.zeros(6); lib().orc_get_primitive_action_grad(self.h, i, s, n, _d(o)); return o

    def sdf(self, i, f, pos):
        p = _arr(pos); return lib().orc_prim_sdf(self.h, i, f, _d(p))

    def normal(self, i, f, pos):
        p = _arr(pos); o = np.zeros(3); lib().orc_prim_normal(self.h, i, f, _d(p), _d(o)); return o

    # ---- particle state -----------------------------------------------------------------------
    def set_frame(self, f, st24):
        s = _arr(st24, (self.n, 24)); lib().orc_set_frame(self.h, f, _d(s))

    def get_frame(self, f):
        o = np.zeros((self.n, 24)); lib().orc_get_frame(self.h, f, _d(o)); return o

    def add_frame_grad(self, f, g24):
        g = _arr(g24, (self.n, 24)); lib().orc_add_frame_grad(self.h, f, _d(g))

    def get_frame_grad(self, f):
        o = np.zeros((self.n, 24)); lib().orc_get_frame_grad(self.h, f, _d(o)); return o

    def clear_grads(self):
        lib().orc_clear_grads(self.h)

    def set_action(self, a):
        a = _arr(a, (self.n_control, 3)); lib().orc_set_action(self.h, _d(a))

    def get_action_grad(self):
        o = np.zeros((max(self.n_control, 1), 3)); lib().orc_get_action_grad(self.h, _d(o)); return o[:self.n_control]

    def set_control_idx(self, idx):
        i = np.ascontiguousarray(idx, dtype=np.int32); lib().orc_set_control_idx(self.h, i.ctypes.data_as(C.POINTER(C.c_int)))

    def substep(self, f):
        lib().orc_substep(self.h, f)

    def substep_grad(self, f):
        lib().orc_substep_grad(self.h, f)

    def get_grid(self):
        G = self.n_grid ** 3
        a, m, o = np.zeros((G, 3)), np.zeros(G), np.zeros((G, 3))
        lib().orc_get_grid(self.h, _d(a), _d(m), _d(o))
        return a, m, o


def svd3(F):
    F = _arr(F, (3, 3)); U, S, V = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
    lib().orc_svd3(_d(F), _d(U), _d(S), _d(V)); return U, S, V


def backward_svd(gu, gs, gv, u, s, v):
    a = [_arr(t, (3, 3)) for t in (gu, gs, gv, u, s, v)]; o = np.zeros((3, 3))
    lib().orc_backward_svd(*[_d(t) for t in a], _d(o)); return o


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
