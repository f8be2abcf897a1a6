"""``MPMSimulator`` -- drop-in for ``softmac/engine/mpm_simulator.py:MPMSimulator`` backed by
libsoftmac_b200.so (hand-written sm_100a kernels; include/softmac_b200.h).

Same constructor, attributes and methods as the reference class; numpy float64 at the boundary, fp32
SoA checkpoints in HBM behind it.  The state / gradient calls also take float32 host arrays (no conversion
pass; ``pin`` them once for plain DMA) and device arrays (anything with ``__cuda_array_interface__``, e.g. the
optimiser's torch CUDA tensors): ``reset``, ``get_state``, ``get_grad``, ``get_state_grad``, ``add_x_grad``,
``add_state_grad``.  New seams the Taichi version exposed implicitly:
``add_x_grad`` / ``add_state_grad`` (loss -> adjoint seeds), ``clear_all_gradients`` and a settable
``primitives_contact`` list.
"""
import ctypes as C

import numpy as np

from .._capi import SmxConfig, lib, check, as_d, d_ptr, vp, SMX_FLAG_DENSE_GRID, SMX_FLAG_NO_SORT  # noqa: F401

fp = C.POINTER(C.c_float)


def _f32(a):
    """a contiguous float32 host array (the fp32 entry points take the caller's buffer as it is)"""
    return isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous


def _dev(a):
    """(device pointer, shape) of a float32 C-contiguous device array (torch CUDA tensor, cupy array, ...) or None"""
    cai = getattr(a, "__cuda_array_interface__", None)
    if cai is None:
        return None
    shape = tuple(int(v) for v in cai["shape"])
    dense = tuple(4 * int(np.prod(shape[i + 1:])) for i in range(len(shape)))
    if cai["typestr"] not in ("<f4", "=f4") or (cai.get("strides") is not None and tuple(cai["strides"]) != dense):
        raise TypeError("device arrays handed to MPMSimulator must be float32 and C-contiguous")
    return int(cai["data"][0]), tuple(cai["shape"])


def _f_ptr(a):
    return a.ctypes.data_as(fp)


MODEL_COROTATED, MODEL_NEOHOOKEAN = 0, 1
MAT_PLASTIC, MAT_ELASTIC, MAT_LIQUID = 0, 1, 2
CONTACT_GRID, CONTACT_PARTICLE, CONTACT_MIXED = 0, 1, 2


class _ContactFlags(list):
    """``sim.primitives_contact = [False, True, True]`` / ``sim.primitives_contact[i] = x`` (demo_grip.py:117)."""

    def __init__(self, sim, vals):
        super().__init__(vals)
        self._sim = sim

    def __setitem__(self, i, v):
        super().__setitem__(i, v)
        self._sim._push_contact()


class MPMSimulator:
    def __init__(self, cfg, primitives=(), env_dt=2e-3, rigid_velocity_control=False, device=0, sort_every=None,
                 flags=0, stream=None, n_batch=1):
        dim = self.dim = cfg.dim
        assert dim == 3, "the B200 path implements the 3-D simulator used by every reference config"
        assert cfg.dtype == "float64"      # mpm_simulator.py:19 -- boundary dtype; device storage is fp32
        self.dtype = np.float64
        self._yield_stress = cfg.yield_stress
        self.ground_friction = cfg.ground_friction
        self.default_gravity = cfg.gravity
        self.n_primitive = len(primitives)
        quality = cfg.quality * 0.5
        self.n_particles = cfg.n_particles          # per rollout
        self.n_batch = max(int(n_batch), 1)         # independent rollouts batched in this handle (new; reference: 1)
        self.n_total = self.n_particles * self.n_batch
        self.n_grid = int(128 * quality)
        self.dx, self.inv_dx = 1 / self.n_grid, float(self.n_grid)
        self.dt = cfg.dt
        self.p_vol, self.p_rho = (self.dx * 0.5) ** 2, 1
        self.p_mass = self.p_vol * self.p_rho
        self.ptype, self.material_model = cfg.ptype, cfg.material_model
        E, nu = cfg.E, cfg.nu
        self._mu, self._lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
        if self.ptype == 1:
            self._mu, self._lam = 0.3 * self._mu, 0.3 * self._lam
        elif self.ptype == 2:
            self._mu = 0.0
        self.max_steps = cfg.max_steps
        self.substeps = int(env_dt / self.dt)
        self.primitives = primitives
        self.rigid_velocity_control = rigid_velocity_control
        self.n_control = cfg.n_controllers
        self.collision_type = cfg.collision_type
        self.cur = 0
        self.use_graphs = False         # step / step_grad as one CUDA-graph launch per call (small, launch-latency-bound scenes)

        c = SmxConfig()
        c.n_particles, c.n_grid, c.max_steps = self.n_particles, self.n_grid, self.max_steps
        c.dt, c.E, c.nu = cfg.dt, E, nu
        c.gravity = (C.c_double * 3)(*[float(g) for g in cfg.gravity])
        c.ground_friction = cfg.ground_friction
        c.material_model, c.ptype, c.collision_type = cfg.material_model, cfg.ptype, cfg.collision_type
        c.substeps = max(self.substeps, 1)
        c.n_control = self.n_control
        c.rigid_velocity_control = int(rigid_velocity_control)
        c.sort_every = max(self.substeps, 4) if sort_every is None else sort_every
        c.device, c.flags = device, flags
        c.stream = stream
        c.n_batch = self.n_batch
        h = vp()
        check(lib().smx_create(C.byref(c), C.byref(h)))
        self._h = h
        self._primitives_contact = _ContactFlags(self, [True] * self.n_primitive)
        for i in range(self.n_primitive):
            pid = self.primitives[i]._attach(self._h, True)
            assert pid == i
        # cfg.plasticity = "von_mises": the soft_cloth variant's flow rule with cfg.yield_stress (soft_cloth/engine/mpm_simulator.py:20,232)
        self.plasticity = "clip"
        if str(getattr(cfg, "plasticity", "clip")) == "von_mises":
            self.set_plasticity("von_mises", float(cfg.yield_stress))

    def set_plasticity(self, mode, yield_stress=50.):
        """"clip": sigma clip of softmac (mpm_simulator.py:226-229); "von_mises": the return mapping the soft_cloth variant runs
        (soft_cloth/engine/mpm_simulator.py:172-189, call site :232) with cfg.yield_stress."""
        m = {"clip": 0, "von_mises": 1, 0: 0, 1: 1}[mode]
        check(lib().smx_set_plasticity(self._h, m, float(yield_stress)))
        self.plasticity, self._yield_stress = ("von_mises" if m else "clip"), float(yield_stress)

    # ------------------------------------------------------------------------------------------------
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().smx_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def primitives_contact(self):
        return self._primitives_contact

    @primitives_contact.setter
    def primitives_contact(self, vals):
        vals = list(vals)
        assert len(vals) == self.n_primitive
        self._primitives_contact = _ContactFlags(self, vals)
        self._push_contact()

    def _push_contact(self):
        for i, v in enumerate(self._primitives_contact):
            check(lib().smx_set_primitive_contact(self._h, i, int(bool(v))))

    def initialize(self):
        # gravity, mu, lam, yield_stress are scalars baked into the handle at construction (mpm_simulator.py:86-90)
        pass

    # -- the hot path (mpm_simulator.py:320-378) --------------------------------------------------------
    def substep(self, s, action=None):
        if action is not None:
            self.set_action(action)
        check(lib().smx_substep(self._h, int(s)))

    def substep_grad(self, s, action=None, ext_f_grad=None):
        if action is not None:
            self.set_action(action)
        if ext_f_grad is not None:
            for i in range(self.n_primitive):
                self.primitives[i].set_ext_f_grad(ext_f_grad[i])
        check(lib().smx_substep_grad(self._h, int(s)))
        if action is None:
            return None
        g = np.zeros((self.n_batch * self.n_control, self.dim))
        check(lib().smx_get_action_grad(self._h, d_ptr(g)))
        return g.reshape(np.shape(action)) if np.size(action) == g.size else g

    def get_action_grad(self):
        """action.grad accumulated since the last set_action (the reference returns it from substep_grad, mpm_simulator.py:376-378)."""
        g = np.zeros((self.n_batch * self.n_control, self.dim))
        check(lib().smx_get_action_grad(self._h, d_ptr(g)))
        return g

    def step(self, s0, count, graph=None):
        """`count` substeps in one native call (the inner loop of TaichiEnv.step, taichi_env.py:101-102).  graph: replay the call as ONE
        CUDA-graph launch (smx_step_graph; default: the simulator's `use_graphs` attribute) -- for launch-latency-bound small scenes."""
        if self.use_graphs if graph is None else graph:
            check(lib().smx_step_graph(self._h, int(s0), int(count)))
        else:
            check(lib().smx_step(self._h, int(s0), int(count)))

    def step_grad(self, s1, count, graph=None):
        if self.use_graphs if graph is None else graph:
            check(lib().smx_step_grad_graph(self._h, int(s1), int(count)))
        else:
            check(lib().smx_step_grad(self._h, int(s1), int(count)))

    def graph_status(self):
        out = (C.c_int64 * 3)()
        check(lib().smx_graph_status(self._h, out))
        return dict(graph_calls=int(out[0]), plain_calls=int(out[1]), reinstantiations=int(out[2]))

    # -- IO (mpm_simulator.py:448-574) ----------------------------------------------------------------------
    def get_state(self, f, dtype=np.float64, out=None):
        """(n, 24) [x v F C].  dtype float32 (or a float32 ``out``): the fp32 rows as stored, no conversion; a device ``out``
        (``__cuda_array_interface__``) is filled on the simulator's stream without a host copy."""
        if out is not None and _dev(out) is not None:
            ptr, shape = _dev(out)
            assert int(np.prod(shape)) == self.n_total * 24
            check(lib().smx_get_state_dev(self._h, int(f), C.c_void_p(ptr)))
            return out
        if out is not None or np.dtype(dtype) == np.float32:
            out = np.empty((self.n_total, 24), dtype=np.float32) if out is None else out
            assert _f32(out) and out.size == self.n_total * 24
            check(lib().smx_get_state_f32(self._h, int(f), _f_ptr(out)))
            return out
        out = np.zeros((self.n_total, 24))
        check(lib().smx_get_state(self._h, int(f), d_ptr(out)))
        return out

    def pin(self, *arrays):
        """Page-lock caller float32 buffers once (cudaHostRegister) so that reset / get_grad / add_x_grad on them are plain DMA."""
        for a in arrays:
            assert isinstance(a, np.ndarray) and a.flags.c_contiguous
            check(lib().smx_host_register(C.c_void_p(a.ctypes.data), a.nbytes))

    def unpin(self, *arrays):
        for a in arrays:
            check(lib().smx_host_unregister(C.c_void_p(a.ctypes.data)))

    def set_state(self, f, state):
        x, v, F, Cm = [as_d(a) for a in state[:4]]
        n = self.n_total
        assert x.size == 3 * n and v.size == 3 * n and F.size == 9 * n and Cm.size == 9 * n
        check(lib().smx_set_frame(self._h, int(f), d_ptr(x), d_ptr(v), d_ptr(F), d_ptr(Cm)))

    def reset(self, x):
        d = _dev(x)
        if d is not None:               # (n_total, 24) float32 rows already on the device: no host round trip
            assert d[1] == (self.n_total, 24), "device reset takes (n_total, 24) float32 rows"
            check(lib().smx_reset_dev(self._h, C.c_void_p(d[0])))
            self.cur = 0
            return
        if _f32(x) and x.ndim == 2 and x.shape[1] in (self.dim, 24) and x.shape[0] == self.n_total:
            check(lib().smx_reset_f32(self._h, _f_ptr(x), int(x.shape[1])))
            self.cur = 0
            return
        x = as_d(x)
        assert x.ndim == 2 and x.shape[1] in (self.dim, 24)
        if x.shape[0] == self.n_particles and self.n_batch > 1:
            x = np.ascontiguousarray(np.tile(x, (self.n_batch, 1)))     # same initial state for every rollout
        assert x.shape[0] == self.n_total
        check(lib().smx_reset(self._h, d_ptr(x), int(x.shape[1])))
        self.cur = 0

    def get_x(self, f):
        out = np.zeros((self.n_total, self.dim))
        check(lib().smx_get_x(self._h, int(f), d_ptr(out)))
        return out

    def set_x(self, f, x):
        x = as_d(x, (self.n_total, self.dim))
        check(lib().smx_set_x(self._h, int(f), d_ptr(x)))

    def get_v(self, f):
        out = np.zeros((self.n_total, self.dim))
        check(lib().smx_get_v(self._h, int(f), d_ptr(out)))
        return out

    def set_v(self, f, v):
        v = as_d(v, (self.n_total, self.dim))
        check(lib().smx_set_v(self._h, int(f), d_ptr(v)))

    def copyframe(self, source, target):
        check(lib().smx_copy_frame(self._h, int(source), int(target)))

    def get_grad(self, f, dtype=np.float64, out=None):
        """(x.grad[f], v.grad[f]) (mpm_simulator.py:561-574).  dtype float32 / out=(xg, vg) float32 arrays: no conversion pass."""
        if out is not None or np.dtype(dtype) == np.float32:
            xg, vg = out if out is not None else (np.empty((self.n_total, self.dim), dtype=np.float32), np.empty((self.n_total, self.dim), dtype=np.float32))
            assert _f32(xg) and _f32(vg) and xg.size == vg.size == self.n_total * self.dim
            check(lib().smx_get_grad_f32(self._h, int(f), _f_ptr(xg), _f_ptr(vg)))
            return xg, vg
        xg, vg = np.zeros((self.n_total, self.dim)), np.zeros((self.n_total, self.dim))
        check(lib().smx_get_grad(self._h, int(f), d_ptr(xg), d_ptr(vg)))
        return xg, vg

    # -- adjoint seeds (what Taichi losses do by writing x.grad[f] directly) --------------------------------
    def add_x_grad(self, f, g):
        if _f32(g) and g.size == self.n_total * self.dim:
            check(lib().smx_add_x_grad_f32(self._h, int(f), _f_ptr(g)))
            return
        g = as_d(g, (self.n_total, self.dim))
        check(lib().smx_add_x_grad(self._h, int(f), d_ptr(g)))

    def add_state_grad(self, f, g24):
        d = _dev(g24)
        if d is not None:
            assert int(np.prod(d[1])) == self.n_total * 24
            check(lib().smx_add_state_grad_dev(self._h, int(f), C.c_void_p(d[0])))
            return
        if _f32(g24) and g24.size == self.n_total * 24:
            check(lib().smx_add_state_grad_f32(self._h, int(f), _f_ptr(g24)))
            return
        g = as_d(g24, (self.n_total, 24))
        check(lib().smx_add_state_grad(self._h, int(f), d_ptr(g)))

    def get_state_grad(self, f, dtype=np.float64, out=None):
        if out is not None and _dev(out) is not None:
            ptr, shape = _dev(out)
            assert int(np.prod(shape)) == self.n_total * 24
            check(lib().smx_get_state_grad_dev(self._h, int(f), C.c_void_p(ptr)))
            return out
        if out is not None or np.dtype(dtype) == np.float32:
            out = np.empty((self.n_total, 24), dtype=np.float32) if out is None else out
            assert _f32(out) and out.size == self.n_total * 24
            check(lib().smx_get_state_grad_f32(self._h, int(f), _f_ptr(out)))
            return out
        out = np.zeros((self.n_total, 24))
        check(lib().smx_get_state_grad(self._h, int(f), d_ptr(out)))
        return out

    def clear_all_gradients(self):
        """ti.ad.clear_all_gradients() for this simulator (demo_grip.py:135)."""
        check(lib().smx_clear_grads(self._h))

    # -- control (mpm_simulator.py:579-602) ---------------------------------------------------------------------
    def set_action(self, action):
        a = as_d(np.asarray(action, dtype=np.float64)).reshape(-1, self.dim)
        if a.shape[0] == self.n_control and self.n_batch > 1:
            a = np.ascontiguousarray(np.tile(a, (self.n_batch, 1)))
        assert a.shape[0] == self.n_batch * self.n_control
        check(lib().smx_set_action(self._h, d_ptr(a)))

    def set_control_idx(self, idx=None):
        idx = np.asarray(idx)
        if idx.shape[0] == self.n_particles and self.n_batch > 1:
            idx = np.tile(idx, self.n_batch)
        if self.n_control == 0:
            idx = idx * 0
        i = np.ascontiguousarray(idx, dtype=np.int32)
        check(lib().smx_set_control_idx(self._h, i.ctypes.data_as(C.POINTER(C.c_int32))))

    # -- introspection -------------------------------------------------------------------------------------------
    def synchronize(self):
        check(lib().smx_synchronize(self._h))

    def sort_keys(self, f):
        k = np.zeros(self.n_total, dtype=np.uint32)
        check(lib().smx_get_sort_keys(self._h, int(f), k.ctypes.data_as(C.POINTER(C.c_uint32))))
        return k

    def permutation(self, f):
        p = np.zeros(self.n_total, dtype=np.uint32)
        check(lib().smx_get_permutation(self._h, int(f), p.ctypes.data_as(C.POINTER(C.c_uint32))))
        return p

    def get_grid(self):
        G = self.n_grid ** 3
        a, b = np.zeros((G, 4), dtype=np.float32), np.zeros((G, 4), dtype=np.float32)
        check(lib().smx_get_grid(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), b.ctypes.data_as(C.POINTER(C.c_float))))
        return a, b

    def counters(self):
        c = (C.c_int64 * 4)()
        check(lib().smx_get_counters(self._h, c))
        return dict(clamped=c[0], left_active_region=c[1], resorts=c[2], active_blocks=c[3])

    def timer_start(self):
        check(lib().smx_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().smx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def stream_ptr(self):
        """cudaStream_t of the handle (for torch.cuda.ExternalStream ordering with collectives)."""
        p = C.c_void_p()
        check(lib().smx_stream(self._h, C.byref(p)))
        return p.value or 0

    def grad_summary_dev(self, f, out_ptr):
        """16-float gradient summary of the adjoint of frame f into device memory at out_ptr (stream-ordered, no host sync)."""
        check(lib().smx_grad_summary_dev(self._h, int(f), C.c_void_p(int(out_ptr))))

    def profile_step(self, f, n, backward=False):
        """{kernel class: (total device ms, launches)} of smx_step(f, n) / smx_step_grad(f, n): CUDA events on the handle's stream."""
        names, ms, nl, cnt = (C.c_char_p * 32)(), (C.c_float * 32)(), (C.c_int32 * 32)(), C.c_int32(0)
        check(lib().smx_profile_step(self._h, int(f), int(n), int(bool(backward)), names, ms, nl, C.byref(cnt)))
        return {names[i].decode(): (float(ms[i]), int(nl[i])) for i in range(cnt.value)}

    def launch_count(self):
        return int(lib().smx_launch_count(self._h))
