"""ctypes binding of libsoftmac_b200.so (C ABI: include/softmac_b200.h).

This is the whole host<->device boundary of the package: plain pointers and sizes, no torch types.
The library is required -- there is no CPU path -- and a missing/unbuildable .so raises ImportError
with the build command.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMX_LIB") or os.path.join(_HERE, "lib", "libsoftmac_b200.so")

SMX_FLAG_DENSE_GRID = 1
SMX_FLAG_NO_SORT = 2
SMX_FLAG_DIRECT_RED = 4
SMX_FLAG_NO_GRID_CKPT = 8
SMX_FLAG_EXTERNAL_STREAM = 16
SMX_FLAG_NO_FUSION = 32
SMX_FLAG_NO_SVD_REC = 64
SMX_FLAG_NO_TMA = 128


class SmxConfig(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int32), ("n_grid", C.c_int32), ("max_steps", C.c_int32),
        ("dt", C.c_double), ("E", C.c_double), ("nu", C.c_double), ("gravity", C.c_double * 3),
        ("ground_friction", C.c_double),
        ("material_model", C.c_int32), ("ptype", C.c_int32), ("collision_type", C.c_int32), ("substeps", C.c_int32),
        ("n_control", C.c_int32), ("rigid_velocity_control", C.c_int32), ("sort_every", C.c_int32),
        ("device", C.c_int32), ("flags", C.c_int32), ("stream", C.c_void_p), ("n_batch", C.c_int32),
    ]


class SmxRigidLinear(C.Structure):
    """smx_rigid_linear of include/softmac_b200.h."""
    _fields_ = [
        ("state_dim", C.c_int32), ("action_dim", C.c_int32), ("max_env_steps", C.c_int32), ("fp32_bridge", C.c_int32),
        ("ext_grad_scale", C.c_double),
        ("As", C.POINTER(C.c_double)), ("Aa", C.POINTER(C.c_double)), ("Aw", C.POINTER(C.c_double)), ("c", C.POINTER(C.c_double)),
        ("body", C.POINTER(C.c_double)), ("init_state", C.POINTER(C.c_double)),
        ("joint", C.POINTER(C.c_int32)), ("enable", C.POINTER(C.c_int32)),
    ]


class SmxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsoftmac_b200 error {code}: {msg}")
        self.code = code


_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
up = C.POINTER(C.c_uint32)
fp = C.POINTER(C.c_float)
vp = C.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/softmac_b200.h one to one
_SIGS = {
    "smx_create": [C.POINTER(SmxConfig), C.POINTER(vp)],
    "smx_destroy": [vp],
    "smx_synchronize": [vp],
    "smx_add_primitive": [vp, dp, dp, ip, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int32],
    "smx_set_primitive_params": [vp, C.c_int32, C.c_double, C.c_double],
    "smx_set_primitive_contact": [vp, C.c_int32, C.c_int32],
    "smx_reset": [vp, dp, C.c_int32],
    "smx_set_frame": [vp, C.c_int32, dp, dp, dp, dp],
    "smx_get_state": [vp, C.c_int32, dp],
    "smx_get_x": [vp, C.c_int32, dp],
    "smx_set_x": [vp, C.c_int32, dp],
    "smx_get_v": [vp, C.c_int32, dp],
    "smx_set_v": [vp, C.c_int32, dp],
    "smx_copy_frame": [vp, C.c_int32, C.c_int32],
    "smx_set_primitive_state": [vp, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_get_primitive_state": [vp, C.c_int32, C.c_int32, dp],
    "smx_get_primitive_state_grad": [vp, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_add_primitive_state_grad": [vp, C.c_int32, C.c_int32, dp],
    "smx_get_ext_f": [vp, C.c_int32, dp],
    "smx_clear_ext_f": [vp, C.c_int32],
    "smx_set_ext_f_grad": [vp, C.c_int32, dp],
    "smx_set_primitive_state_b": [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_get_primitive_state_b": [vp, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_get_primitive_state_grad_b": [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_add_primitive_state_grad_b": [vp, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_get_ext_f_b": [vp, C.c_int32, C.c_int32, dp],
    "smx_clear_ext_f_b": [vp, C.c_int32, C.c_int32],
    "smx_set_ext_f_grad_b": [vp, C.c_int32, C.c_int32, dp],
    "smx_get_ext_f_all": [vp, dp],
    "smx_clear_ext_f_all": [vp],
    "smx_set_ext_f_grads_all": [vp, dp],
    "smx_set_primitive_states_all": [vp, C.c_int32, C.c_int32, dp],
    "smx_get_primitive_state_grads_all": [vp, C.c_int32, C.c_int32, dp],
    "smx_set_primitive_action": [vp, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_get_primitive_action_grad": [vp, C.c_int32, C.c_int32, C.c_int32, dp],
    "smx_reset_primitive": [vp, C.c_int32],
    "smx_rigid_linear_create": [vp, vp],
    "smx_rigid_linear_reset": [vp],
    "smx_rigid_linear_set_actions": [vp, C.c_int32, dp],
    "smx_rigid_linear_step": [vp, C.c_int32],
    "smx_rigid_linear_step_grad": [vp, C.c_int32],
    "smx_rigid_linear_finish": [vp],
    "smx_rigid_linear_get_states": [vp, C.c_int32, dp],
    "smx_rigid_linear_get_action_grads": [vp, C.c_int32, C.c_int32, dp],
    "smx_rigid_linear_get_state_grad": [vp, dp],
    "smx_set_plasticity": [vp, C.c_int32, C.c_double],
    "smx_set_action": [vp, dp],
    "smx_set_control_idx": [vp, ip],
    "smx_get_action_grad": [vp, dp],
    "smx_substep": [vp, C.c_int32],
    "smx_substep_grad": [vp, C.c_int32],
    "smx_substep_begin": [vp, C.c_int32],
    "smx_substep_end": [vp, C.c_int32],
    "smx_substep_mid": [vp, C.c_int32],
    "smx_substep_grad_mid": [vp, C.c_int32],
    "smx_substep_grad_begin": [vp, C.c_int32],
    "smx_substep_grad_end": [vp, C.c_int32],
    "smx_set_slab": [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32],
    "smx_slab_halo_export": [vp, C.POINTER(vp), vp],
    "smx_slab_halo_connect": [vp, C.c_int32, vp, vp],
    "smx_slab_halo_push": [vp, C.c_int32, C.c_int32],
    "smx_slab_halo_add": [vp, C.c_int32, C.c_int32],
    "smx_slab_halo_status": [vp, C.POINTER(C.c_int64)],
    "smx_grid_dev": [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_int64)],
    "smx_stream": [vp, C.POINTER(vp)],
    "smx_step": [vp, C.c_int32, C.c_int32],
    "smx_step_grad": [vp, C.c_int32, C.c_int32],
    "smx_step_graph": [vp, C.c_int32, C.c_int32],
    "smx_step_grad_graph": [vp, C.c_int32, C.c_int32],
    "smx_graph_status": [vp, C.POINTER(C.c_int64)],
    "smx_add_state_grad": [vp, C.c_int32, dp],
    "smx_add_x_grad": [vp, C.c_int32, dp],
    "smx_set_chamfer_target": [vp, dp, C.c_int32],
    "smx_chamfer_loss": [vp, C.c_int32, C.c_double, dp],
    "smx_contact_distance_loss": [vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, dp],
    "smx_get_state_grad": [vp, C.c_int32, dp],
    "smx_host_register": [vp, C.c_uint64],
    "smx_host_unregister": [vp],
    "smx_reset_f32": [vp, fp, C.c_int32],
    "smx_get_state_f32": [vp, C.c_int32, fp],
    "smx_add_x_grad_f32": [vp, C.c_int32, fp],
    "smx_add_state_grad_f32": [vp, C.c_int32, fp],
    "smx_get_state_grad_f32": [vp, C.c_int32, fp],
    "smx_get_grad_f32": [vp, C.c_int32, fp, fp],
    "smx_reset_dev": [vp, vp],
    "smx_get_state_dev": [vp, C.c_int32, vp],
    "smx_get_state_grad_dev": [vp, C.c_int32, vp],
    "smx_add_state_grad_dev": [vp, C.c_int32, vp],
    "smx_get_grad": [vp, C.c_int32, dp, dp],
    "smx_clear_grads": [vp],
    "smx_get_sort_keys": [vp, C.c_int32, up],
    "smx_get_permutation": [vp, C.c_int32, up],
    "smx_get_grid": [vp, fp, fp],
    "smx_get_counters": [vp, C.POINTER(C.c_int64)],
    "smx_frame_component_dev": [vp, C.c_int32, C.c_int32, C.POINTER(vp)],
    "smx_timer_start": [vp],
    "smx_timer_stop": [vp, fp],
    "smx_launch_count": [vp],
    "smx_profile_substep": [vp, C.c_int32, C.c_int32, C.POINTER(C.c_char_p), fp, C.POINTER(C.c_int32)],
    "smx_profile_step": [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_char_p), fp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "smx_grad_summary_dev": [vp, C.c_int32, vp],
    "smx_build_sdf_table": [dp, C.c_int32, ip, C.c_int32, ip, dp, C.c_double, dp, dp, C.c_int32],
    "smx_last_error": [],
}
_RESTYPES = {"smx_last_error": C.c_char_p, "smx_launch_count": C.c_int64}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Load (once) and return the shared library.  Never falls back to a CPU implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and not os.environ.get("SMX_LIB"):
            try:                            # a fresh checkout: compile the CUDA library in-tree (never a CPU substitute)
                from . import build as _build
                _build.build()
            except Exception as e:          # noqa: BLE001
                raise ImportError(
                    f"{LIB_PATH} is missing and could not be built ({e}); build it with `python -m softmac_b200.build` "
                    "(nvcc, sm_100a). softmac_b200 has no CPU fallback.") from e
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing. softmac_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


def check(code):
    if code < 0:
        raise SmxError(code, lib().smx_last_error().decode())
    return code


def as_d(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def d_ptr(a):
    return a.ctypes.data_as(dp)
