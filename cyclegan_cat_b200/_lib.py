"""ctypes binding of libcyclegan_b200.so (include/cyclegan_b200.h).

The library is built in-tree by `cyclegan_cat_b200.build.build()` (nvcc, sm_100a).
There is NO fallback: if the shared library is missing or a call fails, the
host code raises.
"""
import ctypes
import os

from .ir import LayerDesc, TrainCfg, VarInfo

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcyclegan_b200.so")

c_int, c_i64, c_size_t, c_void_p, c_float_p = (ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p,
                                               ctypes.POINTER(ctypes.c_float))

# name -> (restype, argtypes); every symbol include/cyclegan_b200.h declares
SIGNATURES = {
    "cg_init": (c_int, [c_int]),
    "cg_last_error": (ctypes.c_char_p, []),
    "cg_version": (c_int, []),
    "cg_abi_sizeof": (c_int, [c_int]),
    "cg_net_create": (c_int, [ctypes.POINTER(LayerDesc), c_int, c_int, ctypes.POINTER(c_void_p)]),
    "cg_net_destroy": (None, [c_void_p]),
    "cg_net_param_floats": (c_int, [c_void_p, ctypes.POINTER(c_i64)]),
    "cg_net_state_floats": (c_int, [c_void_p, ctypes.POINTER(c_i64)]),
    "cg_net_bind_state": (c_int, [c_void_p, c_void_p]),
    "cg_net_set_training": (c_int, [c_void_p, c_int]),
    "cg_net_set_seed": (c_int, [c_void_p, ctypes.c_uint64]),
    "cg_normalize_u8": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cg_postprocess_u8": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cg_resize_bilinear": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "cg_resize_crop_flip": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "cg_net_var_count": (c_int, [c_void_p, ctypes.POINTER(c_int)]),
    "cg_net_var_info": (c_int, [c_void_p, c_int, ctypes.POINTER(VarInfo)]),
    "cg_net_out_shape": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_int * 4)]),
    "cg_net_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "cg_net_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_int, c_int, c_int, c_int, c_void_p]),
    "cg_net_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                c_void_p]),
    "cg_net_fetch_tensor": (c_int, [c_void_p, c_int, c_void_p, ctypes.POINTER(c_int * 4), c_void_p]),
    "cg_trainer_fetch_tensor": (c_int, [c_void_p, c_int, c_int, c_void_p, ctypes.POINTER(c_int * 4), c_void_p]),
    "cg_trainer_create": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(TrainCfg),
                                  ctypes.POINTER(c_void_p)]),
    "cg_trainer_destroy": (None, [c_void_p]),
    "cg_trainer_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "cg_trainer_bind": (c_int, [c_void_p, ctypes.POINTER(c_void_p * 4), ctypes.POINTER(c_void_p * 4),
                                ctypes.POINTER(c_void_p * 4), ctypes.POINTER(c_void_p * 4), c_void_p, c_size_t]),
    "cg_validate_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cg_train_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cg_trainer_compute_gradients": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                             c_void_p]),
    "cg_trainer_apply_gradients": (c_int, [c_void_p, c_void_p]),
    "cg_optimizer_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_i64, c_void_p]),
    "cg_trainer_get_iterations": (c_int, [c_void_p, ctypes.POINTER(c_i64 * 4)]),
    "cg_trainer_set_iterations": (c_int, [c_void_p, ctypes.POINTER(c_i64 * 4)]),
    "cg_trainer_plan_count": (c_int, [c_void_p, ctypes.POINTER(c_i64), ctypes.POINTER(c_int)]),
    "cg_trainer_fetch_image": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "cg_comm_unique_id": (c_int, [ctypes.c_char * 128]),
    "cg_trainer_comm_init": (c_int, [c_void_p, ctypes.c_char * 128, c_int, c_int]),
    "cg_prof_enable": (c_int, [c_int]),
    "cg_prof_read": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_double)]),
    "cg_launch_count": (c_int, [ctypes.POINTER(c_i64), c_int]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). cyclegan_cat_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    from .ir import AdamCfg
    for which, mirror in enumerate((LayerDesc, VarInfo, TrainCfg, AdamCfg)):
        if lib.cg_abi_sizeof(which) != ctypes.sizeof(mirror):
            raise NativeError(f"{LIB_PATH} is stale: sizeof({mirror.__name__}) is {lib.cg_abi_sizeof(which)} in the "
                              f"library, {ctypes.sizeof(mirror)} in the binding -- rebuild it")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cg_last_error()
        raise NativeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


_initialized_device = None


def init_device(device: int = 0):
    """cg_init once per process; raises on a non-sm_100 device or when CUDA is absent."""
    global _initialized_device
    if _initialized_device != device:
        check(load().cg_init(device), "cg_init")
        _initialized_device = device
