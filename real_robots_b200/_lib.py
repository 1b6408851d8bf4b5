"""ctypes binding of libreal_b200.so (C ABI: include/real_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is usable, every entry point
raises.  The library is built in-tree by `__graft_entry__.build()` / `make -C real_robots_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RR_LIB", os.path.join(_HERE, "libreal_b200.so"))   # RR_LIB: A/B builds during tuning
DEFAULT_BLOB = os.path.join(_HERE, "assets", "scene.rrb")

NUM_JOINTS, NUM_DOF, NUM_TOUCH, EYE_W, EYE_H, MAX_OBJECTS, NUM_LINKS, NUM_SHAPES = 9, 11, 4, 320, 240, 3, 16, 22


class RRConfig(C.Structure):
    """Mirror of `rr_config` (include/real_b200.h)."""
    _fields_ = [
        ("dt", C.c_double), ("solver_iterations", C.c_int), ("erp", C.c_double), ("erp2", C.c_double),
        ("linear_slop", C.c_double), ("warmstart", C.c_double), ("residual_threshold", C.c_double),
        ("lin_damping", C.c_double), ("ang_damping", C.c_double), ("max_coord_vel", C.c_double),
        ("motor_kp", C.c_double), ("motor_kd", C.c_double), ("motor_max_force", C.c_double),
        ("limit_max_impulse", C.c_double), ("use_box_inertia", C.c_int), ("gravity", C.c_double),
        ("enable_contacts", C.c_int), ("contact_breaking", C.c_double), ("margin", C.c_double),
        ("restitution_threshold", C.c_double), ("friction_erp", C.c_double), ("torsional_friction", C.c_int),
    ]


# every symbol include/real_b200.h declares: name -> (restype, argtypes)
_VP, _I, _D = C.c_void_p, C.c_int, C.c_double
SYMBOLS = {
    "rr_config_default": (None, [C.POINTER(RRConfig)]),
    "rr_config_size": (_I, []),
    "rr_create": (_I, [C.POINTER(_VP), _I, _I, _I, C.c_char_p, C.POINTER(RRConfig)]),
    "rr_destroy": (_I, [_VP]),
    "rr_last_error": (C.c_char_p, [_VP]),
    "rr_n_envs": (_I, [_VP]),
    "rr_n_objects": (_I, [_VP]),
    "rr_reset": (_I, [_VP, _VP, _VP]),
    "rr_set_object_pose": (_I, [_VP, _VP, _I, _VP, _VP]),
    "rr_set_state": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "rr_get_state": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "rr_step": (_I, [_VP, _VP, _VP]),
    "rr_read_state": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "rr_render": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "rr_render_camera": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rr_render_frames": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "rr_shape_frames": (_I, [_VP, _VP, _VP]),
    "rr_ik": (_I, [_VP, _VP, _VP, _VP, _I, _D, _VP]),
    "rr_link_pose": (_I, [_VP, _I, _VP, _VP]),
    "rr_score": (_I, [_VP, _VP, _VP, _VP]),
    "rr_step_host": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rr_step_host_begin": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rr_step_host_end": (_I, [_VP]),
    "rr_launch_count": (C.c_longlong, [_VP]),
    "rr_diag": (_I, [_VP, _VP, _VP]),
    "rr_contacts": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "rr_set_contacts": (_I, [_VP, _VP, _VP, _VP]),
    "rr_profile": (_I, [_VP, _I]),
    "rr_profile_read": (_I, [_VP, _VP, _VP]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load libreal_b200.so; raises NativeLibraryError (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C real_robots_b200/csrc). real_robots_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.rr_config_size() != C.sizeof(RRConfig):
        raise NativeLibraryError("rr_config layout mismatch between libreal_b200.so and _lib.py")
    _lib = lib
    return lib


def default_config():
    cfg = RRConfig()
    load().rr_config_default(C.byref(cfg))
    return cfg


def check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.rr_last_error(ctx)
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
