"""TEST INFRASTRUCTURE — ctypes wrapper around the CPU oracle (oracle/liborc.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (real_robots_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
DEFAULT_BLOB = os.path.join(_HERE, "..", "real_robots_b200", "assets", "scene.rrb")

CONFIG_FIELDS = [
    ("dt", C.c_double), ("solver_iterations", C.c_int), ("erp", C.c_double), ("erp2", C.c_double),
    ("linear_slop", C.c_double), ("warmstart", C.c_double), ("residual_threshold", C.c_double),
    ("lin_damping", C.c_double), ("ang_damping", C.c_double), ("max_coord_vel", C.c_double),
    ("motor_kp", C.c_double), ("motor_kd", C.c_double), ("motor_max_force", C.c_double),
    ("limit_max_impulse", C.c_double), ("use_box_inertia", C.c_int), ("gravity", C.c_double),
    ("enable_contacts", C.c_int), ("contact_breaking", C.c_double), ("margin", C.c_double),
    ("restitution_threshold", C.c_double), ("friction_erp", C.c_double), ("torsional_friction", C.c_int),
]


class OrcConfig(C.Structure):
    _fields_ = CONFIG_FIELDS


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        for name in ("orc_destroy", "orc_reset", "orc_step_sim_only"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        assert L.orc_config_size() == C.sizeof(OrcConfig), "OrcConfig layout mismatch"
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_config():
    c = OrcConfig()
    lib().orc_config_default(C.byref(c))
    return c


class Oracle:
    W, H = 320, 240

    def __init__(self, n_objects=3, blob=DEFAULT_BLOB, **cfg):
        self.L = lib()
        self.cfg = default_config()
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise KeyError(k)
            setattr(self.cfg, k, v)
        self.h = self.L.orc_create(os.path.abspath(blob).encode(), n_objects, C.byref(self.cfg))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.n_objects = n_objects

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(C.c_void_p(self.h))
            self.h = None

    def reset(self):
        self.L.orc_reset(C.c_void_p(self.h))

    def get_state(self):
        q = np.zeros(11); qd = np.zeros(11); obj = np.zeros((3, 13))
        self.L.orc_get_state(C.c_void_p(self.h), _p(q), _p(qd), _p(obj))
        return q, qd, obj

    def set_state(self, q=None, qd=None, obj=None):
        q = None if q is None else np.ascontiguousarray(q, np.float64)
        qd = None if qd is None else np.ascontiguousarray(qd, np.float64)
        obj = None if obj is None else np.ascontiguousarray(obj, np.float64)
        self.L.orc_set_state(C.c_void_p(self.h), None if q is None else _p(q), None if qd is None else _p(qd),
                             None if obj is None else _p(obj))

    def set_targets(self, t11):
        t = np.ascontiguousarray(t11, np.float64)
        self.L.orc_set_targets(C.c_void_p(self.h), _p(t))

    def step(self, cmd9):
        c = np.ascontiguousarray(cmd9, np.float64)
        assert c.shape == (9,)
        self.L.orc_step(C.c_void_p(self.h), _p(c))

    def step_sim_only(self):
        self.L.orc_step_sim_only(C.c_void_p(self.h))

    def read(self):
        j = np.zeros(9); t = np.zeros(4)
        self.L.orc_read(C.c_void_p(self.h), _p(j), _p(t))
        return j, t

    def link_pose(self, link):
        o = np.zeros(7)
        self.L.orc_link_pose(C.c_void_p(self.h), int(link), _p(o))
        return o

    def aba(self, tau):
        t = np.ascontiguousarray(tau, np.float64); o = np.zeros(11)
        self.L.orc_aba_qdd(C.c_void_p(self.h), _p(t), _p(o))
        return o

    def minv(self):
        o = np.zeros((11, 11))
        self.L.orc_minv(C.c_void_p(self.h), _p(o))
        return o

    def diag(self):
        r = C.c_int(); i = C.c_int(); c = C.c_int(); res = C.c_double()
        self.L.orc_diag(C.c_void_p(self.h), C.byref(r), C.byref(i), C.byref(c), C.byref(res))
        return dict(rows=r.value, iterations=i.value, contacts=c.value, residual=res.value)

    def diag2(self):
        a = C.c_int(); b = C.c_int(); c = C.c_int()
        self.L.orc_diag2(C.c_void_p(self.h), C.byref(a), C.byref(b), C.byref(c))
        return dict(overlaps=a.value, manifolds=b.value, manifold_overflow=c.value)

    def contacts(self, max_pts=256):
        a = np.zeros(max_pts, np.int32); b = np.zeros(max_pts, np.int32); d = np.zeros((max_pts, 11))
        n = self.L.orc_get_contacts(C.c_void_p(self.h), max_pts, _p(a), _p(b), _p(d))
        return a[:n], b[:n], d[:n]

    def render(self):
        rgb = np.zeros((self.H, self.W, 3), np.uint8); depth = np.zeros((self.H, self.W), np.float32)
        mask = np.zeros((self.H, self.W), np.int32)
        self.L.orc_render(C.c_void_p(self.h), _p(rgb), _p(depth), _p(mask))
        return rgb, depth, mask

    def shape_frames(self):
        f = np.zeros((22, 12))
        self.L.orc_shape_frames(C.c_void_p(self.h), _p(f))
        return f

    def render_frames(self, frames, view16=None, proj16=None):
        if view16 is None:
            view16, proj16 = eye_matrices()
        frames = np.ascontiguousarray(frames, np.float64)
        view16 = np.ascontiguousarray(view16, np.float32); proj16 = np.ascontiguousarray(proj16, np.float32)
        rgb = np.zeros((self.H, self.W, 3), np.uint8); depth = np.zeros((self.H, self.W), np.float32)
        mask = np.zeros((self.H, self.W), np.int32)
        self.L.orc_render_frames(C.c_void_p(self.h), _p(frames), _p(view16), _p(proj16), _p(rgb), _p(depth), _p(mask))
        return rgb, depth, mask

    def targets(self):
        """Motor targets (11 dofs) set by the action prologue of the last step."""
        t = np.zeros(11)
        self.L.orc_get_targets(C.c_void_p(self.h), _p(t))
        return t

    def ik(self, pos, quat, max_iter=1000, residual=1e-3):
        p = np.ascontiguousarray(pos, np.float64); q = np.ascontiguousarray(quat, np.float64); o = np.zeros(11)
        rc = self.L.orc_ik(C.c_void_p(self.h), _p(p), _p(q), int(max_iter), C.c_double(residual), _p(o))
        return o, rc


def eye_matrices():
    v = np.zeros(16, np.float32); p = np.zeros(16, np.float32)
    lib().orc_eye_matrices(_p(v), _p(p))
    return v, p
