"""TEST INFRASTRUCTURE — the shipped CUDA sources (real_robots_b200/csrc/*.cu) compiled with g++ against the SIMT emulator
of tests/simt/simt_emu.h and driven through the same C ABI with numpy arrays standing in for device memory.

This exists so that `-m "not gpu"` tests can hold the actual kernel source to the oracle on the CPU (shuffle / ballot /
barrier semantics included) before GPU time is spent.  It is not a fallback: the package never imports tests/, the library
is built into tests/simt/_build and has no torch front end.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from real_robots_b200 import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
_CSRC = os.path.join(_ROOT, "real_robots_b200", "csrc")
_SO = os.path.join(_HERE, "_build", "libreal_emu.so")
_SRCS = [os.path.join(_CSRC, f) for f in ("rr_api.cu", "rr_physics.cu", "rr_render.cu", "rr_collide.cu")] + [os.path.join(_HERE, "simt_emu.cpp")]
_DEPS = _SRCS + [os.path.join(_CSRC, "rr_ctx.h"), os.path.join(_ROOT, "include", "real_b200.h"), os.path.join(_HERE, "simt_emu.h")]
_libs = {}


def build(flags=(), force=False):
    """flags: extra -D options (tuning macros of the kernels, e.g. a tiny SOLVE_ENV_CAP to force the solver's global-memory
    path); every flag set gets its own library."""
    so = _SO if not flags else _SO.replace(".so", "_" + "_".join(f.strip("-").replace("=", "") for f in flags) + ".so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in _DEPS):
        # -ffp-contract=off mirrors nvcc -fmad=false (rasteriser, narrowphase); the dynamics kernels are compared to 1e-9
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DRR_SIMT_EMU_BUILD", "-I" + _HERE,
               "-Wno-unknown-pragmas"] + list(flags) + ["-x", "c++"] + _SRCS + ["-o", so]
        subprocess.check_call(cmd)
    return so


def lib(flags=()):
    key = tuple(flags)
    if key not in _libs:
        lb = C.CDLL(build(key))
        for name, (res, args) in L.SYMBOLS.items():
            fn = getattr(lb, name)
            fn.restype = res
            fn.argtypes = args
        assert lb.rr_config_size() == C.sizeof(L.RRConfig)
        _libs[key] = lb
    return _libs[key]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuSim:
    """numpy mirror of real_robots_b200.sim.BatchedSim over the emulated library."""

    def __init__(self, n_envs, n_objects=3, build_flags=(), **cfg):
        self.lib = lib(build_flags)
        self.cfg = L.RRConfig()
        self.lib.rr_config_default(C.byref(self.cfg))
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise KeyError(k)
            setattr(self.cfg, k, v)
        self.n_envs, self.n_objects = int(n_envs), int(n_objects)
        ctx = C.c_void_p()
        rc = self.lib.rr_create(C.byref(ctx), 0, self.n_envs, self.n_objects, L.DEFAULT_BLOB.encode(), C.byref(self.cfg))
        if rc != 0:
            raise RuntimeError("rr_create (emulated) failed: %s" % self.lib.rr_last_error(None).decode())
        self.ctx = ctx

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.rr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.rr_last_error(self.ctx).decode()))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._ck(self.lib.rr_reset(self.ctx, _p(m), None), "rr_reset")

    def set_state(self, q=None, qd=None, obj=None):
        q = None if q is None else np.ascontiguousarray(q, np.float64)
        qd = None if qd is None else np.ascontiguousarray(qd, np.float64)
        obj = None if obj is None else np.ascontiguousarray(obj, np.float64)
        self._ck(self.lib.rr_set_state(self.ctx, _p(q), _p(qd), _p(obj), None), "rr_set_state")

    def get_state(self):
        N = self.n_envs
        q = np.zeros((N, 11)); qd = np.zeros((N, 11)); obj = np.zeros((N, 3, 13))
        self._ck(self.lib.rr_get_state(self.ctx, _p(q), _p(qd), _p(obj), None), "rr_get_state")
        return q, qd, obj

    def set_object_pose(self, obj, pos_quat, mask=None):
        pq = np.ascontiguousarray(pos_quat, np.float64)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._ck(self.lib.rr_set_object_pose(self.ctx, _p(m), int(obj), _p(pq), None), "rr_set_object_pose")

    def step(self, cmd):
        c = np.ascontiguousarray(cmd, np.float64)
        assert c.shape == (self.n_envs, 9)
        self._ck(self.lib.rr_step(self.ctx, _p(c), None), "rr_step")

    def read_state(self):
        N = self.n_envs
        j = np.zeros((N, 9)); t = np.zeros((N, 4)); o = np.zeros((N, self.n_objects, 7))
        self._ck(self.lib.rr_read_state(self.ctx, _p(j), _p(t), _p(o), None), "rr_read_state")
        return j, t, o

    def _imgs(self):
        N = self.n_envs
        return (np.zeros((N, 240, 320, 3), np.uint8), np.zeros((N, 240, 320), np.float32), np.zeros((N, 240, 320), np.int32))

    def render(self, mask=None):
        rgb, depth, seg = self._imgs()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._ck(self.lib.rr_render(self.ctx, _p(m), _p(rgb), _p(depth), _p(seg), None), "rr_render")
        return rgb, depth, seg

    def render_frames(self, frames):
        f = np.ascontiguousarray(frames, np.float64)
        assert f.shape == (self.n_envs, 22, 12)
        rgb, depth, seg = self._imgs()
        self._ck(self.lib.rr_render_frames(self.ctx, _p(f), _p(rgb), _p(depth), _p(seg), None), "rr_render_frames")
        return rgb, depth, seg

    def render_camera(self, view16, proj16):
        v = np.ascontiguousarray(view16, np.float32); p = np.ascontiguousarray(proj16, np.float32)
        rgb, depth, seg = self._imgs()
        self._ck(self.lib.rr_render_camera(self.ctx, _p(v), _p(p), None, _p(rgb), _p(depth), _p(seg), None), "rr_render_camera")
        return rgb, depth, seg

    def shape_frames(self):
        f = np.zeros((self.n_envs, 22, 12))
        self._ck(self.lib.rr_shape_frames(self.ctx, _p(f), None), "rr_shape_frames")
        return f

    def ik(self, pos_quat, max_iter=1000, residual=1e-3):
        pq = np.ascontiguousarray(pos_quat, np.float64); out = np.zeros((self.n_envs, 11))
        self._ck(self.lib.rr_ik(self.ctx, None, _p(pq), _p(out), int(max_iter), float(residual), None), "rr_ik")
        return out

    def diag(self):
        d = np.zeros((self.n_envs, 4), np.int32)
        self._ck(self.lib.rr_diag(self.ctx, _p(d), None), "rr_diag")
        return d

    def contacts(self):
        N = self.n_envs
        hdr = np.zeros((N, 32, 2), np.int32); pts = np.zeros((N, 32, 4, 17)); pairs = np.zeros((96, 2), np.int32); n = C.c_int(0)
        self._ck(self.lib.rr_contacts(self.ctx, _p(hdr), _p(pts), _p(pairs), C.byref(n), None), "rr_contacts")
        return hdr, pts, pairs[:n.value]

    def alloc_host(self):
        N = self.n_envs
        rgb, depth, seg = self._imgs()
        return dict(joint_positions=np.zeros((N, 9)), touch_sensors=np.zeros((N, 4)), object_poses=np.zeros((N, self.n_objects, 7)),
                    retina=rgb, depth=depth, mask=seg)

    def _host_args(self, cmd, mask, out):
        return (self.ctx, _p(cmd), _p(mask), _p(out["joint_positions"]), _p(out["touch_sensors"]), _p(out["object_poses"]),
                _p(out["retina"]), _p(out["depth"]), _p(out["mask"]))

    def step_host(self, cmd, render=True, out=None):
        c = np.ascontiguousarray(cmd, np.float64)
        out = out or self.alloc_host()
        m = np.ones(self.n_envs, np.uint8) if render else None
        self._ck(self.lib.rr_step_host(*self._host_args(c, m, out)), "rr_step_host")
        return out

    def step_host_begin(self, cmd, mask, out):
        self._ck(self.lib.rr_step_host_begin(*self._host_args(cmd, mask, out)), "rr_step_host_begin")

    def step_host_end(self):
        self._ck(self.lib.rr_step_host_end(self.ctx), "rr_step_host_end")
