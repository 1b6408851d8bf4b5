"""BatchedSim — thin PyTorch-tensor front end over the C ABI (include/real_b200.h).

PyTorch is plumbing only: it owns device memory and streams; every computation is a kernel of libreal_b200.so.
Tensors passed in must be CUDA, contiguous, of the documented dtype; outputs are allocated once and reused.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


MAX_MANIFOLDS = 32       # RR_MAX_MANIFOLDS_PER_ENV (include/real_b200.h)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class BatchedSim:
    def __init__(self, n_envs, n_objects=3, device=0, blob=None, **cfg):
        if not torch.cuda.is_available():
            raise L.NativeLibraryError("BatchedSim needs a CUDA device; real_robots_b200 has no CPU fallback")
        self.lib = L.load()
        self.cfg = L.default_config()
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise KeyError("unknown config field %r" % k)
            setattr(self.cfg, k, v)
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.n_envs, self.n_objects = int(n_envs), int(n_objects)
        ctx = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)          # make sure torch's primary context exists first
            rc = self.lib.rr_create(C.byref(ctx), self.device.index, self.n_envs, self.n_objects,
                                    (blob or L.DEFAULT_BLOB).encode(), C.byref(self.cfg))
        if rc != 0:
            raise RuntimeError("rr_create failed: %s" % self.lib.rr_last_error(None).decode())
        self.ctx = ctx
        N = self.n_envs
        dev = self.device
        self.joint_pos = torch.zeros(N, L.NUM_JOINTS, dtype=torch.float64, device=dev)
        self.touch = torch.zeros(N, L.NUM_TOUCH, dtype=torch.float64, device=dev)
        self.obj_pose = torch.zeros(N, self.n_objects, 7, dtype=torch.float64, device=dev)
        self._img = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.rr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc, what):
        L.check(self.lib, self.ctx, rc, what)

    def _chk(self, t, shape, dtype, name):
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError("%s must be a contiguous %s tensor of shape %s on %s" % (name, dtype, tuple(shape), self.device))

    def _mask(self, mask):
        if mask is None:
            return None
        m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        if m.shape != (self.n_envs,):
            raise ValueError("mask must have shape [n_envs]")
        return m

    def images(self):
        if self._img is None:
            N, dev = self.n_envs, self.device
            self._img = (torch.empty(N, L.EYE_H, L.EYE_W, 3, dtype=torch.uint8, device=dev),
                         torch.empty(N, L.EYE_H, L.EYE_W, dtype=torch.float32, device=dev),
                         torch.empty(N, L.EYE_H, L.EYE_W, dtype=torch.int32, device=dev))
        return self._img

    # ------------------------------------------------------------------ API
    def reset(self, mask=None):
        m = self._mask(mask)
        self._ck(self.lib.rr_reset(self.ctx, _ptr(m), self._stream()), "rr_reset")

    def set_object_pose(self, obj, pos_quat, mask=None):
        self._chk(pos_quat, (self.n_envs, 7), torch.float64, "pos_quat")
        m = self._mask(mask)
        self._ck(self.lib.rr_set_object_pose(self.ctx, _ptr(m), int(obj), _ptr(pos_quat), self._stream()), "rr_set_object_pose")

    def set_state(self, q=None, qd=None, obj=None):
        if q is not None:
            self._chk(q, (self.n_envs, 11), torch.float64, "q")
        if qd is not None:
            self._chk(qd, (self.n_envs, 11), torch.float64, "qd")
        if obj is not None:
            self._chk(obj, (self.n_envs, 3, 13), torch.float64, "obj")
        self._ck(self.lib.rr_set_state(self.ctx, _ptr(q), _ptr(qd), _ptr(obj), self._stream()), "rr_set_state")

    def get_state(self):
        N, dev = self.n_envs, self.device
        q = torch.empty(N, 11, dtype=torch.float64, device=dev)
        qd = torch.empty(N, 11, dtype=torch.float64, device=dev)
        obj = torch.empty(N, 3, 13, dtype=torch.float64, device=dev)
        self._ck(self.lib.rr_get_state(self.ctx, _ptr(q), _ptr(qd), _ptr(obj), self._stream()), "rr_get_state")
        return q, qd, obj

    def step(self, joint_cmd):
        self._chk(joint_cmd, (self.n_envs, L.NUM_JOINTS), torch.float64, "joint_cmd")
        self._ck(self.lib.rr_step(self.ctx, _ptr(joint_cmd), self._stream()), "rr_step")

    def read_state(self):
        self._ck(self.lib.rr_read_state(self.ctx, _ptr(self.joint_pos), _ptr(self.touch), _ptr(self.obj_pose), self._stream()),
                 "rr_read_state")
        return self.joint_pos, self.touch, self.obj_pose

    def render(self, mask=None, want_mask=True):
        rgb, depth, seg = self.images()
        m = self._mask(mask)
        self._ck(self.lib.rr_render(self.ctx, _ptr(m), _ptr(rgb), _ptr(depth), _ptr(seg) if want_mask else None, self._stream()),
                 "rr_render")
        return rgb, depth, seg

    def render_camera(self, view16, proj16, mask=None):
        """Render with a caller-supplied camera (column-major float32[16] view / projection, pybullet convention)."""
        v = np.ascontiguousarray(view16, np.float32)
        p = np.ascontiguousarray(proj16, np.float32)
        rgb, depth, seg = self.images()
        m = self._mask(mask)
        self._ck(self.lib.rr_render_camera(self.ctx, v.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), _ptr(m),
                                           _ptr(rgb), _ptr(depth), _ptr(seg), self._stream()), "rr_render_camera")
        return rgb, depth, seg

    def render_frames(self, frames):
        self._chk(frames, (self.n_envs, L.NUM_SHAPES, 12), torch.float64, "frames")
        rgb, depth, seg = self.images()
        self._ck(self.lib.rr_render_frames(self.ctx, _ptr(frames), _ptr(rgb), _ptr(depth), _ptr(seg), self._stream()), "rr_render_frames")
        return rgb, depth, seg

    def shape_frames(self):
        f = torch.empty(self.n_envs, L.NUM_SHAPES, 12, dtype=torch.float64, device=self.device)
        self._ck(self.lib.rr_shape_frames(self.ctx, _ptr(f), self._stream()), "rr_shape_frames")
        return f

    def link_pose(self, link):
        out = torch.empty(self.n_envs, 7, dtype=torch.float64, device=self.device)
        self._ck(self.lib.rr_link_pose(self.ctx, int(link), _ptr(out), self._stream()), "rr_link_pose")
        return out

    def score(self, goal_xyz):
        self._chk(goal_xyz, (self.n_envs, self.n_objects, 3), torch.float64, "goal_xyz")
        out = torch.empty(self.n_envs, dtype=torch.float64, device=self.device)
        self._ck(self.lib.rr_score(self.ctx, _ptr(goal_xyz), _ptr(out), self._stream()), "rr_score")
        return out

    def ik(self, pos_quat, max_iter=1000, residual=1e-3, mask=None):
        self._chk(pos_quat, (self.n_envs, 7), torch.float64, "pos_quat")
        out = torch.empty(self.n_envs, 11, dtype=torch.float64, device=self.device)
        m = self._mask(mask)
        self._ck(self.lib.rr_ik(self.ctx, _ptr(m), _ptr(pos_quat), _ptr(out), int(max_iter), float(residual), self._stream()), "rr_ik")
        return out

    def contacts(self):
        """-> (hdr i32[N,32,2], pts f64[N,32,4,17], pairs i32[n_pairs,2]) — cached manifolds of the last step."""
        N, dev = self.n_envs, self.device
        hdr = torch.empty(N, MAX_MANIFOLDS, 2, dtype=torch.int32, device=dev)
        pts = torch.empty(N, MAX_MANIFOLDS, 4, 17, dtype=torch.float64, device=dev)
        pairs = np.zeros((96, 2), np.int32)
        n = C.c_int(0)
        self._ck(self.lib.rr_contacts(self.ctx, _ptr(hdr), _ptr(pts), pairs.ctypes.data_as(C.c_void_p), C.byref(n), self._stream()), "rr_contacts")
        return hdr, pts, pairs[:n.value]

    def set_contacts(self, hdr, pts):
        """Write the persistent manifolds back (layout of contacts()); call after set_state, which clears them."""
        self._chk(hdr, (self.n_envs, MAX_MANIFOLDS, 2), torch.int32, "hdr")
        self._chk(pts, (self.n_envs, MAX_MANIFOLDS, 4, 17), torch.float64, "pts")
        self._ck(self.lib.rr_set_contacts(self.ctx, _ptr(hdr), _ptr(pts), self._stream()), "rr_set_contacts")

    # checkpoint / resume of the simulation state (not in the reference): (q, qd, obj) and the contact
    # manifolds with their warm-start impulses are everything a step depends on, so a resumed run continues bit for bit
    def state_dict(self):
        q, qd, obj = self.get_state()
        hdr, pts, _ = self.contacts()
        return {"n_envs": self.n_envs, "n_objects": self.n_objects, "q": q.cpu(), "qd": qd.cpu(), "obj": obj.cpu(),
                "contact_hdr": hdr.cpu(), "contact_pts": pts.cpu()}

    def load_state_dict(self, sd):
        if int(sd["n_envs"]) != self.n_envs or int(sd["n_objects"]) != self.n_objects:
            raise ValueError("state_dict is for %d envs / %d objects, this simulator has %d / %d"
                             % (sd["n_envs"], sd["n_objects"], self.n_envs, self.n_objects))
        dev = self.device
        self.set_state(sd["q"].to(dev).contiguous(), sd["qd"].to(dev).contiguous(), sd["obj"].to(dev).contiguous())
        self.set_contacts(sd["contact_hdr"].to(dev).contiguous(), sd["contact_pts"].to(dev).contiguous())

    def diag(self):
        out = torch.empty(self.n_envs, 4, dtype=torch.int32, device=self.device)
        self._ck(self.lib.rr_diag(self.ctx, _ptr(out), self._stream()), "rr_diag")
        return out

    def faults(self):
        """i32[N]: bit 0 = a non-finite joint command arrived in the last step (it acted as a full +-max step; the reference
        asserts instead, robot.py:189), bit 1 = the state had become non-finite and the env was re-initialised."""
        return (self.diag()[:, 3] >> 2) & 3

    def profile(self, enable=True):
        self._ck(self.lib.rr_profile(self.ctx, 1 if enable else 0), "rr_profile")

    def profile_read(self):
        """-> dict kind -> (total ms, launches) for kinds step / render_setup / raster / collide."""
        ms = (C.c_double * 4)()
        n = (C.c_longlong * 4)()
        self._ck(self.lib.rr_profile_read(self.ctx, ms, n), "rr_profile_read")
        names = ["step", "render_setup", "raster", "collide"]
        return {names[i]: (ms[i], n[i]) for i in range(4)}

    def launch_count(self):
        return int(self.lib.rr_launch_count(self.ctx))

    # host-buffer path (numpy in / numpy out), synchronous
    def step_host(self, joint_cmd, render_mask=None, want_mask=True, out=None):
        N = self.n_envs
        cmd = np.ascontiguousarray(joint_cmd, dtype=np.float64)
        assert cmd.shape == (N, L.NUM_JOINTS)
        if out is None:
            out = self.alloc_host(render_mask is not None, want_mask)
        rm = None if render_mask is None else np.ascontiguousarray(render_mask, dtype=np.uint8)

        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)
        rc = self.lib.rr_step_host(self.ctx, p(cmd), p(rm), p(out["joint_positions"]), p(out["touch_sensors"]), p(out["object_poses"]),
                                   p(out.get("retina")), p(out.get("depth")), p(out.get("mask")))
        self._ck(rc, "rr_step_host")
        return out

    def step_host_begin(self, joint_cmd, render_mask, out):
        """Enqueue one step (rr_step_host_begin); `joint_cmd`, `render_mask` and the arrays of `out` (alloc_host) must be
        pinned, contiguous and stay untouched until the matching step_host_end().  At most two steps in flight."""
        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)
        rc = self.lib.rr_step_host_begin(self.ctx, p(joint_cmd), p(render_mask), p(out["joint_positions"]), p(out["touch_sensors"]),
                                         p(out["object_poses"]), p(out.get("retina")), p(out.get("depth")), p(out.get("mask")))
        self._ck(rc, "rr_step_host_begin")

    def step_host_end(self):
        self._ck(self.lib.rr_step_host_end(self.ctx), "rr_step_host_end")

    def alloc_host(self, images=True, want_mask=True, pinned=True):
        N = self.n_envs

        tensors = {}

        def mk(shape, dtype):
            t = torch.empty(shape, dtype=dtype, pin_memory=pinned)
            tensors[len(tensors)] = t
            return t.numpy()
        out = dict(joint_positions=mk((N, 9), torch.float64), touch_sensors=mk((N, 4), torch.float64),
                   object_poses=mk((N, self.n_objects, 7), torch.float64))
        if images:
            out["retina"] = mk((N, L.EYE_H, L.EYE_W, 3), torch.uint8)
            out["depth"] = mk((N, L.EYE_H, L.EYE_W), torch.float32)
            if want_mask:
                out["mask"] = mk((N, L.EYE_H, L.EYE_W), torch.int32)
        out["_torch"] = list(tensors.values())        # the (pinned) torch tensors behind the numpy views, in creation order
        return out
