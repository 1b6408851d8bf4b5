"""Host-side mirror of real_robots/envs/robot.py (class Kuka) over the batched simulator.

Same names, argument meaning and error behaviour as the reference (robot.py:10-226); the pybullet objects it
held (`parts`, `object_bodies`, `jdict`) become light views onto env 0 of a BatchedSim.
"""
import json

import numpy as np

from .. import _lib as L
from .. import spaces
from ..blob import read_blob

_MANIFEST = None


def manifest():
    global _MANIFEST
    if _MANIFEST is None:
        b = read_blob(L.DEFAULT_BLOB)
        _MANIFEST = json.loads(bytes(b["manifest.json"]).decode())
    return _MANIFEST


class BodyPartView:
    """Stand-in for pybullet_envs.robot_bases.BodyPart: get_position / get_pose / reset_pose."""

    def __init__(self, owner, kind, index, name):
        self._o, self._kind, self._index, self.name = owner, kind, index, name

    def get_pose(self):
        pos, quat = self._o._pose(self._kind, self._index)
        return np.concatenate([pos, quat])

    def get_position(self):
        return self._o._pose(self._kind, self._index)[0]

    def current_position(self):
        return self.get_position()

    def get_orientation(self):
        return self._o._pose(self._kind, self._index)[1]

    def reset_pose(self, position, orientation):
        if self._kind != "object":
            raise RuntimeError("reset_pose is only supported on the free objects")
        self._o._reset_object_pose(self._index, position, orientation)

    def reset_position(self, position):
        self.reset_pose(position, self.get_orientation())


class Kuka:
    used_objects = ["table", "orange", "mustard", "cube", "tomato"]     # robot.py:11-16 (class default; set per instance below)
    # robot.py:19-24
    object_poses = {
        "table":   [0.00000, 0.00000, 0.08000, 0.00000, 0.00000, 0.00000],
        "orange":  [0.20000, -0.15000, 0.45000, 0.00000, 0.00000, 0.00000],
        "mustard": [-0.10000, 0.30000, 0.45000, 1.57080, 3.14159, 0.00000],
        "cube":    [-0.10000, 0.00000, 0.45000, 0.00000, 0.00000, 0.00000],
        "tomato":  [-0.10000, -0.30000, 0.45000, 0.00000, 0.00000, 0.00000]}

    num_joints = 9
    num_kuka_joints = 7
    num_gripper_joints = 2
    num_touch_sensors = 4
    eye_width = 320
    eye_height = 240

    class ObsSpaces:          # robot.py:33-42
        JOINT_POSITIONS = "joint_positions"
        TOUCH_SENSORS = "touch_sensors"
        RETINA = "retina"
        DEPTH = "depth"
        MASK = "mask"
        OBJ_POS = "object_positions"
        GOAL = "goal"
        GOAL_MASK = "goal_mask"
        GOAL_POS = "goal_positions"

    def __init__(self, additional_obs=False, objects=3):
        if objects not in (1, 2, 3):
            raise ValueError("objects must be 1, 2 or 3")
        self.robot_position = [-0.55, 0, -0.04]          # robot.py:46
        self.contact_threshold = 0.1                      # robot.py:47
        self.used_objects = ["table", "cube", "tomato", "mustard"][:objects + 1]     # robot.py:49-50
        self.action_dim = self.num_joints
        # robot.py:58-67
        self.min_joints = np.ones(9) * -np.pi * 0.944
        self.max_joints = np.ones(9) * np.pi * 0.944
        self.min_joints[0] = -np.pi * 0.666
        self.max_joints[0] = np.pi * 0.666
        self.min_joints[1:9:2] = -np.pi * 0.666
        self.max_joints[1:9:2] = np.pi * 0.666
        self.min_joints[6] = -np.pi * 0.972
        self.max_joints[6] = np.pi * 0.972
        self.min_joints[-2:] = 0
        self.max_joints[-2:] = np.pi / 2
        self.action_space = spaces.Box(low=self.min_joints, high=self.max_joints, dtype=float)
        S = self.ObsSpaces
        H, W = Kuka.eye_height, Kuka.eye_width
        d = {
            S.JOINT_POSITIONS: spaces.Box(-np.inf, np.inf, [self.num_joints], dtype=float),
            S.TOUCH_SENSORS: spaces.Box(0, np.inf, [self.num_touch_sensors], dtype=float),
            S.RETINA: spaces.Box(0, 255, [H, W, 3], dtype=np.uint8),
            S.DEPTH: spaces.Box(0, 1, [H, W], dtype=float),
            S.GOAL: spaces.Box(0, 255, [H, W, 3], dtype=np.uint8),
        }
        if additional_obs:                                # robot.py:73-99
            high = np.array([np.finfo(np.float32).max] * 3)
            obj_obs = {o: spaces.Box(-high, high, dtype=float) for o in self.used_objects[1:]}
            d[S.MASK] = spaces.Box(0, 255, [H, W], dtype=np.int32)
            d[S.GOAL_MASK] = spaces.Box(0, 255, [H, W], dtype=np.int32)
            d[S.OBJ_POS] = spaces.Dict(obj_obs)
            d[S.GOAL_POS] = spaces.Dict(obj_obs)
        self.observation_space = spaces.Dict(d)
        self.target = "orange"
        self.object_names = dict()
        self.object_bodies = dict()
        self.robot_parts = {}
        self.parts = {}
        self._sim = None

    # ------------------------------------------------------------------ binding to the simulator (env 0)
    def _bind(self, sim):
        self._sim = sim
        m = manifest()
        self.parts = {}
        for i, name in enumerate(m["bodies"][0]["links"]):        # 17 entries: base + 16 links
            self.parts[name] = BodyPartView(self, "link", i - 1, name)
        self.parts["kuka0"] = self.parts[m["bodies"][0]["links"][0]]
        self.robot_body = self.parts["kuka0"]
        self.object_bodies = {"kuka": self.robot_body}
        self.object_names = {0: "kuka"}
        for bid, name in enumerate(self.used_objects):
            kind, idx = ("table", 0) if name == "table" else ("object", bid - 1)
            self.object_bodies[name] = BodyPartView(self, kind, idx, name)
            self.object_names[bid + 1] = name
        self.robot_parts = {i - 1: n for i, n in enumerate(m["bodies"][0]["links"])}

    def _pose(self, kind, index):
        import torch
        if kind == "link":
            pq = self._sim.link_pose(index)[0].cpu().numpy()
        elif kind == "table":
            return np.array(self.object_poses["table"][:3]), np.array([0.0, 0.0, 0.0, 1.0])
        else:
            pq = self._sim.read_state()[2][0, index].cpu().numpy()
        torch.cuda.current_stream().synchronize()
        return pq[:3].copy(), pq[3:].copy()

    def _reset_object_pose(self, index, position, orientation):
        import torch
        pq = torch.tensor([list(position) + list(orientation)], dtype=torch.float64, device=self._sim.device)
        self._sim.set_object_pose(index, pq.expand(self._sim.n_envs, 7).contiguous())

    def reset_object(self, obj_name):                      # robot.py:125-129
        from ..util import euler_to_quat
        position = self.object_poses[obj_name][:3]
        orientation = euler_to_quat(self.object_poses[obj_name][3:])
        self.object_bodies[obj_name].reset_pose(position, orientation)

    # ------------------------------------------------------------------ observations
    def calc_state(self):                                  # robot.py:203-211 (a python list of 9 floats)
        return [float(v) for v in self._sim.read_state()[0][0].cpu().numpy()]

    def get_touch_sensors(self):                           # robot.py:152-163
        return self._sim.read_state()[1][0].cpu().numpy().copy()

    def get_contacts(self, forces=False):                  # robot.py:131-150
        hdr, pts, pairs = self._sim.contacts()
        hdr = hdr[0].cpu().numpy()
        pts = pts[0].cpu().numpy()
        m = manifest()
        shapes = m["shapes"]
        contact_dict = {}
        for slot in range(hdr.shape[0]):
            pi, n = int(hdr[slot, 0]), int(hdr[slot, 1])
            if pi < 0:
                break
            sa, sb = shapes[int(pairs[pi, 0])], shapes[int(pairs[pi, 1])]
            # contact_list() = getContactPoints(body, -1, link, -1): robot link vs the BASE link of another body
            if sa["body"] != 0 or sb["body"] == 0 or sb["link"] != -1:
                continue
            part_name = sa["name"]
            other = self.object_names.get(sb["body"], m["bodies"][sb["body"]]["name"])
            for k in range(n):
                if abs(pts[slot, k, 9]) < self.contact_threshold:
                    entry = [other, float(pts[slot, k, 10]) / self._sim.cfg.dt] if forces else other
                    contact_dict.setdefault(part_name, []).append(entry)
        return contact_dict
