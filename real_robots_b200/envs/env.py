"""Host-side stand-in for real_robots/envs/env.py: Goal, REALRobotEnv (the N = 1 gym facade) and the eye camera record.

The public surface (method and attribute names, argument meaning, return types, error behaviour) is the reference's
(env.py:15-467) and is checked against vectors recorded from the reference's own code (tests/test_reference_golden.py).
Nothing below the reference's `self._p.*` seam is Python here: one `rr_step` call does the whole
limitActionByJoint / control_objects_limits / apply_action / stepSimulation sequence, `rr_render` the retina, `rr_ik` the
inverse kinematics (include/real_b200.h, through BatchedSim).  The facade is an n_envs = 1 view that copies results to numpy
to honour the reference's return types; the zero-copy batched interface is envs/vec_env.py.
"""
import os

import numpy as np

from .. import spaces
from ..util import euler_to_quat
from .robot import Kuka

# ---- constants of the reference (file:line in the comments)
CARTESIAN_BOUNDS = ([-0.25, -0.5, 0.40, -1, -1, -1, -1], [0.25, 0.5, 0.60, 1, 1, 1, 1])      # env.py:44-47 xyz + quaternion
MACRO_BOUNDS = ([[-0.25, -0.5], [-0.25, -0.5]], [[0.05, 0.5], [0.05, 0.5]])                   # env.py:49-52 two table points
PLAN_ORIENTATION_EULER = [0, 3.14, -1.57]                                                       # env.py:421 gripper pointing down
PLAN_HIGH_Z, PLAN_LOW_Z, PLAN_SPACING = 0.6, 0.46, 0.05                                         # env.py:425-444
SCORE_DECAY = -np.log(0.25) / 0.10                                                              # env.py:192: 0.25 at 10 cm
# action type -> (command keys of the action dict, name of the step method, request-cache attributes reset to None)
ACTION_TYPES = {
    'joints': (("joint_command",), "step_joints", ()),
    'cartesian': (("cartesian_command", "gripper_command"), "step_cartesian", ("requested_coords", "requested_orient", "last_ik")),
    'macro_action': (("macro_action",), "step_macro", ("requested_action",)),
}


def DefaultRewardFunc(observation):     # env.py:11-12: the task has no extrinsic reward
    return 0


class Goal:
    """env.py:15-24; instances are pickled inside goal files, hence the module path set below."""

    def __init__(self, initial_state=None, final_state=None, retina=None,
                 retina_before=None, challenge=None, mask=None):
        self.initial_state, self.final_state = initial_state, final_state
        self.retina, self.retina_before = retina, retina_before
        self.challenge, self.mask = challenge, mask


Goal.__module__ = "real_robots.envs.env"     # goal files stay loadable by the reference (generate_goals.py:435, env.py:143-145)


def macro_plan(goToPosXY, point_1, point_2):
    """The 1000 x 9 joint plan of a macro action (env.py:410-459): home2 100 rows, above p1 100, down at p1 50, a 500-row
    sweep to p2 at table height, above p2 50, home2 100, home 100.  goToPosXY(xyz) -> 9 joint values.

    The sweep visits `pieces` = min(int(|p2 - p1| / 0.05) + 1, 500) equally spaced way-points; way-point i owns the rows from
    i * (500 // pieces) on, the last one to the end.  The reference also solves the IK of the low point above p2 and never uses
    it; that solve is skipped here (one IK call fewer per plan, same plan)."""
    p1, p2 = np.asarray(point_1, dtype=np.float64), np.asarray(point_2, dtype=np.float64)
    at = lambda p, z: np.append(p, z)                                                           # noqa: E731
    home = np.zeros(9)
    home2 = np.zeros(9)
    home2[[5, 6]] = np.pi / 2
    above_1, down_1, above_2 = goToPosXY(at(p1, PLAN_HIGH_Z)), goToPosXY(at(p1, PLAN_LOW_Z)), goToPosXY(at(p2, PLAN_HIGH_Z))
    n_rows = 500
    a, b = at(p1, PLAN_LOW_Z), at(p2, PLAN_LOW_Z)
    pieces = min(int(np.linalg.norm(b - a) / PLAN_SPACING) + 1, n_rows)
    solved = np.array([goToPosXY(w) for w in np.linspace(a, b, pieces + 1)[1:]])
    owner = np.minimum(np.arange(n_rows) // (n_rows // pieces), pieces - 1)
    blocks = [(home2, 100), (above_1, 100), (down_1, 50), (None, n_rows), (above_2, 50), (home2, 100), (home, 100)]
    return np.vstack([solved[owner] if row is None else np.tile(row, (count, 1)) for row, count in blocks])


class REALRobotEnv:
    """ Create a REALCompetion environment (drop-in for real_robots.envs.env.REALRobotEnv). """

    intrinsic_timesteps = int(15e6)     # env.py:32-34
    extrinsic_timesteps = int(10e3)
    extrinsic_trials = int(50)
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, render=False, objects=3, action_type='joints', additional_obs=True, device=0):
        if action_type not in ACTION_TYPES:
            raise ValueError("action_type must be one 'joints', 'cartesian' or 'macro_action'")
        self.robot = Kuka(additional_obs, objects)
        self.isRender, self.additional_obs, self.action_type = render, additional_obs, action_type
        self._device, self._sim = device, None
        # spaces (env.py:42-81)
        box = lambda bounds: spaces.Box(low=np.array(bounds[0]), high=np.array(bounds[1]), dtype=float)     # noqa: E731
        self.joints_space = self.robot.action_space
        self.cartesian_space, self.macro_space = box(CARTESIAN_BOUNDS), box(MACRO_BOUNDS)
        self.gripper_space = spaces.Box(low=0, high=np.pi / 2, shape=(2,), dtype=float)
        by_key = {"joint_command": self.joints_space, "cartesian_command": self.cartesian_space,
                  "gripper_command": self.gripper_space, "macro_action": self.macro_space}
        keys, step_name, cache_attrs = ACTION_TYPES[action_type]
        self.action_space = spaces.Dict(dict([(k, by_key[k]) for k in keys] + [("render", spaces.MultiBinary(1))]))
        self.observation_space = self.robot.observation_space
        self.step = getattr(self, step_name)
        for attr in cache_attrs:
            setattr(self, attr, None)
        # debug camera of render('rgb_array') (env.py:83-90) and the eye (env.py:136)
        self._cam_dist, self._cam_yaw, self._cam_roll, self._cam_pitch, self._cam_pos = 1.2, 30, 0, -30, [0, 0, .4]
        self._render_width, self._render_height = 320, 240
        self.eyes = {"eye": EyeCamera([0.01, 0, 1.2], [0, 0, 0])}
        self.setCamera()                      # env.py:90: envCamera exists from construction on
        self.reward_func = DefaultRewardFunc
        # goals (env.py:97-113) and the zero images returned while the camera is off (env.py:114-122)
        obs = self.observation_space.spaces
        K = self.robot.ObsSpaces
        blank = lambda key: obs[key].sample() * 0                                                             # noqa: E731
        self.goal = Goal(retina=blank(K.GOAL))
        self.goals, self.goal_idx = None, -1
        from .. import getPackageDataPath
        self.goals_dataset_path = os.path.join(getPackageDataPath(), "goals_dataset.npy.npz")
        self.no_retina, self.no_depth = blank(K.RETINA), blank(K.DEPTH)
        if additional_obs:
            self.no_mask = blank(K.MASK)
            self.get_observation = self.get_observation_extended
        self.timestep = 0

    # ------------------------------------------------------------------ goals (env.py:143-200)
    def set_goals_dataset_path(self, path):
        assert os.path.exists(path), "Non existent path {}".format(path)
        self.goals_dataset_path = path

    def load_goals(self):
        """The file holds one pickled object array of Goal (np.savez_compressed(path, list_of_goals))."""
        archive = np.load(self.goals_dataset_path, allow_pickle=True)
        self.goals = archive[archive.files[0]]

    def set_goal(self):
        """Next goal of the dataset: objects go to its initial poses; its final state keeps positions only."""
        if self.goals is None:
            self.load_goals()
        self.goal_idx += 1
        goal = self.goal = self.goals[self.goal_idx]
        for name, pose in goal.initial_state.items():
            self.robot.object_bodies[name].reset_pose(pose[:3], pose[3:])
        goal.final_state = {name: pose[:3] for name, pose in goal.final_state.items()}
        return self.get_observation()

    def evaluateGoal(self):
        """(challenge, sum over goal objects of exp(-ln 4 / 0.1 m x distance to the goal position))."""
        bodies = self.robot.object_bodies                     # KeyError for unknown objects, as in the reference
        score = 0
        for name, target in self.goal.final_state.items():
            offset = np.array(target[:3]) - np.array(bodies[name].get_position())
            score += np.exp(-SCORE_DECAY * np.linalg.norm(offset))
        return self.goal.challenge, score

    def extrinsicFormula(self, p_goal, p, a_goal, a, w=1):
        """Unused by the reference's own scoring (env.py:168-179) but part of its surface: weighted position / orientation
        closeness, 0.25 at 5 cm and at a quaternion distance of 0.3."""
        near = lambda dist, scale: np.exp(np.log(0.25) / scale * dist)                                        # noqa: E731
        orient_dist = min(np.linalg.norm(a_goal - a), np.linalg.norm(a_goal + a))
        return w * near(np.linalg.norm(p_goal - p), 0.05) + (1 - w) * near(orient_dist, 0.30)

    # The next two run on the device inside every rr_step; the host versions exist because they are public methods of the
    # reference (a policy can preview the rate-limited action, a script can force the object check between steps).
    def limitActionByJoint(self, desired_joints):
        """env.py:314-321: the request clipped to +-[.2 .2 .2 .2 .2 .3 .3 .1 .1] rad around the current joints."""
        now = np.asarray(self.robot.calc_state(), dtype=np.float64)
        reach = np.array([0.2] * 5 + [0.3] * 2 + [0.1] * 2)
        return now + np.clip(np.asarray(desired_joints, dtype=np.float64) - now, -reach, reach)

    def control_objects_limits(self):
        """env.py:257-264: objects below the table top level, or beyond the shelf front and below it, go back to their start."""
        floor = self.robot.object_poses['table'][2]
        for name in self.robot.used_objects:
            x, _, z = self.robot.object_bodies[name].get_position()
            if z < floor or (x > 0.11 and z < 0.29):
                self.robot.reset_object(name)

    def set_eye(self, name, eye_pos=(0.01, 0, 1.2), target_pos=(0, 0, 0)):     # env.py:124-141
        self.eyes[name] = EyeCamera(list(eye_pos), list(target_pos))

    def setCamera(self):                                                         # env.py:124-134
        from ..videomaker import EnvCamera
        self.envCamera = EnvCamera(distance=self._cam_dist, yaw=self._cam_yaw, pitch=self._cam_pitch, roll=self._cam_roll,
                                   pos=self._cam_pos, width=self._render_width, height=self._render_height)

    # ------------------------------------------------------------------ lifecycle
    def _ensure_sim(self):
        if self._sim is None:
            from ..sim import BatchedSim
            self._sim = BatchedSim(1, len(self.robot.used_objects) - 1, device=self._device)
            self.robot._bind(self._sim)
            import torch
            self._torch = torch
            self._cmd = torch.zeros(1, 9, dtype=torch.float64, device=self._sim.device)
            # step_joints goes through rr_step_host: one C call and one synchronisation per step, pinned host buffers
            self._host = self._sim.alloc_host(images=True, want_mask=True, pinned=True)
            self._host_cmd = torch.zeros(1, 9, dtype=torch.float64).pin_memory().numpy()
            self._host_on = torch.ones(1, dtype=torch.uint8).pin_memory().numpy()

    def reset(self):                      # env.py:206-219
        self._ensure_sim()
        self._sim.reset()
        self.timestep = 0
        return self.get_observation()

    def render(self, mode='human', close=False):     # env.py:221-228
        if mode == "human":
            self.isRender = True
        if mode != "rgb_array":
            return np.array([])
        # EnvCamera.render (env.py:470-513): yaw-pitch-roll debug camera.  The reference asks pybullet for the OpenGL
        # renderer here; this build serves the same camera from the TinyRenderer-convention rasteriser (no OpenGL).
        from .. import cameras
        self._ensure_sim()
        view = cameras.view_from_yaw_pitch_roll(self._cam_pos, self._cam_dist, self._cam_yaw, self._cam_pitch, self._cam_roll, 2)
        proj = cameras.projection_fov(80, float(self._render_width) / self._render_height, 0.1, 100.0)
        rgb, _, _ = self._sim.render_camera(view, proj)
        return rgb[0].cpu().numpy()

    def close(self):
        if self._sim is not None:
            self._sim.close()
            self._sim = None

    # ------------------------------------------------------------------ accessors (env.py:230-255)
    def get_part_pos(self, name):
        return self.robot.parts[name].get_position()

    def get_obj_pos(self, name):
        return self.robot.object_bodies[name].get_position()

    def get_obj_pose(self, name):
        return self.robot.object_bodies[name].get_pose()

    def get_all_used_objects(self):
        bodies = self.robot.object_bodies
        return {name: bodies[name].get_position() for name in self.robot.used_objects[1:]}      # [0] is the table

    def get_contacts(self):
        return self.robot.get_contacts()

    def get_retina(self):
        rgb, depth, seg = self._sim.render()
        return rgb[0].cpu().numpy(), seg[0].cpu().numpy(), depth[0].cpu().numpy()

    # ------------------------------------------------------------------ observations (env.py:266-312)
    def _images(self, camera_on):
        if camera_on:
            return self.get_retina()
        return self.no_retina, getattr(self, "no_mask", None), self.no_depth

    def get_observation(self, camera_on=True):
        K = Kuka.ObsSpaces
        retina, _, depth = self._images(camera_on)
        return {K.JOINT_POSITIONS: self.robot.calc_state(), K.TOUCH_SENSORS: self.robot.get_touch_sensors(),
                K.RETINA: retina, K.DEPTH: depth, K.GOAL: self.goal.retina}

    def get_observation_extended(self, camera_on=True):
        K = Kuka.ObsSpaces
        retina, mask, depth = self._images(camera_on)
        return {K.JOINT_POSITIONS: self.robot.calc_state(), K.TOUCH_SENSORS: self.robot.get_touch_sensors(),
                K.RETINA: retina, K.DEPTH: depth, K.MASK: mask, K.OBJ_POS: self.get_all_used_objects(),
                K.GOAL: self.goal.retina, K.GOAL_MASK: self.goal.mask, K.GOAL_POS: self.goal.final_state}

    def _stock_observation(self):
        """True while nobody replaced the observation hooks (then step_joints may assemble the dict from rr_step_host's buffers)."""
        cls = type(self)
        own = self.__dict__.get("get_observation")
        bound_ok = own is None or getattr(own, "__func__", None) in (REALRobotEnv.get_observation, REALRobotEnv.get_observation_extended)
        return (bound_ok and cls.get_observation is REALRobotEnv.get_observation and cls.get_observation_extended is REALRobotEnv.get_observation_extended
                and cls.get_retina is REALRobotEnv.get_retina and cls.get_all_used_objects is REALRobotEnv.get_all_used_objects
                and type(self.robot).calc_state is Kuka.calc_state and type(self.robot).get_touch_sensors is Kuka.get_touch_sensors)

    def _observation_from_host(self, camera_on):
        """The dict get_observation / get_observation_extended would build (env.py:266-312), from the buffers rr_step_host filled."""
        K = Kuka.ObsSpaces
        h = self._host
        if camera_on:
            retina, depth, mask = h["retina"][0].copy(), h["depth"][0].copy(), h["mask"][0].copy()
        else:
            retina, depth, mask = self.no_retina, self.no_depth, getattr(self, "no_mask", None)
        obs = {K.JOINT_POSITIONS: [float(v) for v in h["joint_positions"][0]], K.TOUCH_SENSORS: h["touch_sensors"][0].copy(),
               K.RETINA: retina, K.DEPTH: depth}
        if self.additional_obs:
            obs[K.MASK] = mask
            obs[K.OBJ_POS] = {name: h["object_poses"][0, i, :3].copy() for i, name in enumerate(self.robot.used_objects[1:])}
            obs[K.GOAL] = self.goal.retina
            obs[K.GOAL_MASK] = self.goal.mask
            obs[K.GOAL_POS] = self.goal.final_state
        else:
            obs[K.GOAL] = self.goal.retina
        return obs

    # ------------------------------------------------------------------ stepping (env.py:314-467)
    def step(self, action):
        return self.step_joint(action)          # env.py:323-324 (rebound in __init__, exactly as in the reference)

    def _phase_limit(self):
        """Steps allowed in the current phase: intrinsic until the first goal is set, extrinsic afterwards (env.py:345-352)."""
        return self.intrinsic_timesteps if self.goal_idx < 0 else self.extrinsic_timesteps

    def step_joints(self, action):
        command, camera_on = action['joint_command'], action['render']
        command = np.zeros(9) if command is None else np.asarray(command, dtype=np.float64)
        assert np.isfinite(command).all()                   # robot.py:189
        assert len(command) == self.robot.num_joints        # robot.py:190
        # limitActionByJoint, control_objects_limits, apply_action and stepSimulation all run inside rr_step
        camera_on = bool(np.any(camera_on))
        if self._stock_observation():
            # the whole step — command in, physics, read-out, retina, every observation back — is one rr_step_host call
            self._host_cmd[0, :] = command
            self._sim.step_host(self._host_cmd, self._host_on if camera_on else None, out=self._host)
            observation = self._observation_from_host(camera_on)
        else:       # a subclass / monkeypatch replaced an observation hook: keep calling it
            self._cmd.copy_(self._torch.from_numpy(command.reshape(1, 9)))
            self._sim.step(self._cmd)
            observation = self.get_observation(camera_on)
        self.timestep += 1
        return observation, self.reward_func(observation), bool(self.timestep >= self._phase_limit()), {}

    def _ik(self, coords, orientation):
        t = self._torch
        pq = t.tensor([list(coords) + list(orientation)], dtype=t.float64, device=self._sim.device)
        return self._sim.ik(pq, max_iter=1000, residual=0.001)[0].cpu().numpy()

    def _joint_action(self, joints, action):
        return self.step_joints({"joint_command": joints, "render": action['render']})

    def step_cartesian(self, action):
        """env.py:358-386: the IK is solved again only when the request differs from the previous one (exact comparison)."""
        request = action['cartesian_command']
        if request is None:
            return self._joint_action(np.zeros(9), action)
        coords, orient = request[:3], request[3:]
        if not (np.all(coords == self.requested_coords) and np.all(orient == self.requested_orient)):
            self.last_ik = self._ik(coords, orient)
            self.requested_coords, self.requested_orient = coords, orient
        return self._joint_action(np.hstack([self.last_ik[:7], action['gripper_command']]), action)

    def step_macro(self, action):
        """env.py:388-408: a repeated macro action replays the next row of its plan; a new one, or an exhausted plan, re-plans."""
        request = action['macro_action']
        if request is None:
            return self._joint_action(np.zeros(9), action)
        row = self.next_step() if np.all(request == self.requested_action) else None
        if row is None:
            self.requested_action = request
            self.generate_plan(request)
            row = self.next_step()
        return self._joint_action(row, action)

    def generate_plan(self, macro_action):
        down = euler_to_quat(PLAN_ORIENTATION_EULER)
        self.planned_actions = macro_plan(lambda xyz: self._ik(xyz, down)[:9], macro_action[0], macro_action[1])
        self.plan_step = -1

    def next_step(self):
        self.plan_step += 1
        plan = self.planned_actions
        return plan[self.plan_step, :] if self.plan_step < len(plan) else None


class EyeCamera:
    """env.py:516-567: eye position / target bookkeeping.  The image itself comes from rr_render, whose view
    (eye [0.01,0,1.2] -> table position, up z) and projection (fov 80, 4/3, .1, 100) are fixed to the reference's."""

    def __init__(self, eyePosition, targetPosition, fov=80, width=320, height=240):
        self.eyePosition, self.targetPosition, self.upVector = eyePosition, targetPosition, [0, 0, 1]
        self.fov, self.render_width, self.render_height = fov, width, height
        self.pitch_roll = False

    def renderTarget(self, env):
        """(rgb u8[240,320,3], depth f32[240,320], mask i32[240,320]) of `env` seen from the eye (env.py:536-567)."""
        rgb, mask, depth = env.get_retina()
        return rgb, depth, mask

    render = renderTarget
