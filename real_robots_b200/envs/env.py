"""Host-side mirror of real_robots/envs/env.py: Goal, REALRobotEnv (the N = 1 gym facade) and the cameras.

Same method names, argument meaning and error behaviour as the reference (env.py:15-467); every pybullet call
below the reference's `self._p.*` seam is replaced by a C-ABI call (include/real_b200.h) through BatchedSim.
The facade is an n_envs = 1 view that copies results to numpy to honour the reference's return types; the
zero-copy batched interface is envs/vec_env.py.
"""
import os

import numpy as np

from .. import spaces
from ..util import euler_to_quat
from .robot import Kuka


def DefaultRewardFunc(observation):     # env.py:11-12
    return 0


class Goal:                             # env.py:15-24 (pickled inside goals .npy.npz files)
    def __init__(self, initial_state=None, final_state=None, retina=None,
                 retina_before=None, challenge=None, mask=None):
        self.initial_state = initial_state
        self.final_state = final_state
        self.retina = retina
        self.retina_before = retina_before
        self.challenge = challenge
        self.mask = mask


Goal.__module__ = "real_robots.envs.env"     # goal files stay loadable by the reference (generate_goals.py:435, env.py:143-145)


def macro_plan(goToPosXY, point_1, point_2):
    """env.py:410-459 generate_plan: 1000 x 9 joint plan from two table points; goToPosXY(coords) -> 9 joints."""
    home = np.zeros(9)
    home2 = np.zeros(9)
    home2[5] = np.pi / 2
    home2[6] = np.pi / 2

    def interpolate3D(p1, p2, steps):
        p1 = np.array(p1)
        p2 = np.array(p2)
        dist = np.linalg.norm(p2 - p1)
        pieces = int(dist / 0.05) + 1
        pieces = min(pieces, steps)
        coords = np.linspace(p1, p2, pieces + 1)
        joints = np.zeros((steps, 9))
        chunk = int(steps / pieces)
        for i, coord in enumerate(coords[1:]):
            joints[i * chunk:, :] = goToPosXY(coord)
        return joints

    point_1_h = goToPosXY(np.hstack([point_1, 0.6]))
    point_1_l = goToPosXY(np.hstack([point_1, 0.46]))
    point_2_h = goToPosXY(np.hstack([point_2, 0.6]))
    point_2_l = goToPosXY(np.hstack([point_2, 0.46]))  # noqa: F841 (computed by the reference too; seeds nothing here)
    parts = [np.tile(home2, (100, 1)), np.tile(point_1_h, (100, 1)), np.tile(point_1_l, (50, 1)),
             interpolate3D(np.hstack([point_1, 0.46]), np.hstack([point_2, 0.46]), 500),
             np.tile(point_2_h, (50, 1)), np.tile(home2, (100, 1)), np.tile(home, (100, 1))]
    return np.vstack(parts)


class REALRobotEnv:
    """ Create a REALCompetion environment (drop-in for real_robots.envs.env.REALRobotEnv). """

    intrinsic_timesteps = int(15e6)     # env.py:32-34
    extrinsic_timesteps = int(10e3)
    extrinsic_trials = int(50)
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, render=False, objects=3, action_type='joints', additional_obs=True, device=0):
        self.robot = Kuka(additional_obs, objects)
        self.isRender = render
        self._device = device
        self._sim = None
        self.joints_space = self.robot.action_space
        self.cartesian_space = spaces.Box(low=np.array([-0.25, -0.5, 0.40, -1, -1, -1, -1]),
                                          high=np.array([0.25, 0.5, 0.60, 1, 1, 1, 1]), dtype=float)
        self.macro_space = spaces.Box(low=np.array([[-0.25, -0.5], [-0.25, -0.5]]),
                                      high=np.array([[0.05, 0.5], [0.05, 0.5]]), dtype=float)
        self.gripper_space = spaces.Box(low=0, high=np.pi / 2, shape=(2,), dtype=float)
        if action_type == 'joints':
            self.action_space = spaces.Dict({"joint_command": self.joints_space, "render": spaces.MultiBinary(1)})
            self.step = self.step_joints
        elif action_type == 'cartesian':
            self.action_space = spaces.Dict({"cartesian_command": self.cartesian_space, "gripper_command": self.gripper_space,
                                             "render": spaces.MultiBinary(1)})
            self.step = self.step_cartesian
            self.requested_coords = None
            self.requested_orient = None
            self.last_ik = None
        elif action_type == 'macro_action':
            self.action_space = spaces.Dict({"macro_action": self.macro_space, "render": spaces.MultiBinary(1)})
            self.step = self.step_macro
            self.requested_action = None
        else:
            raise ValueError("action_type must be one 'joints', 'cartesian' or 'macro_action'")
        self.action_type = action_type
        self.observation_space = self.robot.observation_space
        self._cam_dist, self._cam_yaw, self._cam_roll, self._cam_pitch = 1.2, 30, 0, -30
        self._render_width, self._render_height = 320, 240
        self._cam_pos = [0, 0, .4]
        self.eyes = {"eye": EyeCamera([0.01, 0, 1.2], [0, 0, 0])}
        self.reward_func = DefaultRewardFunc
        S = self.robot.ObsSpaces
        self.goal = Goal(retina=self.observation_space.spaces[S.GOAL].sample() * 0)
        from .. import getPackageDataPath
        self.goals_dataset_path = os.path.join(getPackageDataPath(), "goals_dataset.npy.npz")
        self.goals = None
        self.goal_idx = -1
        self.no_retina = self.observation_space.spaces[S.RETINA].sample() * 0
        self.no_depth = self.observation_space.spaces[S.DEPTH].sample() * 0
        self.additional_obs = additional_obs
        if additional_obs:
            self.get_observation = self.get_observation_extended
            self.no_mask = self.observation_space.spaces[S.MASK].sample() * 0
        self.timestep = 0

    # ------------------------------------------------------------------ goals (env.py:143-200)
    def load_goals(self):
        self.goals = list(np.load(self.goals_dataset_path, allow_pickle=True).items())[0][1]

    def set_goals_dataset_path(self, path):
        assert os.path.exists(path), "Non existent path {}".format(path)
        self.goals_dataset_path = path

    def set_goal(self):
        if self.goals is None:
            self.load_goals()
        self.goal_idx += 1
        self.goal = self.goals[self.goal_idx]
        for obj in self.goal.initial_state.keys():
            position = self.goal.initial_state[obj][:3]
            orientation = self.goal.initial_state[obj][3:]
            self.robot.object_bodies[obj].reset_pose(position, orientation)
        for obj in self.goal.final_state.keys():
            self.goal.final_state[obj] = self.goal.final_state[obj][:3]
        return self.get_observation()

    def evaluateGoal(self):
        final_state = self.goal.final_state
        current_state = self.robot.object_bodies
        score = 0
        for obj in final_state.keys():
            p = np.array(current_state[obj].get_position())      # KeyError for unknown objects, as in the reference
            p_goal = np.array(final_state[obj][:3])
            pos_dist = np.linalg.norm(p_goal - p)
            pos_const = -np.log(0.25) / 0.10                      # score goes down to 0.25 within 10 cm
            score += np.exp(-pos_const * pos_dist)
        return self.goal.challenge, score

    # ------------------------------------------------------------------ lifecycle
    def _ensure_sim(self):
        if self._sim is None:
            from ..sim import BatchedSim
            self._sim = BatchedSim(1, len(self.robot.used_objects) - 1, device=self._device)
            self.robot._bind(self._sim)
            import torch
            self._torch = torch
            self._cmd = torch.zeros(1, 9, dtype=torch.float64, device=self._sim.device)

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
        return {obj: self.robot.object_bodies[obj].get_position() for obj in self.robot.used_objects[1:]}

    def get_contacts(self):
        return self.robot.get_contacts()

    def get_retina(self):
        rgb, depth, seg = self._sim.render()
        return rgb[0].cpu().numpy(), seg[0].cpu().numpy(), depth[0].cpu().numpy()

    # ------------------------------------------------------------------ observations (env.py:266-312)
    def get_observation(self, camera_on=True):
        joints = self.robot.calc_state()
        sensors = self.robot.get_touch_sensors()
        if camera_on:
            retina, _, depth = self.get_retina()
        else:
            retina, depth = self.no_retina, self.no_depth
        K = Kuka.ObsSpaces
        return {K.JOINT_POSITIONS: joints, K.TOUCH_SENSORS: sensors, K.RETINA: retina, K.DEPTH: depth, K.GOAL: self.goal.retina}

    def get_observation_extended(self, camera_on=True):
        joints = self.robot.calc_state()
        sensors = self.robot.get_touch_sensors()
        if camera_on:
            retina, mask, depth = self.get_retina()
        else:
            retina, mask, depth = self.no_retina, self.no_mask, self.no_depth
        K = Kuka.ObsSpaces
        return {K.JOINT_POSITIONS: joints, K.TOUCH_SENSORS: sensors, K.RETINA: retina, K.DEPTH: depth, K.MASK: mask,
                K.OBJ_POS: self.get_all_used_objects(), K.GOAL: self.goal.retina, K.GOAL_MASK: self.goal.mask,
                K.GOAL_POS: self.goal.final_state}

    # ------------------------------------------------------------------ stepping (env.py:314-467)
    def step(self, action):
        return self.step_joint(action)          # env.py:323-324 (rebound in __init__, exactly as in the reference)

    def step_joints(self, action):
        joint_action = action['joint_command']
        camera_on = action['render']
        if joint_action is None:
            joint_action = np.zeros(9)
        joint_action = np.asarray(joint_action, dtype=np.float64)
        assert np.isfinite(joint_action).all()           # robot.py:189
        assert len(joint_action) == self.robot.num_joints  # robot.py:190
        # limitActionByJoint, control_objects_limits, apply_action and stepSimulation all run inside rr_step
        self._cmd.copy_(self._torch.from_numpy(joint_action.reshape(1, 9)))
        self._sim.step(self._cmd)
        observation = self.get_observation(bool(np.any(camera_on)))
        reward = self.reward_func(observation)
        done = False
        self.timestep += 1
        if self.goal_idx < 0:
            if self.timestep >= self.intrinsic_timesteps:
                done = True
        else:
            if self.timestep >= self.extrinsic_timesteps:
                done = True
        return observation, reward, done, {}

    def _ik(self, coords, orientation):
        t = self._torch
        pq = t.tensor([list(coords) + list(orientation)], dtype=t.float64, device=self._sim.device)
        return self._sim.ik(pq, max_iter=1000, residual=0.001)[0].cpu().numpy()

    def step_cartesian(self, action):
        if action['cartesian_command'] is None:
            joint_action = {"joint_command": np.zeros(9), "render": action['render']}
        else:
            coords = action['cartesian_command'][:3]
            desiredOrientation = action['cartesian_command'][3:]
            sameCoords = np.all(coords == self.requested_coords)
            sameOrient = np.all(desiredOrientation == self.requested_orient)
            if sameCoords and sameOrient:
                arm_joints = self.last_ik
            else:
                arm_joints = self._ik(coords, desiredOrientation)
                self.last_ik = arm_joints
                self.requested_coords = coords
                self.requested_orient = desiredOrientation
            gripper_joints = action['gripper_command']
            all_joints = np.hstack([arm_joints[:7], gripper_joints])
            joint_action = {"joint_command": all_joints, "render": action['render']}
        return self.step_joints(joint_action)

    def step_macro(self, action):
        macro_action = action['macro_action']
        if macro_action is None:
            joint_action = {"joint_command": np.zeros(9), "render": action['render']}
        else:
            sameAction = np.all(macro_action == self.requested_action)
            joints = None
            if sameAction:
                joints = self.next_step()
            if not sameAction or joints is None:
                self.requested_action = macro_action
                self.generate_plan(macro_action)
                joints = self.next_step()
            joint_action = {"joint_command": joints, "render": action['render']}
        return self.step_joints(joint_action)

    def generate_plan(self, macro_action):
        orientation = euler_to_quat([0, 3.14, -1.57])

        def goToPosXY(coords):
            return self._ik(coords, orientation)[:9]
        self.planned_actions = macro_plan(goToPosXY, macro_action[0], macro_action[1])
        self.plan_step = -1

    def next_step(self):
        self.plan_step += 1
        if self.plan_step < len(self.planned_actions):
            return self.planned_actions[self.plan_step, :]
        return None


class EyeCamera:
    """env.py:516-567: eye position / target bookkeeping.  The image itself comes from rr_render, whose view
    (eye [0.01,0,1.2] -> table position, up z) and projection (fov 80, 4/3, .1, 100) are fixed to the reference's."""

    def __init__(self, eyePosition, targetPosition, fov=80, width=320, height=240):
        self.eyePosition, self.targetPosition, self.upVector = eyePosition, targetPosition, [0, 0, 1]
        self.fov, self.render_width, self.render_height = fov, width, height
        self.pitch_roll = False
