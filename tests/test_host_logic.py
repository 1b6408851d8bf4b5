"""CPU tests of the host-side mirror of the reference interface (no GPU): ids, spaces, goals, plans, scoring."""
import pickle

import numpy as np
import pytest

import real_robots
import real_robots_b200
from real_robots_b200.envs.env import Goal, REALRobotEnv, macro_plan
from real_robots_b200.envs.robot import Kuka
from real_robots_b200.evaluate import build_score_object, shard_range, EvaluationService
from real_robots_b200.policy import BasePolicy


def test_registered_ids_match_reference():
    # real_robots/__init__.py:16-28
    ids = sorted(real_robots_b200.REGISTRY)
    assert len(ids) == 18
    assert 'REALRobot2020-R2J3-v0' in ids and 'REALRobot2020-R1M1-v0' in ids
    kw = real_robots_b200.REGISTRY['REALRobot2020-R2J3-v0']
    assert kw == dict(additional_obs=False, objects=3, action_type='joints')
    with pytest.raises(KeyError):
        real_robots_b200.make('REALRobot-v0')         # stale id used by the reference's examples


def test_action_and_observation_spaces():
    k = Kuka(additional_obs=True, objects=3)
    # SURVEY App. A.4 (robot.py:58-67)
    lo, hi = k.action_space.low, k.action_space.high
    assert np.allclose(hi[[0, 1, 3, 5]], np.pi * 0.666) and np.allclose(hi[[2, 4]], np.pi * 0.944)
    assert np.isclose(hi[6], np.pi * 0.972) and np.allclose(lo[7:], 0) and np.allclose(hi[7:], np.pi / 2)
    assert np.allclose(lo[:7], -hi[:7])
    sp = k.observation_space.spaces
    assert sp['retina'].shape == (240, 320, 3) and sp['retina'].dtype == np.uint8
    assert sp['mask'].dtype == np.int32 and sp['depth'].shape == (240, 320)
    assert set(sp['object_positions'].spaces) == {'cube', 'tomato', 'mustard'}
    assert set(Kuka(False, 1).observation_space.spaces) == {'joint_positions', 'touch_sensors', 'retina', 'depth', 'goal'}
    assert Kuka(True, 2).used_objects == ['table', 'cube', 'tomato']


def test_env_construction_and_errors():
    env = real_robots.make('REALRobot2020-R1C2-v0')
    assert isinstance(env, REALRobotEnv)
    assert set(env.action_space.spaces) == {'cartesian_command', 'gripper_command', 'render'}
    assert np.allclose(env.cartesian_space.low, [-0.25, -0.5, 0.40, -1, -1, -1, -1])
    assert np.allclose(env.macro_space.high, [[0.05, 0.5], [0.05, 0.5]])
    assert env.intrinsic_timesteps == int(15e6) and env.extrinsic_timesteps == int(10e3) and env.extrinsic_trials == 50
    assert (env.goal.retina == 0).all() and env.goal_idx == -1 and env.goals is None
    with pytest.raises(ValueError):
        REALRobotEnv(action_type='bogus')
    with pytest.raises(AssertionError):
        env.set_goals_dataset_path('/nonexistent/goals.npz')
    assert env.render('human').size == 0 and env.isRender


def test_goal_pickle_roundtrip_under_reference_class_path(tmp_path):
    g = Goal(initial_state={'cube': np.arange(7.0)}, final_state={'cube': np.arange(7.0) + 1}, challenge='2D',
             retina=np.zeros((240, 320, 3), np.uint8), mask=np.zeros((240, 320), np.int32))
    assert pickle.loads(pickle.dumps(g)).__class__ is Goal
    assert b"real_robots.envs.env" in pickle.dumps(g)
    # the reference's file format: np.savez_compressed(path, list_of_goals) (generate_goals.py:435, env.py:143-145)
    path = str(tmp_path / "goals.npy")
    np.savez_compressed(path, np.array([g, g], dtype=object))
    env = real_robots.make('REALRobot2020-R1J1-v0')
    env.set_goals_dataset_path(path + ".npz")
    env.load_goals()
    assert len(env.goals) == 2 and env.goals[0].challenge == '2D'


def test_macro_plan_layout():
    # env.py:447-458: 100 home2, 100 p1 high, 50 p1 low, 500 sweep, 50 p2 high, 100 home2, 100 home = 1000 rows
    calls = []

    def fake_ik(coords):
        calls.append(np.array(coords))
        return np.full(9, float(len(calls)))
    plan = macro_plan(fake_ik, np.array([-0.2, -0.4]), np.array([0.0, 0.4]))
    assert plan.shape == (1000, 9)
    assert np.allclose(plan[0], [0, 0, 0, 0, 0, np.pi / 2, np.pi / 2, 0, 0]) and np.allclose(plan[-1], 0)
    dist = np.linalg.norm([0.2, 0.8])
    pieces = int(dist / 0.05) + 1
    assert len(calls) == 3 + pieces                       # 4 + pieces IK solves in the reference, one of them unused
    assert np.allclose(plan[100], 1) and np.allclose(plan[200], 2)
    chunk = 500 // pieces
    assert np.allclose(plan[250], 4) and np.allclose(plan[250 + chunk], 5)     # first two sweep pieces (IK calls 4 and 5)
    assert np.allclose(plan[750], 3) and np.allclose(plan[749], 3 + pieces)    # p2 high after the last sweep piece
    assert np.allclose(calls[0], [-0.2, -0.4, 0.6]) and np.allclose(calls[1], [-0.2, -0.4, 0.46])
    assert np.allclose(calls[3], [-0.2 + 0.2 / pieces, -0.4 + 0.8 / pieces, 0.46])


def test_score_object_and_sharding():
    so = build_score_object({'2D': [0.2, 0.4], '3D': [1.0]})
    assert np.isclose(so['score_2D'], 0.3) and so['score_2.5D'] == 0 and np.isclose(so['score_3D'], 1.0)
    assert np.isclose(so['score_total'], np.mean([0.2, 0.4, 1.0]))
    assert [shard_range(50, r, 8) for r in range(8)] == [(0, 7), (7, 14), (14, 21), (21, 28), (28, 35), (35, 42), (42, 49), (49, 50)]
    assert shard_range(3, 5, 8) == (3, 3)


def test_evaluation_service_checks():
    class P(BasePolicy):
        pass
    with pytest.raises(Exception, match="macro_action in Round 2"):
        EvaluationService(P, environment='R2', action_type='macro_action', visualize=False)
    with pytest.raises(Exception, match="R1 or R2"):
        EvaluationService(P, environment='R3', visualize=False)
    with pytest.raises(Exception, match="1, 2 or 3"):
        EvaluationService(P, environment='R1', action_type='joints', n_objects=4, visualize=False)
    with pytest.raises(Exception, match="Sub-Class"):
        EvaluationService(object, environment='R1', action_type='joints', n_objects=1, visualize=False,
                          goals_dataset_path=__file__)


def test_random_batched_policy_is_the_vectorised_cli_random_policy():
    """cli.py:12-20 per env: an action drawn from the joints Box, kept until a p = 0.05 redraw; seeded and reproducible."""
    import torch
    from real_robots_b200.envs.robot import Kuka
    from real_robots_b200.envs.vec_env import VecREALRobotEnv
    from real_robots_b200.policy import RandomBatchedPolicy
    n = 2000
    env = object.__new__(VecREALRobotEnv)
    env.n_envs, env.device = n, torch.device("cpu")
    robot = Kuka(True, 3)
    env.joints_low = torch.tensor(robot.min_joints, dtype=torch.float64)
    env.joints_high = torch.tensor(robot.max_joints, dtype=torch.float64)
    pi, pi2 = RandomBatchedPolicy(env, seed=7), RandomBatchedPolicy(env, seed=7)
    prev = pi.action.clone()
    changed = 0
    for _ in range(50):
        a = pi.step(None, None, None)
        assert a["render"] is True and a["joint_command"].shape == (n, 9)
        cmd = a["joint_command"]
        assert bool((cmd >= env.joints_low).all()) and bool((cmd <= env.joints_high).all())
        moved = (cmd != prev).any(dim=1)
        assert bool(((cmd != prev).all(dim=1) == moved).all())        # a redraw replaces the whole 9-vector
        changed += int(moved.sum())
        prev = cmd.clone()
        assert torch.equal(pi2.step(None, None, None)["joint_command"], cmd)
    rate = changed / (50 * n)
    assert 0.045 < rate < 0.055, rate
    spread = (cmd.max(dim=0).values - cmd.min(dim=0).values) / (env.joints_high - env.joints_low)
    assert bool((spread > 0.95).all())                                 # uniform over the whole Box


def test_camera_helpers_against_the_oracle_and_geometry():
    """cameras.py restates pybullet's computeViewMatrix / computeViewMatrixFromYawPitchRoll / computeProjectionMatrixFOV
    (env.py:489-500, 543-551) in float32: the eye camera must equal the oracle's matrices bit for bit, and the
    yaw / pitch camera must sit `distance` from its target, look at it, and follow pybullet's angle conventions."""
    from oracle.orc import eye_matrices
    from real_robots_b200 import cameras
    v, p = eye_matrices()
    assert (cameras.view_from_positions([0.01, 0, 1.2], [0, 0, 0.08], [0, 0, 1]).view(np.uint32) == v.view(np.uint32)).all()
    assert (cameras.projection_fov(80, 320.0 / 240, 0.1, 100.0).view(np.uint32) == p.view(np.uint32)).all()
    for target, dist, yaw, pitch in (([0, 0, .4], 1.2, 30, -30), ([-0.3, 0, .4], 1.0, 90, -45), ([0.1, -0.2, 0.3], 2.0, 0, 0),
                                     ([0, 0, 0], 0.7, -120, 20)):
        V = cameras.view_from_yaw_pitch_roll(target, dist, yaw, pitch, 0, 2).reshape(4, 4).T.astype(np.float64)
        R, t = V[:3, :3], V[:3, 3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-6
        eye = -R.T @ t
        assert abs(np.linalg.norm(eye - target) - dist) < 1e-6
        assert np.allclose(R @ np.asarray(target, float) + t, [0, 0, -dist], atol=1e-6)       # the target is straight ahead
        # pybullet: yaw 0 / pitch 0 looks along +y from behind the target; positive yaw turns the camera about +z,
        # negative pitch raises it above the target
        yr, pr = np.radians(yaw), np.radians(pitch)
        want = np.asarray(target, float) + dist * np.array([np.sin(yr) * np.cos(pr), -np.cos(yr) * np.cos(pr), -np.sin(pr)])
        assert np.allclose(eye, want, atol=1e-6), (yaw, pitch, eye, want)
        up_in_view = R @ np.array([0, 0, 1.0])
        assert up_in_view[1] > 0 and abs(up_in_view[0]) < 1e-6                                 # world z stays up in the image
    P = cameras.projection_fov(90, 4 / 3, 0.1, 100.0).reshape(4, 4).T
    assert abs(P[1, 1] - 1.0) < 1e-6 and abs(P[0, 0] - 0.75) < 1e-6 and P[3, 2] == -1


def test_videomaker_frame_selection_and_files_without_a_gpu(tmp_path, monkeypatch):
    """VideoMaker's host logic (videomaker.py:23-130) over a stand-in simulator: default intervals, one frame every 8
    steps inside the selected frames / trials only, GOAL / START insets pasted, reference-style file names."""
    cv2 = pytest.importorskip("cv2")
    import torch
    from real_robots_b200.videomaker import Frames, VideoMaker
    monkeypatch.chdir(tmp_path)

    class Sim:
        calls = 0

        def render_camera(self, view, proj):
            Sim.calls += 1
            assert view.shape == (16,) and proj.shape == (16,) and abs(float(proj[5]) - 1.0) < 1e-6     # fov 90
            return torch.full((1, 240, 320, 3), 200, dtype=torch.uint8), None, None
    env = type("Env", (), {})()
    env.intrinsic_timesteps, env.extrinsic_trials, env._sim = 15e6, 50, Sim()
    vm = VideoMaker(env, True, True)                                    # defaults: first / middle / last minute, 5 trials
    minute = 60 * 25 * 8
    assert all(x in vm.intrinsic_frames for x in (0, minute, 7.5e6, 15e6 - minute, 15e6))
    assert minute + 1 not in vm.intrinsic_frames and 7.5e6 - 1 not in vm.intrinsic_frames
    assert len(vm.extrinsic_trials) == 5 and sum(t in vm.extrinsic_trials for t in range(50)) == 5
    assert len(VideoMaker(env, None, False).intrinsic_frames) == 0
    with pytest.raises(Exception):
        VideoMaker(env, 3, None)
    vm = VideoMaker(env, Frames([16, 40]), Frames(2), debug=True)
    vm.start_intrinsic()
    for t in range(1, 57):
        vm.update_intrinsic(t)
    vm.end_intrinsic()
    assert Sim.calls == 4                                               # steps 16, 24, 32, 40
    goal = np.zeros((240, 320, 3), np.uint8)
    start = np.full((240, 320, 3), 90, np.uint8)
    for trial in (1, 2):
        vm.start_trial({'goal': goal, 'retina': start}, trial)
        for t in range(1, 25):
            vm.extrinsic_trial(None, None, t, {'score_total': 0.5, 'score_2D': 0.5, 'score_2.5D': 0, 'score_3D': 0})
        vm.end_trial()
    assert Sim.calls == 4 + 3                                           # only trial 2 is filmed: steps 8, 16, 24
    assert len(vm.files) == 2 and vm.files[0].endswith("-intrinsic.avi") and vm.files[1].endswith("-trial-2.avi")
    assert vm.files[0].startswith("Simulation-%d-y" % vm.seed)
    cap = cv2.VideoCapture(vm.files[1])
    ok, frame = cap.read()
    assert ok and frame.shape == (240, 320, 3)
    assert frame[10, 310].mean() < 60 and 60 < frame[10, 10].mean() < 130 and frame[200, 160].mean() > 170   # GOAL inset, START inset, scene


def test_goal_generation_constraints_with_a_stand_in_simulator(monkeypatch, tmp_path):
    """generate_goals (generate_goals.py:275-365) over a stand-in VecREALRobotEnv whose objects drop straight down onto
    the table (x <= 0.05) or the shelf: accepted goals satisfy the challenge's constraints, rejected candidates never
    appear, and the file round-trips through REALRobotEnv.load_goals' format."""
    import torch
    import real_robots_b200.envs.vec_env as ve
    from real_robots_b200 import generate_goals as gg

    class Sim:
        def __init__(self, n, k):
            self.n, self.k, self.t = n, k, 0
            self.pose = torch.zeros(n, k, 7, dtype=torch.float64); self.pose[:, :, 6] = 1

        def step(self, cmd):
            self.t += 1
            rest = torch.where(self.pose[:, :, 0] <= 0.05, torch.tensor(0.32, dtype=torch.float64), torch.tensor(0.42, dtype=torch.float64))
            self.pose[:, :, 2] = torch.maximum(self.pose[:, :, 2] - 0.02, rest)        # falls 2 cm a step, then rests
            self.pose[0, :, 1] += 1e-3                                                 # env 0 never settles

        def read_state(self):
            return None, None, self.pose

        def render(self):
            return (torch.zeros(self.n, 240, 320, 3, dtype=torch.uint8), torch.zeros(self.n, 240, 320), torch.zeros(self.n, 240, 320, dtype=torch.int32))

    class Env:
        def __init__(self, n_envs, device=0, objects=3, **kw):
            self.n_envs, self.device, self.sim = n_envs, torch.device("cpu"), Sim(n_envs, objects)

        def reset(self, render=True):
            self.sim.pose[:] = 0; self.sim.pose[:, :, 6] = 1; self.sim.pose[:, :, 2] = 0.45

        def set_object_poses(self, pq, mask=None):
            self.sim.pose = pq.clone() if mask is None else torch.where(mask[:, None, None], pq, self.sim.pose)

        def close(self):
            pass
    monkeypatch.setattr(ve, "VecREALRobotEnv", Env)
    for challenge, n_obj in (('2D', 3), ('2.5D', 2), ('3D', 1)):
        goals = gg.generate_goals(6, n_obj, challenge, seed=3, batch=24, min_start_goal_dist=0.2, min_objects_dist=0.1)
        assert len(goals) == 6
        for g in goals:
            assert g.challenge == challenge and g.subtype == str(n_obj) and list(g.initial_state) == gg.OBJECTS[:n_obj]
            a = np.array([g.initial_state[o] for o in gg.OBJECTS[:n_obj]]); b = np.array([g.final_state[o] for o in gg.OBJECTS[:n_obj]])
            assert (np.linalg.norm(a[:, :2] - b[:, :2], axis=1) >= 0.2).all()          # start and goal apart in the plane
            assert gg._min_separation(a[None])[0] >= 0.1 and gg._min_separation(b[None])[0] >= 0.1
            if challenge == '2D':
                assert (a[:, 0] <= 0.05).all() and (b[:, 0] <= 0.05).all() and (a[:, 2] < 0.33).all()   # everything on the table
            else:
                assert (a[:, 2] > 0.40).any() or (b[:, 2] > 0.40).any()                  # the shelf is involved
            if challenge != '3D':
                for o, name in enumerate(gg.OBJECTS[:n_obj]):
                    q = np.array([0.0, 0.0, 0.0, 1.0])       # the orientation the stand-in's objects settle in after a reset (basePosition, :415-418)
                    assert min(np.linalg.norm(a[o, 3:] - q), np.linalg.norm(a[o, 3:] + q)) < 1e-12
            assert g.retina.shape == (240, 320, 3) and g.mask.shape == (240, 320)
    path = str(tmp_path / "goals.npy")
    gg.save_goals(goals, path)
    back = np.load(path + ".npz", allow_pickle=True)["arr_0"]
    assert len(back) == 6 and type(back[0]).__module__ == "real_robots.envs.env" and back[0].challenge == '3D'


def test_get_contacts_reads_manifolds_like_the_reference():
    """Kuka.get_contacts (robot.py:131-150) over synthetic manifolds: robot part -> names of the bodies it touches, only
    against BASE links of other bodies (contact_list passes linkIndexB = -1), distance filter 0.1, forces =
    normal impulse / dt, first-seen order."""
    import torch
    from real_robots_b200.envs.robot import Kuka, manifest
    shapes = manifest()["shapes"]
    idx = {s["name"]: i for i, s in enumerate(shapes)}
    pairs = np.array([[idx["skin_00"], idx["cube_0"]], [idx["finger_00"], idx["table_upper"]], [idx["skin_01"], idx["table_base"]],
                      [idx["cube_0"], idx["table_base"]], [idx["skin_00"], idx["tomato"]]], np.int32)
    hdr = torch.full((1, 32, 2), -1, dtype=torch.int32); hdr[0, :, 1] = 0
    pts = torch.zeros(1, 32, 4, 17, dtype=torch.float64)
    for slot, (pi, n) in enumerate([(0, 2), (1, 1), (2, 1), (3, 4), (4, 1)]):
        hdr[0, slot, 0], hdr[0, slot, 1] = pi, n
        pts[0, slot, :n, 9] = -0.001
        pts[0, slot, :n, 10] = 0.01 * (slot + 1)
    pts[0, 0, 1, 9] = 0.2                                   # second skin_00 / cube point is beyond the 0.1 threshold

    class Sim:
        cfg = type("C", (), {"dt": 0.005})()

        def contacts(self):
            return hdr, pts, pairs
    robot = Kuka(True, 3)
    robot._bind(Sim())
    assert robot.get_contacts() == {"skin_00": ["cube", "tomato"], "skin_01": ["table"]}
    forces = robot.get_contacts(forces=True)
    assert forces["skin_00"][0][0] == "cube" and abs(forces["skin_00"][0][1] - 0.01 / 0.005) < 1e-12
    assert forces["skin_00"][1][0] == "tomato" and abs(forces["skin_00"][1][1] - 0.05 / 0.005) < 1e-12
    assert robot.object_names == {0: "kuka", 1: "table", 2: "cube", 3: "tomato", 4: "mustard"}


def test_euler_to_quat_is_pybullets_convention():
    """pybullet.getQuaternionFromEuler([roll, pitch, yaw]) = rotation about fixed x, then y, then z (xyzw): equal to scipy's
    extrinsic 'xyz' — used for object_poses (robot.py:19-24), the macro plan's gripper orientation (env.py:417) and goals."""
    from scipy.spatial.transform import Rotation
    from real_robots_b200.util import euler_to_quat
    rng = np.random.default_rng(0)
    for _ in range(200):
        e = rng.uniform(-np.pi, np.pi, 3)
        q, s = np.array(euler_to_quat(e.tolist())), Rotation.from_euler('xyz', e).as_quat()
        assert min(np.abs(q - s).max(), np.abs(q + s).max()) < 1e-15
    assert np.allclose(euler_to_quat([0, 0, 0]), [0, 0, 0, 1])
    assert np.allclose(euler_to_quat([0, 0, np.pi / 2]), [0, 0, np.sqrt(0.5), np.sqrt(0.5)])
