"""GPU tests of the reference-facing Python surface: REALRobotEnv (N = 1 gym facade), VecREALRobotEnv, the debug
camera, macro / cartesian actions and evaluate(), checked against the oracle where arithmetic is involved."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import real_robots  # noqa: E402
from oracle.orc import Oracle, eye_matrices  # noqa: E402
from real_robots_b200 import cameras  # noqa: E402
from real_robots_b200.envs.env import Goal  # noqa: E402
from real_robots_b200.envs.vec_env import VecREALRobotEnv  # noqa: E402
from real_robots_b200.policy import BasePolicy, RandomBatchedPolicy  # noqa: E402


def test_gym_facade_joints_matches_oracle():
    env = real_robots.make('REALRobot2020-R1J3-v0')
    obs = env.reset()
    assert set(obs) == {'joint_positions', 'touch_sensors', 'retina', 'depth', 'mask', 'object_positions', 'goal',
                        'goal_mask', 'goal_positions'}
    assert isinstance(obs['joint_positions'], list) and len(obs['joint_positions']) == 9      # robot.py:204-211 returns a list
    assert obs['retina'].shape == (240, 320, 3) and obs['retina'].dtype == np.uint8
    assert obs['mask'].dtype == np.int32 and set(np.unique(obs['mask'])) == {-1, 0, 1, 2, 3, 4}
    assert (obs['goal'] == 0).all() and obs['goal_mask'] is None
    o = Oracle(3)
    r0, d0, m0 = o.render()
    assert (obs['mask'] == m0).all() and (obs['retina'] == r0).all() and (obs['depth'] == d0).all()
    rng = np.random.default_rng(0)
    a = rng.uniform(env.action_space['joint_command'].low, env.action_space['joint_command'].high)
    for t in range(40):
        obs, reward, done, info = env.step({'joint_command': a, 'render': t == 39})
        o.step(a)
    assert reward == 0 and done is False and info == {}
    assert np.allclose(obs['joint_positions'], o.read()[0], atol=1e-9)
    assert np.allclose(obs['object_positions']['cube'], o.get_state()[2][0, :3], atol=1e-7)
    assert set(obs['object_positions']) == {'cube', 'tomato', 'mustard'}
    assert (obs['mask'] == o.render()[2]).mean() > 0.9999
    # render off -> cached zero arrays (env.py:273-275)
    obs, _, _, _ = env.step({'joint_command': a, 'render': False})
    assert (obs['retina'] == 0).all() and (obs['depth'] == 0).all() and (obs['mask'] == 0).all()
    # None command = zeros(9) (env.py:333-334); part / object accessors (env.py:230-247)
    env.step({'joint_command': None, 'render': False})
    assert np.allclose(env.get_part_pos('base'), env.robot.parts['base'].get_position())
    assert env.get_obj_pose('cube').shape == (7,) and isinstance(env.get_contacts(), dict)
    # done flag (env.py:345-352)
    env.intrinsic_timesteps = env.timestep + 1
    assert env.step({'joint_command': a, 'render': False})[2] is True
    env.close()


def test_render_rgb_array_debug_camera_matches_oracle():
    env = real_robots.make('REALRobot2020-R2J2-v0')
    env.reset()
    img = env.render('rgb_array')
    assert img.shape == (240, 320, 3) and img.dtype == np.uint8
    view = cameras.view_from_yaw_pitch_roll([0, 0, .4], 1.2, 30, -30, 0, 2)
    proj = cameras.projection_fov(80, 320.0 / 240, 0.1, 100.0)
    o = Oracle(2)
    ref = o.render_frames(o.shape_frames(), view, proj)[0]
    assert (img == ref).all()
    assert (img != 255).any(axis=2).mean() > 0.05          # something is visible
    # the Python float32 camera helpers reproduce the oracle's eye camera bit for bit
    v, p = eye_matrices()
    assert (cameras.view_from_positions([0.01, 0, 1.2], [0, 0, 0.08], [0, 0, 1]) == v).all()
    assert (cameras.projection_fov(80, 320.0 / 240, 0.1, 100.0) == p).all()
    # and the eye render is unchanged afterwards (the constants were restored)
    assert (env.get_retina()[0] == o.render()[0]).all()
    env.close()


def test_cartesian_and_macro_actions():
    env = real_robots.make('REALRobot2020-R1C1-v0')
    env.reset()
    quat = real_robots_b200_quat()
    cmd = np.concatenate([[-0.1, 0.1, 0.5], quat])
    for t in range(300):
        obs, _, _, _ = env.step({'cartesian_command': cmd, 'gripper_command': np.array([0.3, 0.2]), 'render': False})
    assert np.linalg.norm(env.get_part_pos('base') - cmd[:3]) < 0.01       # tests/test_actions.py:69 tolerance
    assert abs(obs['joint_positions'][7] - 0.3) < 0.02 and abs(obs['joint_positions'][8] - 0.2) < 0.02
    assert env.last_ik is not None and np.all(env.requested_coords == cmd[:3])
    env.close()
    # macro action: the plan passes the reference test script's checkpoints (tests/test_actions.py:143-152)
    env = real_robots.make('REALRobot2020-R1M1-v0')
    env.reset()
    env.robot.object_bodies['cube'].reset_pose([-0.1, 0.45, 0.6], [0, 0, 0, 1])     # out of the way, as :94-97 does
    macro = np.array([[-0.10, -0.30], [-0.05, 0.20]])
    checkpoints = {199: None, 249: None, 749: None, 849: None, 999: None}
    for t in range(1000):
        env.step({'macro_action': macro, 'render': False})
        if t in checkpoints:
            checkpoints[t] = env.get_part_pos('base').copy()
    assert env.planned_actions.shape == (1000, 9) and env.plan_step == 999
    assert np.linalg.norm(checkpoints[199] - [-0.10, -0.30, 0.6]) < 0.01
    assert np.linalg.norm(checkpoints[249] - [-0.10, -0.30, 0.46]) < 0.01
    assert np.linalg.norm(checkpoints[749] - [-0.05, 0.20, 0.46]) < 0.015
    # joint motion is rate limited to 10% of min(error, 0.2..0.3 rad) per step (env.py:314-321 + the motor gain), so 100
    # steps after leaving p2 the arm is still approaching home2
    assert np.linalg.norm(checkpoints[849] - [-0.419, 0, 1.14]) < 0.12      # home2
    assert np.linalg.norm(checkpoints[999] - [-0.55, 0, 1.271]) < 0.05      # home
    env.close()


def real_robots_b200_quat():
    from real_robots_b200.util import euler_to_quat
    return euler_to_quat([0, 3.14, -1.57])


def test_vec_env_matches_single_env_semantics():
    N = 8
    env = VecREALRobotEnv(N, objects=3, action_type='joints', additional_obs=True)
    obs = env.reset()
    assert obs['retina'].shape == (N, 240, 320, 3) and obs['mask'].shape == (N, 240, 320) and obs['object_positions'].shape == (N, 3, 3)
    pol = RandomBatchedPolicy(env, seed=3)
    orcs = [Oracle(3) for _ in range(N)]
    for t in range(50):
        act = pol.step(obs, None, None)
        act['render'] = t == 49
        obs, reward, done, _ = env.step(act)
        a = act['joint_command'].cpu().numpy()
        for e in range(N):
            orcs[e].step(a[e])
    assert reward.shape == (N,) and not bool(done.any())
    assert np.allclose(obs['joint_positions'].cpu().numpy(), np.array([o.read()[0] for o in orcs]), atol=1e-8)
    # macro actions, batched plans == the N = 1 reference logic run per env
    menv = VecREALRobotEnv(4, objects=1, action_type='macro_action')
    menv.reset(render=False)
    macro = torch.tensor([[[-0.2, -0.3], [0.0, 0.3]], [[-0.1, 0.2], [-0.2, -0.2]], [[0.0, 0.0], [0.05, 0.4]], [[-0.25, -0.45], [-0.25, 0.45]]],
                         dtype=torch.float64, device='cuda')
    menv.step({'macro_action': macro, 'render': False})
    plan = menv.plan.cpu().numpy()
    from real_robots_b200.envs.env import macro_plan
    quat = real_robots_b200_quat()
    for e in range(4):
        o = Oracle(1)
        ref = macro_plan(lambda c: o.ik(c, quat)[0][:9], macro[e, 0].cpu().numpy(), macro[e, 1].cpu().numpy())
        assert np.abs(plan[e] - ref).max() < 1e-7
    assert int(menv.plan_step[0]) == 0
    menv.step({'macro_action': macro, 'render': False})
    assert int(menv.plan_step[0]) == 1
    env.close(); menv.close()


def test_evaluate_with_goal_file(tmp_path):
    """real_robots.evaluate() end to end on a synthetic goals file in the reference's pickle format."""
    goals = []
    for k in range(2):
        pos = np.array([-0.1 + 0.02 * k, 0.05 * k, 0.45, 0, 0, 0, 1.0])
        goals.append(Goal(initial_state={'cube': pos}, final_state={'cube': pos + [0.1, 0, 0.3214 - 0.45, 0, 0, 0, 0]}, challenge='2D',
                          retina=np.zeros((240, 320, 3), np.uint8), mask=np.zeros((240, 320), np.int32)))
    path = str(tmp_path / "goals.npy")
    np.savez_compressed(path, np.array(goals, dtype=object))

    class Still(BasePolicy):
        calls = 0

        def step(self, observation, reward, done):
            Still.calls += 1
            return {'joint_command': np.zeros(9), 'render': False}

    score, scores = real_robots.evaluate(Still, environment='R1', action_type='joints', n_objects=1, intrinsic_timesteps=5,
                                         extrinsic_timesteps=120, extrinsic_trials=2, visualize=False, goals_dataset_path=path + ".npz")
    assert Still.calls == 5 + 2 * 120
    assert set(score) == {'score_2D', 'score_2.5D', 'score_3D', 'score_total'} and len(scores['2D']) == 2
    # the cube is dropped at its start pose and rests there: final xy ~ initial, goal is 10 cm away -> score ~ 0.25 (env.py:192-194)
    assert all(0.15 < s < 0.30 for s in scores['2D'])


def test_evaluate_vectorised_single_rank():
    from real_robots_b200.evaluate import evaluate_vectorised
    G = 6
    rng = np.random.default_rng(2020)
    init = np.zeros((G, 1, 7)); init[:, 0, :3] = np.c_[rng.uniform(-0.25, 0.05, G), rng.uniform(-0.45, 0.45, G), np.full(G, 0.40)]; init[:, 0, 6] = 1
    final = init[:, :, :3].copy(); final[:, 0, 0] += 0.1
    score, scores, vec = evaluate_vectorised(lambda env: _Still(env), dict(initial=init, final=final), n_objects=1, action_type='joints',
                                             extrinsic_timesteps=100)
    assert vec.shape == (G,) and np.isfinite(vec).all() and (vec > 0.1).all() and (vec < 0.35).all()
    assert np.isclose(score['score_total'], vec.mean())


class _Still:
    def __init__(self, env):
        self.z = torch.zeros(env.n_envs, 9, dtype=torch.float64, device=env.device)

    def step(self, obs, reward, done):
        return {'joint_command': self.z, 'render': False}


def test_goal_generation_roundtrip(tmp_path):
    """generate_goals on the batched sim -> reference file format -> REALRobotEnv.set_goal / evaluateGoal."""
    from real_robots_b200.generate_goals import generate_goals, save_goals
    goals = generate_goals(3, n_objects=1, challenge='2D', seed=5, batch=8)
    assert len(goals) == 3
    for g in goals:
        assert set(g.initial_state) == {'cube'} and g.initial_state['cube'].shape == (7,) and g.challenge == '2D'
        assert g.retina.shape == (240, 320, 3) and g.mask.shape == (240, 320) and (g.mask == 2).any()
        assert np.linalg.norm(g.initial_state['cube'][:3] - g.final_state['cube'][:3]) >= 0.10      # min_start_goal_dist
        assert 0.30 < g.initial_state['cube'][2] < 0.33 and 0.30 < g.final_state['cube'][2] < 0.33  # settled on the table
    path = str(tmp_path / "goals")
    save_goals(goals, path)
    env = real_robots.make('REALRobot2020-R1J1-v0')
    env.set_goals_dataset_path(path + ".npz")
    env.reset()
    obs = env.set_goal()
    assert env.goal_idx == 0 and (obs['goal'] == goals[0].retina).all() and obs['goal_positions']['cube'].shape == (3,)
    assert np.allclose(env.get_obj_pose('cube')[:3], goals[0].initial_state['cube'][:3], atol=1e-9)
    ch, score = env.evaluateGoal()
    assert ch == '2D' and 0 < score < 0.3
    env.close()
    # 2.5D: x is drawn over table and shelf; an accepted goal has an object on the shelf at one end of the move (:314-351)
    from real_robots_b200.generate_goals import SHELF_Z
    goals = generate_goals(2, n_objects=2, challenge='2.5D', seed=9, batch=32, min_start_goal_dist=0.2, min_objects_dist=0.25)
    assert len(goals) == 2
    for g in goals:
        assert g.challenge == '2.5D' and g.subtype == '2'
        assert any(st[n][2] > SHELF_Z[n] for st in (g.initial_state, g.final_state) for n in st)
        for n in g.initial_state:
            assert np.linalg.norm(g.initial_state[n][:2] - g.final_state[n][:2]) >= 0.2
        p = np.array([g.final_state[n][:3] for n in g.final_state])
        assert np.linalg.norm(p[0] - p[1]) >= 0.25


def test_videomaker_writes_frames(tmp_path, monkeypatch):
    """VideoMaker (videomaker.py:11-190) filmed through rr_render_camera: files named like the reference's, one frame
    every 8 simulation steps, robot visible from the video camera (distance 1, yaw 90, pitch -45, fov 90)."""
    import cv2
    from real_robots.videomaker import VideoMaker, Frames
    monkeypatch.chdir(tmp_path)
    env = real_robots.make('REALRobot2020-R1J1-v0')
    env.intrinsic_timesteps, env.extrinsic_trials = 40, 1
    obs = env.reset()
    vm = VideoMaker(env, Frames([0, 40]), Frames(0), debug=True)
    frame = vm.camera.render(env)
    assert frame.shape == (240, 320, 3) and frame.dtype == np.uint8
    assert (frame != 255).any(axis=2).mean() > 0.1            # table and robot fill a good part of the view
    vm.start_intrinsic()
    for t in range(1, 41):
        obs, _, _, _ = env.step({'joint_command': np.zeros(9), 'render': t == 40})
        vm.update_intrinsic(t)
    vm.end_intrinsic()
    vm.start_trial({'goal': obs['retina'], 'retina': obs['retina']}, 0)
    for t in range(1, 17):
        vm.extrinsic_trial(obs, None, t, {'score_total': 0.5, 'score_2D': 0.5, 'score_2.5D': 0, 'score_3D': 0})
    vm.end_trial()
    assert len(vm.files) == 2 and vm.files[0].endswith("-intrinsic.avi") and vm.files[1].endswith("-trial-0.avi")
    counts = []
    for f in vm.files:
        cap = cv2.VideoCapture(f)
        n = 0
        while cap.read()[0]:
            n += 1
        counts.append(n)
    assert counts == [5, 2]
    env.close()


def test_cartesian_8192_envs_full_observation():
    """BASELINE config 5 shape: 8192 envs on one GPU, cartesian action (xyz in the reference's Box env.py:44-47, un-normalised
    quaternion components in [-1, 1], gripper in [0, pi/2] :54-55), full R1 observation.  Size-independent properties:
    the IK of a sampled env agrees with the oracle's, duplicated commands give identical envs, the request cache
    (env.py:366-370) suppresses the second solve, outputs have the reference's shapes / dtypes / value sets."""
    N = 8192
    env = VecREALRobotEnv(N, objects=3, action_type='cartesian', additional_obs=True)
    obs = env.reset(render=False)
    g = torch.Generator(device='cuda'); g.manual_seed(11)
    lo = torch.tensor([-0.25, -0.5, 0.42, -1, -1, -1, -1], dtype=torch.float64, device='cuda')
    hi = torch.tensor([0.05, 0.5, 0.6, 1, 1, 1, 1], dtype=torch.float64, device='cuda')
    cart = lo + (hi - lo) * torch.rand(N, 7, dtype=torch.float64, device='cuda', generator=g)
    cart[N // 2:] = cart[:N // 2]                                            # second half repeats the first
    grip = torch.rand(N, 2, dtype=torch.float64, device='cuda', generator=g) * (np.pi / 2)
    grip[N // 2:] = grip[:N // 2]
    def launches(**kw):
        n = env.sim.launch_count()
        out = env.step(dict({'cartesian_command': cart, 'gripper_command': grip}, **kw))
        return env.sim.launch_count() - n, out
    _, (obs_off, reward, done, _) = launches(render=False)     # IK for every env, narrowphase in line (the reset dropped the look-ahead one)
    launches(render=False)
    plain, _ = launches(render=False)                           # steady state without render: no IK (request cache), no raster launch at all
    again, _ = launches(render=False)                           # (render off for every env: cached zero images, env.py:273-275)
    assert plain == again
    rendered, (obs, reward, done, _) = launches(render=True)
    assert rendered - plain == 2                                # + render set-up + rasteriser
    assert not obs_off['retina'].any() and obs_off['retina'].shape == (N, 240, 320, 3)
    for e in (0, 1, 77, 4095):
        ref = Oracle(3).ik(cart[e, :3].cpu().numpy(), cart[e, 3:].cpu().numpy())[0]
        assert np.abs(env.last_ik[e].cpu().numpy() - ref).max() < 1e-7
    assert obs['retina'].shape == (N, 240, 320, 3) and obs['retina'].dtype == torch.uint8
    assert obs['depth'].shape == (N, 240, 320) and obs['depth'].dtype == torch.float32
    assert obs['mask'].shape == (N, 240, 320) and obs['mask'].dtype == torch.int32
    assert obs['joint_positions'].shape == (N, 9) and obs['touch_sensors'].shape == (N, 4) and obs['object_positions'].shape == (N, 3, 3)
    assert bool(torch.isfinite(obs['joint_positions']).all()) and bool(torch.isfinite(obs['depth']).all())
    vals = torch.unique(obs['mask'][::511])
    assert set(vals.tolist()) <= {-1, 0, 1, 2, 3, 4} and {0, 1}.issubset(set(vals.tolist()))
    for k in ('retina', 'depth', 'mask', 'joint_positions', 'object_positions'):
        assert bool((obs[k][:N // 2] == obs[k][N // 2:]).all()), k            # an env does not depend on its slot
    assert reward.shape == (N,) and float(reward.abs().sum()) == 0.0 and not bool(done.any())
    launches(render=False)                                      # (a step that follows a rendered one is split over two solver kernels:
    launches(render=False)                                      #  let the schedule settle again)
    cart2 = cart.clone(); cart2[:, 0] += 0.01
    n = env.sim.launch_count()
    env.step({'cartesian_command': cart2, 'gripper_command': grip, 'render': False})
    assert env.sim.launch_count() - n - plain == 1              # a new request: + the IK kernel
    env.close()


def test_public_helpers_of_the_reference_surface():
    """control_objects_limits (env.py:257-264) and EyeCamera.renderTarget (env.py:536-567) as host-callable methods."""
    env = real_robots.make('REALRobot2020-R1J2-v0')
    obs = env.reset()
    env.robot.object_bodies['cube'].reset_pose([0.0, 0.0, 0.02], [0, 0, 0, 1])          # under the table top level
    env.robot.object_bodies['tomato'].reset_pose([0.2, 0.0, 0.25], [0, 0, 0, 1])        # past the shelf front, below it
    env.control_objects_limits()
    assert np.allclose(env.get_obj_pos('cube'), [-0.1, 0.0, 0.45], atol=1e-12)
    assert np.allclose(env.get_obj_pos('tomato'), [-0.1, -0.3, 0.45], atol=1e-12)
    rgb, depth, mask = env.eyes['eye'].renderTarget(env)
    assert rgb.shape == (240, 320, 3) and depth.dtype == np.float32 and mask.dtype == np.int32 and (mask == 2).any()
    lim = env.limitActionByJoint(np.full(9, 3.0))
    assert np.allclose(lim, [0.2] * 5 + [0.3] * 2 + [0.1] * 2, atol=1e-12)              # from q = 0
    env.close()


def test_c4_macro_action_evaluate_vectorised_equals_the_serial_facade(tmp_path):
    """BASELINE configs[3]: macro_action, 3 objects, the extrinsic phase of evaluate() on 50 seeded synthetic
    goals (object xy ~ U([-0.25, 0.05] x [-0.45, 0.45]), dropped from z = 0.40, seed 2020), one 1000-step trial per goal.
    evaluate_vectorised runs the 50 trials side by side; real_robots.evaluate() (the reference's serial flow,
    evaluate.py:281-332, through the N = 1 gym facade) runs the first three one after the other with the same policy — a
    macro action that sweeps from the cube's start to its goal.  Same kernels, same inputs: the scores must be equal."""
    from real_robots_b200.evaluate import evaluate_vectorised
    G, names = 50, ['cube', 'tomato', 'mustard']
    rng = np.random.default_rng(2020)
    init = np.zeros((G, 3, 7)); init[:, :, 6] = 1.0
    init[:, :, 0] = rng.uniform(-0.25, 0.05, (G, 3)); init[:, :, 1] = rng.uniform(-0.45, 0.45, (G, 3)); init[:, :, 2] = 0.40
    final = np.zeros((G, 3, 3))
    final[:, :, 0] = rng.uniform(-0.25, 0.05, (G, 3)); final[:, :, 1] = rng.uniform(-0.45, 0.45, (G, 3)); final[:, :, 2] = 0.3214
    macro = np.stack([init[:, 0, :2], final[:, 0, :2]], axis=1)                 # [G, 2, 2]: cube start -> cube goal

    class Sweep:
        def __init__(self, env):
            self.m = torch.from_numpy(macro[:env.n_envs]).to(env.device).contiguous()

        def step(self, obs, reward, done):
            return {'macro_action': self.m, 'render': False}

    score, scores, vec = evaluate_vectorised(lambda env: Sweep(env), dict(initial=init, final=final), n_objects=3,
                                             action_type='macro_action', extrinsic_timesteps=1000)
    assert vec.shape == (G,) and np.isfinite(vec).all() and (vec > 0).all() and (vec <= 3.0).all()
    assert np.isclose(score['score_total'], vec.mean()) and len(scores['2D']) == G
    assert vec.std() > 0.05                                                     # the trials differ (the sweep moves the cubes)

    K = 3
    goals = [Goal(initial_state={n: init[g, i] for i, n in enumerate(names)}, final_state={n: np.r_[final[g, i], 0, 0, 0, 1.0] for i, n in enumerate(names)},
                  challenge='2D', retina=np.zeros((240, 320, 3), np.uint8), mask=np.zeros((240, 320), np.int32)) for g in range(K)]
    path = str(tmp_path / "goals_c4.npy")
    np.savez_compressed(path, np.array(goals, dtype=object))

    class SerialSweep(BasePolicy):
        trial = -1

        def start_extrinsic_trial(self):
            SerialSweep.trial += 1

        def step(self, observation, reward, done):
            return {'macro_action': macro[SerialSweep.trial], 'render': False}

    s_obj, s_scores = real_robots.evaluate(SerialSweep, environment='R1', action_type='macro_action', n_objects=3, intrinsic_timesteps=0,
                                           extrinsic_timesteps=1000, extrinsic_trials=K, visualize=False, goals_dataset_path=path + ".npz")
    serial = np.array(s_scores['2D'])
    assert serial.shape == (K,)
    assert np.abs(serial - vec[:K]).max() < 1e-9, (serial, vec[:K])


def test_integration_md_binding_stub_on_the_gpu():
    """INTEGRATION.md section B's ctypes binding, executed verbatim against libreal_b200.so: a step through it equals the oracle."""
    from real_robots_b200 import _lib
    from tests.scenarios import integration_stub_client
    client = integration_stub_client(_lib.LIB_PATH, n_objects=1)
    o = Oracle(1)
    a = np.array([0.3, -0.2, 0.1, -0.4, 0.0, 0.5, 0.2, 0.3, 0.2])
    for _ in range(3):
        client.step(a, True)
        o.step(a)
    assert np.abs(client.joints[0] - o.read()[0]).max() < 1e-9
    r, d, m = o.render()
    assert (client.mask[0] == m).all() and (client.rgb[0] == r).all() and (client.depth[0] == d).all()
