"""CPU tests against tests/golden/reference_host_golden.npz — vectors produced by running the reference's own Python code
(tests/golden/make_reference_golden.py imports the unmodified package from /root/reference behind inert stand-ins for
pybullet / gym).  They pin (a) the oracle's action prologue, object-limit teleport and scoring, and (b) the host mirror
(spaces, plans, request caches, done logic, score aggregation) to the reference; the GPU parity tests then hold the CUDA
path to the oracle."""
import json
import os

import numpy as np
import pytest

import real_robots
from real_robots_b200.envs.env import Goal, REALRobotEnv, macro_plan
from real_robots_b200.envs.robot import Kuka
from real_robots_b200.evaluate import build_score_object
from real_robots_b200.util import euler_to_quat

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_host_golden.npz"))
META = json.loads(bytes(G["meta_json"]).decode())


def fake_ik(coords, orient=None):          # the same stand-in IK the generator gave the reference
    c = np.asarray(coords, dtype=np.float64)
    o = np.zeros(4) if orient is None else np.asarray(orient, dtype=np.float64)
    ph = c @ np.array([1.3, 2.1, 0.7]) + o @ np.array([0.11, 0.23, 0.31, 0.43])
    return np.sin(np.arange(11) * 0.37 + ph)


@pytest.fixture(scope="module")
def oracle():
    from oracle.orc import Oracle
    return Oracle(3)


def test_joint_limits_spaces_and_constants():
    k = Kuka(True, 3)
    assert np.array_equal(k.min_joints, G["limits_min"]) and np.array_equal(k.max_joints, G["limits_max"])
    assert np.array_equal(k.action_space.low, G["limits_min"]) and np.array_equal(k.action_space.high, G["limits_max"])
    env = real_robots.make('REALRobot2020-R1J3-v0')
    for name, sp in (("cartesian", env.cartesian_space), ("macro", env.macro_space), ("gripper", env.gripper_space)):
        assert np.array_equal(np.asarray(sp.low, np.float64), G["spaces_%s_low" % name])
        assert np.array_equal(np.asarray(sp.high, np.float64), G["spaces_%s_high" % name])
    assert sorted(env.observation_space.spaces) == META["obs_keys_R1"]
    assert sorted(real_robots.make('REALRobot2020-R2J1-v0').observation_space.spaces) == META["obs_keys_R2"]
    for key, (shape, dtype) in META["obs_shapes_R1"].items():
        sp = env.observation_space.spaces[key]
        assert list(sp.shape) == shape and str(sp.dtype) == dtype, key
    assert [int(env.intrinsic_timesteps), int(env.extrinsic_timesteps), int(env.extrinsic_trials)] == META["timesteps"]
    for n in (1, 2, 3):
        assert Kuka(True, n).used_objects == META["used_objects"][str(n)]
    assert list(k.robot_position) == META["robot_position"] and k.contact_threshold == META["contact_threshold"]
    names = ["table", "cube", "tomato", "mustard"]
    mine = np.array([list(Kuka.object_poses[n][:3]) + list(euler_to_quat(Kuka.object_poses[n][3:])) for n in names])
    assert np.abs(mine - G["quat_object_poses"]).max() < 1e-15


def test_oracle_action_prologue_matches_reference(oracle):
    """limitActionByJoint + apply_action (env.py:314-321, robot.py:188-201): the 11 motor targets for 64 (state, command) pairs."""
    q, cmd, want = G["prologue_q"], G["prologue_cmd"], G["prologue_targets"]
    for i in range(len(q)):
        oracle.reset()
        oracle.set_state(q=q[i], qd=np.zeros(11))
        oracle.step(cmd[i])
        assert np.abs(oracle.targets() - want[i]).max() < 1e-15, i


def test_oracle_object_limits_match_reference(oracle):
    """control_objects_limits (env.py:257-264): the same objects are sent back to their initial pose."""
    pos, want = G["objlimit_pos"], G["objlimit_reset"]
    init = G["quat_object_poses"][1:, :3]
    for i in range(len(pos)):
        oracle.reset()
        _, _, obj = oracle.get_state()
        obj[:, :3] = pos[i]
        obj[:, 3:7] = [0, 0, 0, 1]
        obj[:, 7:] = 0
        oracle.set_state(obj=obj)
        oracle.step(np.zeros(9))
        _, _, after = oracle.get_state()
        got = (np.linalg.norm(after[:, :3] - init, axis=1) < 5e-3).astype(np.int32)
        far = np.linalg.norm(pos[i] - init, axis=1) > 2e-2          # otherwise the check cannot tell
        assert np.array_equal(got[far], want[i][far]), (i, pos[i], got, want[i])


def test_scores_match_reference():
    cur, goal, want = G["score_cur"], G["score_goal"], G["score_value"]
    env = real_robots.make('REALRobot2020-R1J3-v0')
    objs = ["cube", "tomato", "mustard"]
    for i in range(len(cur)):
        env.goal = Goal(initial_state={}, final_state={n: np.concatenate([goal[i, o], [0, 0, 0, 1]]) for o, n in enumerate(objs)}, challenge='2D')
        env.robot.object_bodies = {n: type("B", (), {"get_position": (lambda p: (lambda self: p))(cur[i, o])})() for o, n in enumerate(objs)}
        ch, s = env.evaluateGoal()
        assert ch == '2D' and abs(s - want[i]) < 1e-14
    assert abs(want[0] - 3.0) < 1e-15 and abs(want[1] - 0.75) < 1e-12        # 0.25 per object at 10 cm (env.py:192-194)
    so = build_score_object({"2D": [0.5, 0.25, 0.75], "3D": [1.0], "2.5D": [0.125]})
    assert {k: float(v) for k, v in so.items()} == META["scoreobj"]
    assert {k: float(v) for k, v in build_score_object({"2D": [0.5]}).items()} == META["scoreobj_single"]


def test_macro_plans_match_reference_bit_for_bit():
    calls = []

    def ik9(coords):
        calls.append(np.array(coords))
        return fake_ik(coords, euler_to_quat([0, 3.14, -1.57]))[:9]
    for m, want, ncall in zip(G["plan_macro"], G["plan_rows"], G["plan_ik_calls"]):
        calls.clear()
        plan = macro_plan(ik9, m[0], m[1])
        assert plan.shape == (1000, 9) and np.array_equal(plan, want)
        assert len(calls) == ncall - 1        # the reference also solves the unused low point above p2 (env.py:447)


def test_macro_sequencing_and_cartesian_cache_match_reference():
    env = real_robots.make('REALRobot2020-R1M1-v0')
    sent = []
    env._ik = lambda coords, orient: fake_ik(coords, orient)
    env.step_joints = lambda a: sent.append(np.array(a["joint_command"]))
    macros = G["plan_macro"]
    steps = []
    for idx in G["macro_seq_action_index"]:
        env.step_macro({"macro_action": macros[idx], "render": False})
        steps.append(env.plan_step)
    assert np.array_equal(np.array(steps), G["macro_seq_plan_step"])
    assert np.array_equal(np.array(sent[:8]), G["macro_seq_first_rows"]) and np.array_equal(np.array(sent[-3:]), G["macro_seq_last_rows"])
    env.step_macro({"macro_action": None, "render": True})
    assert np.all(sent[-1] == 0)
    # cartesian request cache (env.py:366-370)
    env = real_robots.make('REALRobot2020-R1C1-v0')
    calls, sent = [], []

    def ik(coords, orient):
        calls.append(1)
        return fake_ik(coords, orient)
    env._ik = ik
    env.step_joints = lambda a: sent.append(np.array(a["joint_command"]))
    counts = []
    for c, g in zip(G["cart_cmd"], G["cart_grip"]):
        env.step_cartesian({"cartesian_command": c, "gripper_command": g, "render": False})
        counts.append(len(calls))
    assert np.array_equal(np.array(counts), G["cart_ik_calls"]) and np.array_equal(np.array(sent), G["cart_joint_cmd"])


def test_done_logic_matches_reference():
    """step_joints (env.py:345-352) with a stand-in simulator: done after `intrinsic_timesteps` steps without a goal,
    after `extrinsic_timesteps` with one; reward 0, info {}."""
    env = real_robots.make('REALRobot2020-R1J1-v0')

    class FakeTensor:
        def copy_(self, other):
            return self
    env._sim = type("S", (), {"step": lambda self, c: None})()
    env._cmd = FakeTensor()
    env._torch = type("T", (), {"from_numpy": staticmethod(lambda a: a)})
    env.get_observation = lambda cam=True: {}
    env.intrinsic_timesteps, env.extrinsic_timesteps, env.timestep, env.goal_idx = 3, 2, 0, -1
    dones = []
    for t in range(3):
        o, r, d, info = env.step_joints({"joint_command": None if t == 0 else np.zeros(9), "render": t == 1})
        assert r == 0 and info == {}
        dones.append(bool(d))
    env.goal_idx, env.timestep = 0, 0
    for t in range(2):
        dones.append(bool(env.step_joints({"joint_command": np.zeros(9), "render": False})[2]))
    assert dones == META["done_sequence_intrinsic3_then_extrinsic2"]
    # the order of operations inside one step that the reference executes and rr_step reproduces (env.py:336-342)
    assert META["step_joints_call_order"] == ["limitActionByJoint", "control_objects_limits", "apply_action", "global_step", "get_observation(False)"]
    env._sim = None


def test_touch_sensor_semantics_of_the_reference():
    """robot.py:131-163: a sensor reads the LARGEST normal force among its link's contact points closer than 0.1 (strict)."""
    assert np.array_equal(G["touch_sensors"], [7.25, 0.0, 0.5, 0.0])
    assert META["touch_contacts"] == {"finger_00": ["cube"], "skin_00": ["cube", "tomato"], "skin_10": ["table"]}


def test_evaluation_service_flow_matches_reference(monkeypatch):
    """The reference's EvaluationService was driven end to end with tests/golden/fake_env.py; this repo's must make the same
    calls in the same order (controller hooks, env.reset / set_goal / step / evaluateGoal), end in the same
    evaluation_state and raise on the same bad arguments (evaluate.py:44-71, 148-191, 203-417)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import fake_env
    import real_robots_b200
    from real_robots_b200.evaluate import EvaluationService
    from real_robots_b200.policy import BasePolicy
    log, made = [], []
    fake_env.FakeEnv.log = log
    monkeypatch.setattr(real_robots_b200, "make", lambda env_id: (made.append(env_id), fake_env.FakeEnv())[1])
    svc = EvaluationService(fake_env.make_controller(BasePolicy, log), environment='R1', action_type='macro_action', n_objects=2,
                            intrinsic_timesteps=3, extrinsic_timesteps=2, extrinsic_trials=4, visualize=True,
                            goals_dataset_path="goals.npz", video=None)
    svc.run_intrinsic_phase()
    svc.run_extrinsic_phase()
    assert made == META["eval_env_id"][:1]
    assert log == META["eval_log"]
    assert fake_env.plain(svc.scores) == META["eval_scores"]
    assert fake_env.plain(svc.evaluation_state) == META["eval_state"]
    log2 = []
    fake_env.FakeEnv.log = log2
    svc = EvaluationService(fake_env.make_controller(BasePolicy, log2), 'R2', 'joints', 1, 0, 1, 1, False, "g.npz", None)
    svc.run_intrinsic_phase()
    assert fake_env.plain(svc.evaluation_state) == META["eval_skip_state"] and log2 == META["eval_skip_log"]
    for args, msg in META["eval_errors"].items():
        if args == "not_a_policy":
            with pytest.raises(Exception):
                EvaluationService(object, 'R1', 'joints', 1, 1, 1, 1, False, "g.npz", None)
            continue
        with pytest.raises(Exception) as ei:
            EvaluationService(fake_env.make_controller(BasePolicy, []), *eval(args), 1, 1, 1, False, "g.npz", None)
        assert str(ei.value) == msg


def test_registry_and_policy_interface_match_reference():
    """The 18 env ids with their constructor kwargs (real_robots/__init__.py:16-28, recorded from the reference's own
    register() calls) and the BasePolicy method signatures (policy.py)."""
    import inspect
    import real_robots_b200
    from real_robots_b200.policy import BasePolicy
    ref = META["registry"]
    assert sorted(real_robots_b200.REGISTRY) == sorted(ref)
    for env_id, entry in ref.items():
        assert real_robots_b200.REGISTRY[env_id] == entry["kwargs"], env_id
        assert entry["entry_point"] == "real_robots.envs:REALRobotEnv"
    assert real_robots.envs.REALRobotEnv is REALRobotEnv                       # the entry point resolves under the alias package
    mine = {n: str(inspect.signature(f)) for n, f in inspect.getmembers(BasePolicy, inspect.isfunction)}
    for name, sig in META["base_policy"].items():
        assert mine.get(name) == sig, (name, mine.get(name), sig)


def test_settle_detector_and_placement_thresholds_match_reference():
    """generate_goals.runEnv (:21-68) was run on scripted object motions (tests/golden/fake_env.py); the batched settle()
    must call the same poses still / not still for every script and step budget.  Also isOnTable / isOnShelf (:245-272)."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import fake_env
    from real_robots_b200 import generate_goals as gg
    scripts = fake_env.SCRIPTS

    class FakeSim:
        def __init__(self):
            self.s = 0

        def step(self, cmd):
            self.s += 1

        def read_state(self):
            poses = np.stack([fake_env.scripted_pose(sc, self.s) for sc in scripts])[:, None, :]      # [N, 1 object, 7]
            return None, None, torch.from_numpy(poses)
    for max_t in fake_env.MAX_TS:
        env = type("E", (), {})()
        env.n_envs, env.device, env.sim = len(scripts), torch.device("cpu"), FakeSim()
        written = {}
        env.set_object_poses = lambda pq, mask=None: written.update(pq=pq.clone(), mask=mask.clone())
        still, poses = gg.settle(env, max_t=max_t)
        for i, sc in enumerate(scripts):
            failed, t_ref = META["settle"]["%s/%d" % (sc, max_t)]
            assert bool(still[i]) == (not failed), (sc, max_t)
            if not failed:      # the poses are latched at the step the env met the test (runEnv returns there), not at the end of the batch
                assert np.allclose(poses[i, 0].numpy(), fake_env.scripted_pose(sc, int(t_ref) + 1)), (sc, max_t, t_ref)
        assert (written["mask"] == still).all() and torch.equal(written["pq"], poses)
    zs = (0.30, 0.329, 0.331, 0.339, 0.341, 0.394, 0.396, 0.399, 0.401, 0.45)
    for obj, want in META["shelf_table_thresholds"].items():
        mine = [z for z in zs if z < gg.TABLE_Z[obj]] + [-z for z in zs if z > gg.SHELF_Z[obj]]
        assert mine == want, obj


def test_generate_position_matches_reference_draw_for_draw():
    """generatePosition (generate_goals.py:79-108): same legacy-RNG call order, so RandomState(seed) here reproduces the
    reference under np.random.seed(seed) — x range by table plane, y in [-.45, .45], z = 0.40 / 0.50, free orientation."""
    from real_robots_b200 import generate_goals as gg
    for (seed, plane), want in zip(((1, None), (2, True), (3, False), (4, True), (5, None)), G["genpos_free"]):
        rng = np.random.RandomState(seed)
        got = np.array([gg.generate_position(rng, "cube", False, plane) for _ in range(4)])
        assert np.abs(got - want).max() < 1e-15, (seed, plane)


def test_extrinsic_formula_matches_reference():
    env = real_robots.make('REALRobot2020-R1J1-v0')
    for row in G["extrinsic_formula"]:
        p, p_goal, a, a_goal, want = row[0:3], row[3:6], row[6:10], row[10:14], row[14:17]
        got = [env.extrinsicFormula(p_goal, p, a_goal, a, w) for w in (1, 0.5, 0.0)]
        assert np.abs(np.array(got) - want).max() < 1e-14


def test_host_limit_action_matches_reference():
    """The public host-side limitActionByJoint (env.py:314-321) against the recorded prologue: its output, pushed through the
    reference's own clip / fan-out (recorded as motor targets), must reproduce the first 7 arm targets wherever the
    joint-limit clip of apply_action is not active."""
    env = real_robots.make('REALRobot2020-R1J3-v0')
    k = Kuka(True, 3)
    for q, cmd, want in zip(G["prologue_q"], G["prologue_cmd"], G["prologue_targets"]):
        env.robot.calc_state = lambda q=q: [float(v) for v in q[:8]] + [float(-q[8])]
        lim = env.limitActionByJoint(cmd)
        clipped = np.maximum(k.min_joints, np.minimum(lim, k.max_joints))
        clipped[8] = max(0.0, min(2 * clipped[7], clipped[8]))
        assert np.array_equal(clipped[:8], want[:8]) and clipped[8] == -want[8] and want[9] == want[7] and want[10] == want[8]
    env.set_eye("left", [0.1, 0, 1.0], [0, 0, 0])
    env.setCamera()
    assert set(env.eyes) == {"eye", "left"} and env.envCamera.dist == 1.2 and env.envCamera.yaw == 30


def test_vectorised_macro_plans_match_reference_bit_for_bit():
    """VecREALRobotEnv._generate_plans (torch, one plan per env, every IK request batched) against the reference's
    generate_plan rows (env.py:410-459) with the same stand-in IK — on CPU tensors, no simulator behind it."""
    torch = pytest.importorskip("torch")
    from real_robots_b200.envs.vec_env import VecREALRobotEnv
    macros = np.asarray(G["plan_macro"], dtype=np.float64)
    n = len(macros)
    quat = np.array(euler_to_quat([0, 3.14, -1.57]))
    env = object.__new__(VecREALRobotEnv)
    env.n_envs, env.device = n, torch.device("cpu")
    env.plan = torch.zeros(n, 1000, 9, dtype=torch.float64)
    requests = []

    def ik_xyz(xyz, mask):
        requests.append(int(mask.sum()))
        rows = [fake_ik(r, quat)[:9] if m else np.zeros(9) for r, m in zip(xyz.numpy(), mask.numpy())]
        return torch.tensor(np.array(rows), dtype=torch.float64)
    env._ik_xyz = ik_xyz
    env._generate_plans(torch.tensor(macros), torch.ones(n, dtype=torch.bool))
    got = env.plan.numpy()
    for k in range(n):
        assert np.array_equal(got[k], G["plan_rows"][k]), k
    # one batched request per plan stage: p1 high / p1 low / p2 high, then one per sweep piece of the longest plan
    assert len(requests) == 3 + int((np.asarray(G["plan_ik_calls"]) - 4).max())
    # envs that do not need a plan keep theirs
    env.plan = torch.full((n, 1000, 9), 7.0, dtype=torch.float64)
    need = torch.zeros(n, dtype=torch.bool); need[0] = True
    env._generate_plans(torch.tensor(macros), need)
    assert np.array_equal(env.plan[0].numpy(), G["plan_rows"][0]) and bool((env.plan[1:] == 7.0).all())


def _vec_shell(torch, n, action_type):
    """A VecREALRobotEnv without a simulator: the request caches and plan replay are plain torch code."""
    from real_robots_b200.envs.vec_env import VecREALRobotEnv
    env = object.__new__(VecREALRobotEnv)
    env.n_envs, env.device, env.action_type = n, torch.device("cpu"), action_type
    if action_type == 'macro_action':
        env.requested = torch.full((n, 2, 2), float('nan'), dtype=torch.float64)
        env.plan = torch.zeros(n, 1000, 9, dtype=torch.float64)
        env.plan_step = torch.full((n,), 1000, dtype=torch.int64)
    else:
        env.requested = torch.full((n, 7), float('nan'), dtype=torch.float64)
        env.last_ik = torch.zeros(n, 11, dtype=torch.float64)
    return env


def test_vectorised_macro_sequencing_matches_reference_per_env():
    """Two envs replay the reference's macro sequence (env.py:388-408, 462-467), the second one delayed by three steps of
    another action, so that at most times only one of them re-plans: plan_step and the rows sent must match the reference's
    for each env independently."""
    torch = pytest.importorskip("torch")
    quat = np.array(euler_to_quat([0, 3.14, -1.57]))
    macros = np.asarray(G["plan_macro"], dtype=np.float64)
    seq = [int(i) for i in G["macro_seq_action_index"]]
    lag = 3
    env = _vec_shell(torch, 2, 'macro_action')
    env._ik_xyz = lambda xyz, mask: torch.tensor(np.array([fake_ik(r, quat)[:9] if m else np.zeros(9)
                                                            for r, m in zip(xyz.numpy(), mask.numpy())]), dtype=torch.float64)
    steps, sent = [[], []], [[], []]
    for t in range(len(seq) + lag):
        a0 = macros[seq[t]] if t < len(seq) else macros[5]
        a1 = macros[seq[t - lag]] if t >= lag else macros[5]
        row = env._macro(torch.tensor(np.stack([a0, a1])))
        for e in range(2):
            steps[e].append(int(env.plan_step[e])); sent[e].append(row[e].numpy().copy())
    want = np.asarray(G["macro_seq_plan_step"])
    assert np.array_equal(np.array(steps[0][:len(seq)]), want) and np.array_equal(np.array(steps[1][lag:]), want)
    for rows in (sent[0][:len(seq)], sent[1][lag:]):
        assert np.array_equal(np.array(rows[:8]), G["macro_seq_first_rows"]) and np.array_equal(np.array(rows[-3:]), G["macro_seq_last_rows"])


def test_vectorised_cartesian_cache_matches_reference():
    """env.py:358-386 per env: IK only when the 7-vector request changes (exact equality), gripper passed through."""
    torch = pytest.importorskip("torch")

    class Sim:
        calls = []

        def ik(self, pq, max_iter, residual, mask=None):
            assert max_iter == 1000 and residual == 0.001
            Sim.calls.append(mask.numpy().copy())
            return torch.tensor(np.array([fake_ik(r[:3], r[3:]) for r in pq.numpy()]), dtype=torch.float64)
    env = _vec_shell(torch, 2, 'cartesian')
    env.sim = Sim()
    cmds, grips = np.asarray(G["cart_cmd"]), np.asarray(G["cart_grip"])
    out0, out1, n0 = [], [], []
    for t in range(len(cmds) + 1):                       # env 1 runs the same sequence one step later
        c0 = cmds[min(t, len(cmds) - 1)]; g0 = grips[min(t, len(cmds) - 1)]
        c1 = cmds[max(t - 1, 0)]; g1 = grips[max(t - 1, 0)]
        cmd = env._cartesian(torch.tensor(np.stack([c0, c1])), torch.tensor(np.stack([g0, g1])))
        out0.append(cmd[0].numpy().copy()); out1.append(cmd[1].numpy().copy())
        n0.append(sum(int(m[0]) for m in Sim.calls))
    assert np.array_equal(np.array(out0[:len(cmds)]), G["cart_joint_cmd"]) and np.array_equal(np.array(out1[1:]), G["cart_joint_cmd"])
    assert np.array_equal(np.array(n0[:len(cmds)]), G["cart_ik_calls"])


def test_vectorised_done_logic_and_masked_reset_match_reference():
    """VecREALRobotEnv.step_joints / reset(mask) with a stand-in simulator: per env, done follows the reference sequence
    (env.py:345-352) — also for an env that was reset in between — reward 0, info {}."""
    torch = pytest.importorskip("torch")
    from real_robots_b200.envs.vec_env import VecREALRobotEnv
    n = 3
    env = object.__new__(VecREALRobotEnv)
    env.n_envs, env.device, env.action_type, env.additional_obs = n, torch.device("cpu"), 'joints', False
    env.timestep = torch.zeros(n, dtype=torch.int64)
    env.goal_idx = -1
    env._zero9 = torch.zeros(n, 9, dtype=torch.float64)
    env._all = torch.ones(n, dtype=torch.uint8)
    log = []

    class Sim:
        def step(self, cmd):
            log.append(("step", cmd.clone()))

        def reset(self, mask):
            log.append(("reset", None if mask is None else mask.clone()))
    env.sim = Sim()
    env._observe = lambda render_mask: {"render": render_mask}
    env.intrinsic_timesteps, env.extrinsic_timesteps = 3, 2
    want = META["done_sequence_intrinsic3_then_extrinsic2"]
    dones = []
    for t in range(3):
        obs, r, d, info = env.step({'joint_command': None if t == 0 else torch.ones(n, 9, dtype=torch.float64), 'render': t == 1})
        assert info == {} and r.shape == (n,) and float(r.abs().max()) == 0.0
        assert (obs["render"] is None) == (t != 1)                  # render flag False -> no env is rendered
        dones.append(d.clone())
        if t == 0:
            assert float(log[-1][1].abs().max()) == 0.0             # joint_command None -> zero action (env.py:329-330)
            env.reset(mask=torch.tensor([0, 1, 0], dtype=torch.uint8))   # env 1 starts over after its first step
    assert [bool(d[0]) for d in dones] == want[:3] and [bool(d[2]) for d in dones] == want[:3]
    assert [bool(d[1]) for d in dones] == [False, False, False] and env.timestep.tolist() == [3, 2, 3]
    env.goal_idx = 0
    env.reset()
    assert env.timestep.tolist() == [0, 0, 0] and log[-1] == ("reset", None)
    ext = [bool(env.step({'joint_command': torch.zeros(n, 9, dtype=torch.float64), 'render': False})[2][0]) for _ in range(2)]
    assert ext == want[3:]
