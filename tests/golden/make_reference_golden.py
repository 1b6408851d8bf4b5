"""Golden vectors for the HOST-SIDE logic of the step path, produced by running the reference's own Python code.

    python tests/golden/make_reference_golden.py        # needs /root/reference; writes tests/golden/reference_host_golden.npz

The reference cannot simulate here (pybullet / pybullet_envs / gym are not installable), but everything it does in Python
around the pybullet calls can run: this script puts inert stand-ins for the missing third-party modules into sys.modules,
imports the UNMODIFIED `real_robots` package from /root/reference, and calls its functions / methods on small fake
objects that supply the values pybullet would have supplied (current joint positions, object positions, an injective fake
IK).  What is recorded is therefore the reference's arithmetic and control flow, not a re-reading of it:

  limits           Kuka.__init__ joint limits                                       robot.py:58-67
  prologue_*       REALRobotEnv.limitActionByJoint + Kuka.apply_action              env.py:314-321, robot.py:188-201
  objlimit_*       REALRobotEnv.control_objects_limits (which objects get reset)    env.py:257-264
  score_*          REALRobotEnv.evaluateGoal                                        env.py:181-200
  plan_*           REALRobotEnv.generate_plan with the fake IK                      env.py:410-459
  macro_seq_*      REALRobotEnv.step_macro / next_step sequencing                   env.py:388-408, 462-467
  cart_*           REALRobotEnv.step_cartesian request cache                        env.py:358-386
  done_*           REALRobotEnv.step_joints call order and `done` logic             env.py:326-356
  touch_*          Kuka.get_touch_sensors / get_contacts                            robot.py:131-163
  scoreobj_*       EvaluationService.add_scores / build_score_object                evaluate.py:179-201, 396-417
  spaces_*         action / observation space bounds of the env                     env.py:42-81
  quat_*           object reset orientations (getQuaternionFromEuler of object_poses)  robot.py:19-24, 125-129

tests/test_reference_golden.py (CPU) holds the oracle and the host mirror to these vectors; the GPU parity tests then
hold the CUDA path to the oracle.
"""
import json
import math
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"


def fake_ik(coords, orient=None):
    """Injective, smooth stand-in for calculateInverseKinematics: 11 values that depend on the request."""
    c = np.asarray(coords, dtype=np.float64)
    o = np.zeros(4) if orient is None else np.asarray(orient, dtype=np.float64)
    ph = c @ np.array([1.3, 2.1, 0.7]) + o @ np.array([0.11, 0.23, 0.31, 0.43])
    return tuple(np.sin(np.arange(11) * 0.37 + ph))


def quat_from_euler(e):
    """pybullet.getQuaternionFromEuler (btQuaternion::setEulerZYX, xyzw)."""
    r, p, y = [0.5 * float(v) for v in e]
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def install_stubs():
    sys.path.insert(0, REPO)
    from real_robots_b200 import spaces as shim          # gym.spaces stand-in (Box / Dict / MultiBinary with sample())
    datadir = tempfile.mkdtemp()
    os.makedirs(os.path.join(datadir, "kuka_gripper_description"))   # so that the package does not try to copy its data

    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    ik_calls = []

    def calculateInverseKinematics(body, link, coords, orient, maxNumIterations=None, residualThreshold=None):
        ik_calls.append((np.array(coords, dtype=np.float64), np.array(orient, dtype=np.float64)))
        return fake_ik(coords, orient)
    pb = mod("pybullet", getQuaternionFromEuler=quat_from_euler, calculateInverseKinematics=calculateInverseKinematics,
             ER_BULLET_HARDWARE_OPENGL=1, ER_TINY_RENDERER=2, COV_ENABLE_TINY_RENDERER=12, ik_calls=ik_calls)
    mod("pybullet_data", getDataPath=lambda: datadir)

    class MJCFBaseBulletEnv:
        def __init__(self, robot, render=False):
            self.robot, self.isRender, self.scene, self.camera = robot, render, None, None
            self.action_space, self.observation_space = robot.action_space, robot.observation_space
    mod("pybullet_envs")
    mod("pybullet_envs.env_bases", MJCFBaseBulletEnv=MJCFBaseBulletEnv)
    mod("pybullet_envs.scene_abstract", SingleRobotEmptyScene=type("SingleRobotEmptyScene", (), {}))

    class URDFBasedRobot:
        def __init__(self, model_urdf, robot_name, action_dim, obs_dim, *a, **k):
            self.model_urdf, self.robot_name = model_urdf, robot_name
    mod("robot_bases", URDFBasedRobot=URDFBasedRobot, BodyPart=type("BodyPart", (), {}))
    gym = mod("gym", spaces=shim, Env=object)
    sys.modules["gym.spaces"] = shim
    mod("gym.envs")
    registered = []
    mod("gym.envs.registration", register=lambda **kw: registered.append(kw), registered=registered)
    gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("gym.make is not available in the golden generator"))
    events = []

    class AIcrowdEvents:
        AICROWD_EVENT_INFO, AICROWD_EVENT_ERROR, AICROWD_EVENT_SUCCESS = "INFO", "ERROR", "SUCCESS"

        def register_event(self, event_type, payload):
            events.append((event_type, payload["state"]))
    api = mod("aicrowd_api", event_log=events)
    api.events = mod("aicrowd_api.events", AIcrowdEvents=AIcrowdEvents)
    mod("interval", interval=type("interval", (), {"__init__": lambda self, *a: None, "__len__": lambda self: 0}))
    sys.path.insert(0, REF)
    return pb


class NS(types.SimpleNamespace):
    pass


def main():
    pb = install_stubs()
    import real_robots                                           # the reference package, unmodified
    assert real_robots.__file__.startswith(REF), real_robots.__file__
    from real_robots.envs.env import REALRobotEnv, Goal
    from real_robots.envs.robot import Kuka
    from real_robots.evaluate import EvaluationService
    rng = np.random.default_rng(20201)
    G = {}
    meta = {"reference_version": real_robots.__version__}

    # ---- limits and spaces
    k3 = Kuka(additional_obs=True, objects=3)
    G["limits_min"], G["limits_max"] = k3.min_joints.copy(), k3.max_joints.copy()
    env = REALRobotEnv(render=False, objects=3, action_type='joints', additional_obs=True)
    G["spaces_cartesian_low"], G["spaces_cartesian_high"] = env.cartesian_space.low.astype(np.float64), env.cartesian_space.high.astype(np.float64)
    G["spaces_macro_low"], G["spaces_macro_high"] = env.macro_space.low.astype(np.float64), env.macro_space.high.astype(np.float64)
    G["spaces_gripper_low"], G["spaces_gripper_high"] = env.gripper_space.low.astype(np.float64), env.gripper_space.high.astype(np.float64)
    meta["obs_keys_R1"] = sorted(env.observation_space.spaces)
    meta["obs_keys_R2"] = sorted(REALRobotEnv(objects=1, additional_obs=False).observation_space.spaces)
    meta["obs_shapes_R1"] = {k: [list(v.shape), str(v.dtype)] for k, v in env.observation_space.spaces.items() if getattr(v, "shape", None)}
    meta["timesteps"] = [int(env.intrinsic_timesteps), int(env.extrinsic_timesteps), int(env.extrinsic_trials)]
    meta["used_objects"] = {str(n): Kuka(True, n).used_objects for n in (1, 2, 3)}
    meta["robot_position"] = list(k3.robot_position)
    meta["contact_threshold"] = k3.contact_threshold
    names = ["table", "cube", "tomato", "mustard"]
    G["quat_object_poses"] = np.array([list(Kuka.object_poses[n][:3]) + list(quat_from_euler(Kuka.object_poses[n][3:])) for n in names])

    # ---- action prologue: limitActionByJoint then apply_action; q is the 11-dof state in the order
    #      [iiwa 1..7, base_to_finger00, finger00_to_finger01, base_to_finger10, finger10_to_finger11]
    order = ["lbr_iiwa_joint_%d" % (i + 1) for i in range(7)] + ["base_to_finger00_joint", "finger00_to_finger01_joint",
                                                                  "base_to_finger10_joint", "finger10_to_finger11_joint"]

    class FakeJoint:
        def __init__(self, pos):
            self.pos, self.target = pos, None

        def get_position(self):
            return self.pos

        def set_position(self, t):
            self.target = float(t)
    n = 64
    q = np.zeros((n, 11))
    q[:, :7] = rng.uniform(-2.0, 2.0, (n, 7))
    q[:, 7] = rng.uniform(0, 1.5, n); q[:, 9] = q[:, 7]
    q[:, 8] = -rng.uniform(0, 1.5, n); q[:, 10] = q[:, 8]
    q[0] = 0.0
    cmd = rng.uniform(-3.5, 3.5, (n, 9))
    cmd[:, 7:] = rng.uniform(-0.5, 2.0, (n, 2))
    cmd[1] = q[1, [0, 1, 2, 3, 4, 5, 6, 7, 8]] * [1, 1, 1, 1, 1, 1, 1, 1, -1]          # request == current state
    targets = np.zeros((n, 11))
    for i in range(n):
        robot = Kuka(True, 3)
        robot.jdict = {name: FakeJoint(q[i, j]) for j, name in enumerate(order)}
        fake_env = NS(robot=robot)
        limited = REALRobotEnv.limitActionByJoint(fake_env, cmd[i].copy())
        robot.apply_action(np.array(limited))
        targets[i] = [robot.jdict[name].target for name in order]
    G["prologue_q"], G["prologue_cmd"], G["prologue_targets"] = q, cmd, targets

    # ---- control_objects_limits
    pos = rng.uniform([-0.4, -0.5, 0.0], [0.4, 0.5, 0.6], (48, 3, 3))
    pos[0] = [[0.12, 0, 0.28], [0.10, 0, 0.28], [0.0, 0.0, 0.079]]                     # edges of the two conditions
    pos[1] = [[0.11, 0, 0.28], [0.12, 0, 0.29], [0.0, 0.0, 0.08]]
    reset_flags = np.zeros((48, 3), dtype=np.int32)
    objs = ["cube", "tomato", "mustard"]
    for i in range(48):
        hit = []
        bodies = {"table": NS(get_position=lambda: (0.0, 0.0, 0.08))}
        for o, name in enumerate(objs):
            bodies[name] = NS(get_position=(lambda p: (lambda: tuple(p)))(pos[i, o]))
        robot = NS(used_objects=["table"] + objs, object_bodies=bodies, object_poses=Kuka.object_poses, reset_object=hit.append)
        REALRobotEnv.control_objects_limits(NS(robot=robot))
        assert "table" not in hit
        reset_flags[i] = [int(name in hit) for name in objs]
    G["objlimit_pos"], G["objlimit_reset"] = pos, reset_flags

    # ---- evaluateGoal
    cur = rng.uniform([-0.3, -0.5, 0.25], [0.3, 0.5, 0.55], (32, 3, 3))
    goal = cur + rng.normal(0, 0.08, (32, 3, 3))
    goal[0] = cur[0]
    goal[1] = cur[1] + [[0.10, 0, 0], [0, 0.10, 0], [0, 0, 0.10]]
    scores = np.zeros(32)
    for i in range(32):
        bodies = {name: NS(get_position=(lambda p: (lambda: tuple(p)))(cur[i, o])) for o, name in enumerate(objs)}
        g = Goal(initial_state={}, final_state={name: np.concatenate([goal[i, o], [0, 0, 0, 1]]) for o, name in enumerate(objs)}, challenge='2D')
        ch, s = REALRobotEnv.evaluateGoal(NS(goal=g, robot=NS(object_bodies=bodies)))
        assert ch == '2D'
        scores[i] = s
    G["score_cur"], G["score_goal"], G["score_value"] = cur, goal, scores
    ef = []
    for i in range(8):
        a, b = rng.normal(size=4), rng.normal(size=4)
        a /= np.linalg.norm(a); b /= np.linalg.norm(b)
        ef.append(list(cur[i, 0]) + list(goal[i, 0]) + list(a) + list(b) +
                  [REALRobotEnv.extrinsicFormula(None, goal[i, 0], cur[i, 0], b, a, w) for w in (1, 0.5, 0.0)])
    G["extrinsic_formula"] = np.array(ef)

    # ---- generate_plan with the fake IK
    macros = np.array([[[-0.10, -0.30], [-0.05, 0.20]], [[-0.25, -0.5], [0.05, 0.5]], [[0.0, 0.0], [0.0, 0.0]],
                       [[-0.2, 0.1], [-0.2, 0.149]], [[-0.2, 0.1], [-0.2, 0.151]], [[0.05, -0.5], [-0.25, 0.5]]])
    plans, ncalls = [], []
    for m in macros:
        e = NS()
        pb.ik_calls.clear()
        REALRobotEnv.generate_plan(e, m)
        assert e.plan_step == -1
        plans.append(np.array(e.planned_actions))
        ncalls.append(len(pb.ik_calls))
    G["plan_macro"], G["plan_rows"], G["plan_ik_calls"] = macros, np.array(plans), np.array(ncalls)

    # ---- step_macro sequencing: same action -> next row; new action or exhausted plan -> new plan
    seq_actions = [macros[0]] * 3 + [macros[1]] * 2 + [macros[0]] * 1001
    sent = []
    e = NS(requested_action=None)
    e.generate_plan = lambda m: REALRobotEnv.generate_plan(e, m)
    e.next_step = lambda: REALRobotEnv.next_step(e)
    e.step_joints = lambda a: sent.append(np.array(a["joint_command"]))
    steps_of_plan = []
    for a in seq_actions:
        REALRobotEnv.step_macro(e, {"macro_action": a, "render": False})
        steps_of_plan.append(e.plan_step)
    G["macro_seq_action_index"] = np.array([0] * 3 + [1] * 2 + [0] * 1001)
    G["macro_seq_plan_step"] = np.array(steps_of_plan)
    G["macro_seq_first_rows"] = np.array(sent[:8])
    G["macro_seq_last_rows"] = np.array(sent[-3:])
    e2 = NS(requested_action=None, step_joints=lambda a: sent.append(a))
    REALRobotEnv.step_macro(e2, {"macro_action": None, "render": True})
    assert np.all(sent[-1]["joint_command"] == 0) and sent[-1]["render"] is True

    # ---- step_cartesian request cache
    carts = [np.array([-0.1, 0.1, 0.5, 0, 1, 0, 0.]), np.array([-0.1, 0.1, 0.5, 0, 1, 0, 0.]), np.array([-0.1, 0.1, 0.5, 0, 1, 0, 0.1]),
             np.array([-0.1, 0.1, 0.5, 0, 1, 0, 0.1]), np.array([-0.1, 0.1, 0.45, 0, 1, 0, 0.1]), np.array([-0.1, 0.1, 0.5, 0, 1, 0, 0.])]
    grips = rng.uniform(0, np.pi / 2, (len(carts), 2))
    sent2, counts = [], []
    e = NS(requested_coords=None, requested_orient=None, last_ik=None, step_joints=lambda a: sent2.append(np.array(a["joint_command"])))
    pb.ik_calls.clear()
    for c, gcmd in zip(carts, grips):
        REALRobotEnv.step_cartesian(e, {"cartesian_command": c, "gripper_command": gcmd, "render": False})
        counts.append(len(pb.ik_calls))
    G["cart_cmd"], G["cart_grip"], G["cart_ik_calls"], G["cart_joint_cmd"] = np.array(carts), grips, np.array(counts), np.array(sent2)

    # ---- step_joints: order of operations and `done`
    log = []
    robot = NS(calc_state=lambda: [0.0] * 9, used_objects=["table"], object_bodies={"table": NS(get_position=lambda: (0, 0, 0.08))},
               object_poses=Kuka.object_poses, reset_object=lambda o: log.append("reset_object"),
               apply_action=lambda a: log.append("apply_action"))
    e = NS(robot=robot, scene=NS(multiplayer=False, global_step=lambda: log.append("global_step")), timestep=0, goal_idx=-1,
           intrinsic_timesteps=3, extrinsic_timesteps=2, reward_func=lambda o: 0)
    e.limitActionByJoint = lambda a: (log.append("limitActionByJoint"), REALRobotEnv.limitActionByJoint(e, a))[1]
    e.control_objects_limits = lambda: (log.append("control_objects_limits"), REALRobotEnv.control_objects_limits(e))[1]
    e.get_observation = lambda cam: (log.append("get_observation(%s)" % bool(cam)), {})[1]
    dones = []
    for t in range(3):
        o, r, d, info = REALRobotEnv.step_joints(e, {"joint_command": None if t == 0 else np.zeros(9), "render": t == 1})
        dones.append(bool(d))
        assert r == 0 and info == {}
    e.goal_idx, e.timestep = 0, 0
    for t in range(2):
        dones.append(bool(REALRobotEnv.step_joints(e, {"joint_command": np.zeros(9), "render": False})[2]))
    meta["step_joints_call_order"] = log[:5]
    meta["done_sequence_intrinsic3_then_extrinsic2"] = dones

    # ---- touch sensors from contact lists (contact tuple: [8] distance, [9] normal force, [2] body id of the other side)
    def contact(body, dist, force):
        c = [0] * 14
        c[2], c[8], c[9] = body, dist, force
        return c
    parts = {"skin_00": NS(contact_list=lambda: [contact(2, 0.01, 3.5), contact(3, -0.02, 7.25), contact(2, 0.2, 99.0)]),
             "skin_01": NS(contact_list=lambda: []),
             "skin_10": NS(contact_list=lambda: [contact(1, 0.099, 0.5)]),
             "skin_11": NS(contact_list=lambda: [contact(1, 0.1, 5.0)]),
             "finger_00": NS(contact_list=lambda: [contact(2, 0.0, 11.0)])}
    kk = Kuka(True, 3)
    kk.parts = parts
    kk.object_names = {0: "kuka", 1: "table", 2: "cube", 3: "tomato", 4: "mustard"}
    G["touch_sensors"] = kk.get_touch_sensors()
    meta["touch_contacts"] = {k: sorted(v) for k, v in kk.get_contacts().items()}

    # ---- evaluation score object
    svc = NS(scores={}, evaluation_state={}, sync_evaluation_state=lambda: None)
    for ch, s in [("2D", 0.5), ("2D", 0.25), ("3D", 1.0), ("2.5D", 0.125), ("2D", 0.75)]:
        EvaluationService.add_scores(svc, ch, s)
    so = EvaluationService.build_score_object(svc)
    meta["scoreobj"] = {k: float(v) for k, v in so.items()}
    svc2 = NS(scores={}, evaluation_state={}, sync_evaluation_state=lambda: None)
    EvaluationService.add_scores(svc2, "2D", 0.5)
    meta["scoreobj_single"] = {k: float(v) for k, v in EvaluationService.build_score_object(svc2).items()}

    # ---- EvaluationService driven end to end with a deterministic stand-in env and a recording controller
    sys.path.insert(0, HERE)
    import fake_env
    import gym
    from real_robots.policy import BasePolicy
    log = []
    fake_env.FakeEnv.log = log
    made = []
    gym.make = lambda env_id: (made.append(env_id), fake_env.FakeEnv())[1]
    svc = EvaluationService(fake_env.make_controller(BasePolicy, log), environment='R1', action_type='macro_action', n_objects=2,
                            intrinsic_timesteps=3, extrinsic_timesteps=2, extrinsic_trials=4, visualize=True,
                            goals_dataset_path="goals.npz", video=None)
    svc.run_intrinsic_phase()
    so = svc.run_extrinsic_phase()
    meta["eval_env_id"] = made
    meta["eval_log"] = log
    meta["eval_score_object"] = fake_env.plain(so)
    meta["eval_scores"] = fake_env.plain(svc.scores)
    meta["eval_state"] = fake_env.plain(svc.evaluation_state)
    # skipped intrinsic phase
    log2 = []
    fake_env.FakeEnv.log = log2
    svc = EvaluationService(fake_env.make_controller(BasePolicy, log2), 'R2', 'joints', 1, 0, 1, 1, False, "g.npz", None)
    svc.run_intrinsic_phase()
    meta["eval_skip_state"] = fake_env.plain(svc.evaluation_state)
    meta["eval_skip_log"] = list(log2)
    errors = {}
    for args in (("R3", "joints", 1), ("R2", "macro_action", 1), ("R1", "bogus", 1), ("R1", "joints", 4), ("R1", "joints", "2")):
        try:
            EvaluationService(fake_env.make_controller(BasePolicy, []), *args, 1, 1, 1, False, "g.npz", None)
            errors[str(args)] = None
        except Exception as ex:
            errors[str(args)] = str(ex)
    try:
        EvaluationService(object, 'R1', 'joints', 1, 1, 1, 1, False, "g.npz", None)
        errors["not_a_policy"] = None
    except Exception as ex:
        errors["not_a_policy"] = type(ex).__name__
    meta["eval_errors"] = errors

    # ---- settle detector: runEnv (generate_goals.py:21-68) on scripted object motion
    import contextlib, io
    from real_robots import generate_goals as ref_gg
    settle = {}
    for script in fake_env.SCRIPTS:
        for max_t in fake_env.MAX_TS:
            st = NS(s=0)
            e = NS(robot=NS(used_objects=["table", "cube"]))
            e.get_obj_pose = lambda obj, st=st, script=script: fake_env.scripted_pose(script, st.s)
            e.step = lambda action, st=st: (setattr(st, "s", st.s + 1), ({"retina": 0, "mask": 0}, 0, False, {}))[1]
            with contextlib.redirect_stdout(io.StringIO()):
                _, pos, failed, t, _ = ref_gg.runEnv(e, max_t=max_t)
            settle["%s/%d" % (script, max_t)] = [bool(failed), int(t)]
    meta["settle"] = settle
    # generatePosition (:79-108) with the legacy global RNG: x, y, z and (free orientation) the quaternion of 3 Euler draws
    draws = []
    fake_p = NS(getQuaternionFromEuler=quat_from_euler)
    for seed, plane in ((1, None), (2, True), (3, False), (4, True), (5, None)):
        np.random.seed(seed)
        draws.append([ref_gg.generatePosition(NS(_p=fake_p), "cube", fixed=False, tablePlane=plane) for _ in range(4)])
    G["genpos_free"] = np.array(draws)
    meta["shelf_table_thresholds"] = {obj: [float(z) for z in (0.30, 0.329, 0.331, 0.339, 0.341, 0.394, 0.396, 0.399, 0.401, 0.45)
                                            if ref_gg.isOnTable(obj, {obj: [0, 0, z]})] +
                                           [-float(z) for z in (0.30, 0.329, 0.331, 0.339, 0.341, 0.394, 0.396, 0.399, 0.401, 0.45)
                                            if ref_gg.isOnShelf(obj, {obj: [0, 0, z]})] for obj in ("cube", "tomato", "mustard")}

    # ---- registry (real_robots/__init__.py:16-28) and the BasePolicy interface (policy.py)
    import inspect
    import gym.envs.registration as reg
    meta["registry"] = {r["id"]: {"entry_point": r["entry_point"], "kwargs": r["kwargs"]} for r in reg.registered}
    meta["base_policy"] = {n: str(inspect.signature(f)) for n, f in inspect.getmembers(BasePolicy, inspect.isfunction)}

    G["meta_json"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8)
    out = os.path.join(HERE, "reference_host_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, "with", len(G), "arrays;", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
