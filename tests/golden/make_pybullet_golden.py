#!/usr/bin/env python
"""SURVEY App. D dump — run ONCE on a machine where `import pybullet, pybullet_envs, gym` works:

    python tests/golden/make_pybullet_golden.py --ref /path/to/real_robots [--out tests/golden/pybullet_golden.npz]

It drives the UNMODIFIED reference (`gym.make('REALRobot2020-R1J3-v0')`) and records what only pybullet can say about
the arithmetic under the step path: engine parameters, per-link dynamics info, joint order, a contact-free replay,
contact replays, renders and IK answers.  `tests/test_pybullet_golden.py` holds the oracle (CPU) and the CUDA path
(GPU) to the file when it exists; until then both tests skip and DESIGN.md §2 keeps saying PARITY UNPINNED.

NOT RUN in the build container: pybullet / gym are absent there, so this script is untested against a
live pybullet.  It only uses documented pybullet calls through `env._p` and the reference's public env surface.
"""
import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "pybullet_golden.npz"))
    ap.add_argument("--steps", type=int, default=1000)
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    import gym
    import pybullet
    import real_robots  # noqa: F401  (registers the ids)

    out = {}
    out["pybullet_api_version"] = np.array(pybullet.getAPIVersion())
    out["numpy_enabled"] = np.array(int(pybullet.isNumpyEnabled()))

    env = gym.make('REALRobot2020-R1J3-v0')
    env = getattr(env, "unwrapped", env)
    env.reset()
    p = env._p
    n_bodies = p.getNumBodies()
    assert n_bodies == 5, n_bodies                                       # kuka, table, cube, tomato, mustard

    # (2) engine parameters
    params = p.getPhysicsEngineParameters()
    for k, v in params.items():
        if isinstance(v, (int, float)):
            out["engine_" + k] = np.array(float(v))

    # (3) dynamics / joint info per body and link
    for b in range(n_bodies):
        nj = p.getNumJoints(b)
        out["body%d_name" % b] = np.array(p.getBodyInfo(b)[1].decode())
        out["body%d_base_name" % b] = np.array(p.getBodyInfo(b)[0].decode())
        dyn = [p.getDynamicsInfo(b, l) for l in range(-1, nj)]
        out["body%d_mass" % b] = np.array([d[0] for d in dyn])
        out["body%d_lateral_friction" % b] = np.array([d[1] for d in dyn])
        out["body%d_inertia_diag" % b] = np.array([d[2] for d in dyn])
        out["body%d_inertial_pos" % b] = np.array([d[3] for d in dyn])
        out["body%d_inertial_orn" % b] = np.array([d[4] for d in dyn])
        out["body%d_restitution" % b] = np.array([d[5] for d in dyn])
        out["body%d_rolling_friction" % b] = np.array([d[6] for d in dyn])
        out["body%d_spinning_friction" % b] = np.array([d[7] for d in dyn])
        if nj:
            ji = [p.getJointInfo(b, j) for j in range(nj)]
            out["body%d_joint_names" % b] = np.array([j[1].decode() for j in ji])
            out["body%d_link_names" % b] = np.array([j[12].decode() for j in ji])
            out["body%d_joint_type" % b] = np.array([j[2] for j in ji])
            out["body%d_joint_damping" % b] = np.array([j[6] for j in ji])
            out["body%d_joint_lower" % b] = np.array([j[8] for j in ji])
            out["body%d_joint_upper" % b] = np.array([j[9] for j in ji])
            out["body%d_joint_max_force" % b] = np.array([j[10] for j in ji])
            out["body%d_joint_axis" % b] = np.array([j[13] for j in ji])
            out["body%d_parent_frame_pos" % b] = np.array([j[14] for j in ji])
            out["body%d_parent_index" % b] = np.array([j[16] for j in ji])
        vis = p.getVisualShapeData(b)
        out["body%d_visual_link" % b] = np.array([v[1] for v in vis])
        out["body%d_visual_rgba" % b] = np.array([v[7] for v in vis])

    def robot_state():
        js = p.getJointStates(0, list(range(p.getNumJoints(0))))
        return np.array([s[0] for s in js]), np.array([s[1] for s in js]), np.array([s[3] for s in js])

    def object_state():
        rows = []
        for b in range(2, n_bodies):
            pos, orn = p.getBasePositionAndOrientation(b)
            lin, ang = p.getBaseVelocity(b)
            rows.append(list(pos) + list(orn) + list(lin) + list(ang))
        return np.array(rows)

    def link_poses():
        ls = p.getLinkStates(0, list(range(p.getNumJoints(0))))
        return np.array([list(s[0]) + list(s[1]) for s in ls])           # COM frames, xyz + quat xyzw

    # (4) contact-free replay: objects lifted far above the table every step (tests/test_actions.py:94-97 idiom)
    rng = np.random.RandomState(2020)
    low, high = env.robot.action_space.low, env.robot.action_space.high      # robot.py:58-71
    env.reset()
    acts, qs, qds, taus, links = [], [], [], [], []
    a = rng.uniform(low, high)
    for t in range(args.steps):
        if rng.rand() < 0.05:
            a = rng.uniform(low, high)
        for i, name in enumerate(env.robot.used_objects[1:]):
            env.robot.object_bodies[name].reset_pose([2.0 + i, 2.0, 5.0], [0, 0, 0, 1])
        env.step({'joint_command': a, 'render': False})
        q, qd, tau = robot_state()
        acts.append(a.copy()); qs.append(q); qds.append(qd); taus.append(tau); links.append(link_poses())
    out["free_actions"] = np.array(acts); out["free_q"] = np.array(qs); out["free_qd"] = np.array(qds)
    out["free_tau"] = np.array(taus); out["free_link_poses"] = np.array(links)

    # (5) contact replay: drop and settle (first 300 steps after reset, zero command)
    env.reset()
    objs, ncont, cforce = [], [], []
    for t in range(300):
        env.step({'joint_command': np.zeros(9), 'render': False})
        objs.append(object_state())
        cps = [c for b in range(2, n_bodies) for c in p.getContactPoints(bodyA=b)]
        ncont.append(len(cps)); cforce.append(sum(c[9] for c in cps))
    out["settle_objects"] = np.array(objs); out["settle_n_contacts"] = np.array(ncont)
    out["settle_normal_force_sum"] = np.array(cforce)
    out["settle_touch"] = np.array(env.robot.get_touch_sensors())

    # (6) renders at several arm poses (objects resting on the table)
    poses = [np.zeros(9)] + [rng.uniform(low, high) * 0.6 for _ in range(7)]
    rq, robj, rrgb, rdepth, rmask = [], [], [], [], []
    for target in poses:
        for _ in range(60):
            obs, _, _, _ = env.step({'joint_command': target, 'render': False})
        obs, _, _, _ = env.step({'joint_command': target, 'render': True})
        q, qd, _ = robot_state()
        rq.append(np.stack([q, qd])); robj.append(object_state())
        rrgb.append(obs['retina']); rdepth.append(np.asarray(obs['depth'], np.float32)); rmask.append(obs['mask'])
    out["render_robot"] = np.array(rq); out["render_objects"] = np.array(robj)
    out["render_rgb"] = np.array(rrgb, np.uint8); out["render_depth"] = np.array(rdepth, np.float32)
    out["render_mask"] = np.array(rmask, np.int32)
    out["view_matrix"] = np.array(pybullet.computeViewMatrix([0.01, 0, 1.2], [0, 0, 0.08], [0, 0, 1]), np.float32)
    out["proj_matrix"] = np.array(pybullet.computeProjectionMatrixFOV(80, 320.0 / 240.0, 0.1, 100.0), np.float32)

    # (7) IK from the home pose, as env.py:372-375 calls it
    env.reset()
    targets, answers = [], []
    for _ in range(100):
        xyz = rng.uniform([-0.25, -0.5, 0.40], [0.05, 0.5, 0.60])
        quat = rng.uniform(-1, 1, 4)
        sol = pybullet.calculateInverseKinematics(0, 7, xyz, quat, maxNumIterations=1000, residualThreshold=0.001)
        targets.append(np.concatenate([xyz, quat])); answers.append(np.array(sol))
    out["ik_targets"] = np.array(targets); out["ik_answers"] = np.array(answers)

    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "with", len(out), "arrays; pybullet API", int(out["pybullet_api_version"]))


if __name__ == "__main__":
    main()
