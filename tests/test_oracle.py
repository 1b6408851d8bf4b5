"""CPU tests of the oracle (the parity anchor): known answers from the reference's own test script, an
independent numpy cross-check of the Featherstone ABA, analytic rigid-body cases, renderer invariants, and the
committed golden vectors (tests/golden, regression pins produced by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle.orc import Oracle, eye_matrices
from tests.np_dynamics import Arm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fk_known_answers_from_reference_test_script():
    # /root/reference/tests/test_actions.py:60,66: home -> (-0.55, 0, 1.27); home2 (q6=q7=pi/2) -> (-0.41, 0, 1.14), tol 0.01 (:69)
    o = Oracle(3, enable_contacts=0)
    assert np.allclose(o.link_pose(7)[:3], [-0.55, 0, 1.27], atol=0.01)
    q = np.zeros(11); q[5] = q[6] = np.pi / 2
    o.set_state(q=q)
    assert np.allclose(o.link_pose(7)[:3], [-0.41, 0, 1.14], atol=0.01)
    # exact URDF kinematics (SURVEY §8c known answers)
    assert np.allclose(o.link_pose(7)[:3], [-0.419, 0, 1.140], atol=1e-9)


def test_fk_matches_independent_numpy(blob):
    arm = Arm(blob)
    rng = np.random.default_rng(0)
    o = Oracle(3, enable_contacts=0)
    for _ in range(5):
        q = rng.uniform(-2, 2, 11)
        o.set_state(q=q)
        R, p, _ = arm.fk(q)
        for link in range(16):
            assert np.allclose(o.link_pose(link)[:3], p[link], atol=1e-12)


def test_aba_matches_mass_matrix_and_bias(blob):
    """qdd_ABA == M^-1 (tau - C qd - g) with M, C, g from an independent numpy formulation (damping off)."""
    arm = Arm(blob)
    rng = np.random.default_rng(1)
    o = Oracle(3, enable_contacts=0, lin_damping=0.0, ang_damping=0.0)
    for _ in range(4):
        q = rng.uniform(-1.5, 1.5, 11); qd = rng.uniform(-1, 1, 11); tau = rng.uniform(-5, 5, 11)
        o.set_state(q=q, qd=qd)
        M = arm.mass_matrix(q)
        Minv = o.minv()
        assert np.allclose(Minv @ M, np.eye(11), atol=1e-9)
        qdd = o.aba(tau)
        ref = np.linalg.solve(M, tau + arm.gravity_torque(q) - arm.coriolis(q, qd))
        assert np.allclose(qdd, ref, rtol=1e-5, atol=1e-5)


def test_free_fall_and_motor_tracking():
    o = Oracle(3, enable_contacts=0)
    # free fall of the cube: v = g t (1 - small damping), z decreases
    for _ in range(10):
        o.step(np.zeros(9))
    _, _, obj = o.get_state()
    assert -0.5 < obj[0, 9] < -0.45           # ~ -9.81 * 0.05 with 0.04(1+|v|) damping
    # POSITION_CONTROL closes ~10% of the error per step: after 200 steps the arm is at the (rate-limited) target
    tgt = np.array([0.5, 0.3, -0.4, 0.6, 0.2, -0.3, 0.1, 0.4, 0.2])
    for _ in range(400):
        o.step(tgt)
    j, _ = o.read()
    assert np.allclose(j, tgt, atol=2e-2)


def test_action_prologue_matches_reference_formulae():
    # env.py:314-321 (per-step clip .2 .2 .2 .2 .2 .3 .3 .1 .1) and robot.py:192-193 (range clip, a8 <= 2 a7)
    o = Oracle(3, enable_contacts=0)
    o.step(np.array([5, -5, 5, -5, 5, -5, 5, 5, 5.0]))
    # after one step targets are cur + clip(diff): observable through the motor: q moves by ~0.1 * target
    q, _, _ = o.get_state()
    exp_t = np.array([0.2, -0.2, 0.2, -0.2, 0.2, -0.3, 0.3, 0.1, 0.1])
    # targets fan out 9 -> 11 with sign flip on the distal finger joints (robot.py:198-201)
    exp11 = np.concatenate([exp_t[:7], [exp_t[7], -exp_t[8], exp_t[7], -exp_t[8]]])
    assert np.all(np.sign(q[np.abs(exp11) > 0]) == np.sign(exp11[np.abs(exp11) > 0]))
    assert np.allclose(q, 0.1 * exp11, atol=0.02)


def test_teleport_out_of_bounds_objects():
    # env.py:257-264: z < 0.08 or (x > 0.11 and z < 0.29) -> back to object_poses, velocities zeroed
    o = Oracle(3, enable_contacts=0)
    _, _, obj = o.get_state()
    obj[0, :3] = [0.2, 0.0, 0.2]; obj[0, 7:10] = [1, 2, 3]
    obj[1, 2] = 0.05
    o.set_state(obj=obj)
    o.step(np.zeros(9))
    _, _, obj2 = o.get_state()
    assert np.allclose(obj2[0, :2], [-0.10, 0.0]) and abs(obj2[0, 2] - 0.45) < 1e-3
    assert np.allclose(obj2[1, :2], [-0.10, -0.30]) and abs(obj2[1, 2] - 0.45) < 1e-3


def test_camera_matrices_match_survey():
    v, p = eye_matrices()
    V = v.reshape(4, 4).T; P = p.reshape(4, 4).T
    assert np.allclose(V[1], [-0.99996, 0, 0.008928, -0.000714], atol=1e-5)
    assert np.allclose(V[2], [0.008928, 0, 0.99996, -1.200041], atol=1e-5)
    assert np.allclose([P[0, 0], P[1, 1], P[2, 2], P[2, 3]], [0.893815, 1.191754, -1.002002, -0.2002], atol=1e-5)


def test_render_invariants():
    o = Oracle(3, enable_contacts=0)
    rgb, depth, mask = o.render()
    assert rgb.shape == (240, 320, 3) and rgb.dtype == np.uint8 and depth.dtype == np.float32 and mask.dtype == np.int32
    ids = set(np.unique(mask).tolist())
    assert ids == {-1, 0, 1, 2, 3, 4}          # background + robot, table, cube, tomato, mustard
    assert (rgb[mask == -1] == 255).all()      # white clear colour
    assert depth.max() <= 1.0011 and depth.min() > 0.5
    # table projection: x in [82,237], y in [78,163]; robot base near (160,57)
    ys, xs = np.nonzero(mask == 1)
    assert 65 <= xs.min() <= 90 and 230 <= xs.max() <= 255 and 70 <= ys.min() <= 85 and 155 <= ys.max() <= 175
    assert mask[57, 160] == 0
    # fewer objects -> their ids disappear
    o1 = Oracle(1, enable_contacts=0)
    assert set(np.unique(o1.render()[2]).tolist()) == {-1, 0, 1, 2}


def test_scoring_formula():
    # env.py:192-194: 0.25 at 10 cm
    assert abs(np.exp(-(-np.log(0.25) / 0.10) * 0.10) - 0.25) < 1e-12


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "oracle_golden.npz")), reason="golden file missing")
def test_golden_vectors():
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    from tests.golden.make_golden import physics_case, render_case
    for k in range(int(g["n_physics"])):
        q, qd, obj = physics_case(k, contacts=int(g["physics_contacts"]))
        assert np.allclose(q, g["phys_q_%d" % k], atol=1e-12) and np.allclose(obj, g["phys_obj_%d" % k], atol=1e-12)
    for k in range(int(g["n_render"])):
        rgb, depth, mask = render_case(k)
        assert (mask == g["mask_%d" % k]).all() and (rgb == g["rgb_%d" % k]).all() and (depth == g["depth_%d" % k]).all()
