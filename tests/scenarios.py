"""Scripted contact scenarios shared by the emulated (CPU) and the GPU parity tests."""
import numpy as np

from oracle.orc import Oracle
from real_robots_b200.util import euler_to_quat

DOWN = euler_to_quat([0, 3.14, -1.57])       # gripper pointing at the table (env.py:417)


def scripted_press(x, z_from, z_to, settle=100):
    """An oracle with one cube at rest, the arm teleported above table point (x, 0) with the gripper pointing down, and the
    joint command that lowers it from z_from to z_to.  Returns (oracle, state (q, qd, obj), command[9])."""
    o = Oracle(1)
    for _ in range(settle):
        o.step(np.zeros(9))
    q0, _ = o.ik([x, 0.0, z_from], DOWN)
    q, qd, ob = o.get_state()
    q[:] = 0.0
    q[:7] = q0[:7]
    o.set_state(q, np.zeros(11), ob)
    tgt, _ = o.ik([x, 0.0, z_to], DOWN)
    cmd = np.zeros(9)
    cmd[:7] = tgt[:7]
    return o, (q, np.zeros(11), ob), cmd


def check_touch_scenarios(make_sim, steps_press=40, steps_shelf=40):
    """make_sim() -> object with set_state(q[1,11], qd[1,11], obj[1,3,13]), step(cmd[1,9]), read_state() -> (joints, touch, obj)
    as numpy, contacts() -> (hdr, pts, pairs) as numpy.  Holds the touch sensors to the oracle (robot.py:131-163):
    (1) fingers pressed onto the cube (readings of thousands of newtons).  Fingers flat on a cube are face-to-face contacts:
        the witness point of such a pair jumps on 1e-16 differences of the poses, so the two sides restart from the
        oracle's state every step and a step counts when the reading agrees to 1e-9 relative; >= 85 % of the steps must,
        among them one with a reading above 500 N (returns the fraction);
    (2) fingers pressed onto the shelf (table_upper, link 0 of the table): contacts with non-zero impulses on the skin links,
        but contact_list() only reports contacts with BASE links, so all four sensors stay 0."""
    from real_robots_b200.envs.robot import manifest
    names = [s["name"] for s in manifest()["shapes"]]
    # (1)
    o, st, cmd = scripted_press(-0.10, 0.50, 0.44)
    sim = make_sim()
    good, good_peak = 0, 0.0
    for t in range(steps_press):
        q, qd, ob = o.get_state()
        o.set_state(q, qd, ob)
        sim.set_state(q[None], qd[None], ob[None])
        o.step(cmd)
        sim.step(cmd[None])
        _, rt = o.read()
        touch = np.asarray(sim.read_state()[1])[0]
        if np.abs(touch - rt).max() <= 1e-9 * max(1.0, np.abs(rt).max()):
            good += 1
            good_peak = max(good_peak, rt.max())
    assert good >= 0.85 * steps_press and good_peak > 500.0, (good, steps_press, good_peak)
    # (2)
    o, st, cmd = scripted_press(0.12, 0.62, 0.52)
    sim = make_sim()
    sim.set_state(st[0][None], st[1][None], st[2][None])
    shelf_impulse, skin_on_shelf = 0.0, False
    for t in range(steps_shelf):
        o.step(cmd)
        sim.step(cmd[None])
        _, rt = o.read()
        touch = np.asarray(sim.read_state()[1])[0]
        assert not rt.any() and not touch.any(), (t, touch, rt)
        hdr, pts, pairs = [np.asarray(x) for x in sim.contacts()]
        for slot in range(hdr.shape[1]):
            pi, n = int(hdr[0, slot, 0]), int(hdr[0, slot, 1])
            if pi < 0:
                break
            if n and names[int(pairs[pi, 1])] == "table_upper" and int(pairs[pi, 0]) < 17:
                shelf_impulse = max(shelf_impulse, float(pts[0, slot, :n, 10].max()))
                if names[int(pairs[pi, 0])].startswith("skin_") and (np.abs(pts[0, slot, :n, 9]) < 0.1).any():
                    skin_on_shelf = True         # would be reported if table_upper were a base link (contact threshold 0.1, robot.py:47)
    assert shelf_impulse > 0.0 and skin_on_shelf           # the gripper did press on the shelf, skins within the contact threshold
    return good / steps_press
