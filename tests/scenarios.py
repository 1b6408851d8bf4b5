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


def integration_stub_client(lib_path, n_objects=1):
    """Executes the ctypes binding INTEGRATION.md section B shows a maintainer of the reference (the B200Client class, verbatim
    from the document) against the library at lib_path; returns a client instance."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"## B\..*?```python\n(.*?)```", text, re.S).group(1)
    assert 'C.CDLL("libreal_b200.so")' in block
    ns = {}
    exec(block.replace('C.CDLL("libreal_b200.so")', "C.CDLL(%r)" % lib_path), ns)
    blob = os.path.join(root, "real_robots_b200", "assets", "scene.rrb")
    return ns["B200Client"](n_objects, blob=blob)


def contact_rich_tracking(step_fn, state_fn, set_state_fn, N=32, T=400, seed=3):
    """Random-policy episodes with arm / object / table contacts, CUDA path against the oracle.

    step_fn(actions[N,9]), state_fn() -> (q[N,11], qd, obj[N,3,13]) numpy, set_state_fn(q, qd, obj).
    Returns (tracked, agree, flips): `tracked` = fraction of envs whose FREE-RUNNING joints stay within 1e-6 of the oracle's
    for T steps; then, restarting both sides from the oracle's state every step, `agree` = fraction of (env, step) pairs that
    agree to 1e-9 after one step, and `flips` = the largest one-step difference among the others.  A pair that disagrees is a
    discrete narrowphase decision taken differently (witness point of a face-to-face contact, EPA face of coplanar
    candidates) on poses that differ in the last bits — not accumulated rounding: there is nothing in between."""
    lo = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0]); hi = -lo; hi[7:] = np.pi / 2
    rng = np.random.default_rng(seed)

    def policy(steps):
        a = rng.uniform(lo, hi, size=(N, 9))
        for _ in range(steps):
            redraw = rng.random(N) < 0.05
            a = np.where(redraw[:, None], rng.uniform(lo, hi, size=(N, 9)), a)
            yield a.copy()
    orcs = [Oracle(3) for _ in range(N)]
    tracked = np.ones(N, bool)
    for a in policy(T):
        step_fn(a)
        for e in range(N):
            orcs[e].step(a[e])
        q = state_fn()[0]
        oq = np.array([o.get_state()[0] for o in orcs])
        tracked &= np.abs(q - oq).max(axis=1) < 1e-6
    good, total, flips = 0, 0, 0.0
    for a in policy(T // 4):
        st = [o.get_state() for o in orcs]
        for e in range(N):
            orcs[e].set_state(*st[e])
        set_state_fn(np.stack([s[0] for s in st]), np.stack([s[1] for s in st]), np.stack([s[2] for s in st]))
        step_fn(a)
        for e in range(N):
            orcs[e].step(a[e])
        q, _, obj = state_fn()
        oq = np.array([o.get_state()[0] for o in orcs]); oobj = np.array([o.get_state()[2] for o in orcs])
        err = np.maximum(np.abs(q - oq).max(axis=1), np.abs(obj - oobj).reshape(N, -1).max(axis=1))
        good += int((err < 1e-9).sum()); total += N
        if (err >= 1e-9).any():
            flips = max(flips, float(err[err >= 1e-9].max()))
    return float(tracked.mean()), good / total, flips
