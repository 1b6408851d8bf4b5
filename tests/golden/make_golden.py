"""Generates tests/golden/oracle_golden.npz — regression pins of the CPU oracle.

The reference's own arithmetic (pybullet) cannot be run in this container, so these vectors are
NOT reference outputs: they pin the oracle restatement so that it cannot drift silently, and give the GPU tests
fixed inputs/outputs that travel to the GPU box.   usage: python -m tests.golden.make_golden
"""
import os

import numpy as np

from oracle.orc import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
N_PHYSICS, N_RENDER = 3, 3
PHYSICS_CONTACTS = 1


def actions(k, steps=80):
    rng = np.random.default_rng(100 + k)
    lo = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0]); hi = -lo; hi[7:] = np.pi / 2
    a = rng.uniform(lo, hi)
    out = []
    for t in range(steps):
        if rng.random() < 0.05:
            a = rng.uniform(lo, hi)
        out.append(a.copy())
    return np.array(out)


def physics_case(k, contacts=PHYSICS_CONTACTS):
    o = Oracle(3, enable_contacts=contacts)
    for a in actions(k):
        o.step(a)
    return o.get_state()


def render_state(k):
    rng = np.random.default_rng(200 + k)
    q = np.zeros(11) if k == 0 else rng.uniform(-1.5, 1.5, 11)
    if k:
        q[7:] = [0.5, -0.3, 0.5, -0.3]
    obj = np.zeros((3, 13))
    for o in range(3):
        obj[o, :3] = [rng.uniform(-0.25, 0.05), rng.uniform(-0.4, 0.4), rng.uniform(0.32, 0.5)]
        qq = rng.normal(size=4); obj[o, 3:7] = qq / np.linalg.norm(qq)
    return q, obj


def render_case(k):
    o = Oracle(3, enable_contacts=0)
    if k:
        q, obj = render_state(k)
        o.set_state(q=q, qd=np.zeros(11), obj=obj)
    return o.render()


def main():
    out = dict(n_physics=N_PHYSICS, n_render=N_RENDER, physics_contacts=PHYSICS_CONTACTS)
    for k in range(N_PHYSICS):
        q, qd, obj = physics_case(k)
        out["phys_q_%d" % k] = q; out["phys_qd_%d" % k] = qd; out["phys_obj_%d" % k] = obj
    for k in range(N_RENDER):
        rgb, depth, mask = render_case(k)
        out["rgb_%d" % k] = rgb; out["depth_%d" % k] = depth; out["mask_%d" % k] = mask.astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_golden.npz"))


if __name__ == "__main__":
    main()
