#!/usr/bin/env python
"""How much does each UNPINNED choice of the oracle move the results?

Runs the CPU oracle (test infrastructure) with one config field changed at a time and reports, against the default
configuration: the largest joint-angle difference over a 1000-step contact-free random-policy rollout (the §8d
policy), and the difference of the objects' resting poses / the number of steps they need to settle after reset.
CPU only; prints a markdown table (committed as profiles/r01_oracle_sensitivity.md).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.orc import Oracle  # noqa: E402

LO = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0])
HI = -LO
HI[7:] = np.pi / 2


def actions(steps, seed=1234):
    rng = np.random.default_rng(seed)
    a = rng.uniform(LO, HI)
    out = []
    for _ in range(steps):
        if rng.random() < 0.05:
            a = rng.uniform(LO, HI)
        out.append(a.copy())
    return out


def free_rollout(cfg, acts):
    o = Oracle(3, enable_contacts=0, **cfg)
    qs = []
    for a in acts:
        o.step(a)
        qs.append(o.get_state()[0].copy())
    return np.array(qs)


def settle(cfg, steps=400):
    o = Oracle(3, **cfg)
    prev = o.get_state()[2][:, :7].copy()
    still, settled_at = 0, None
    for t in range(steps):
        o.step(np.zeros(9))
        cur = o.get_state()[2][:, :7].copy()
        ok = np.abs(cur[:, :3] - prev[:, :3]).max() < 1e-4 and np.abs(cur[:, 3:] - prev[:, 3:]).max() < 1e-3   # generate_goals.py:21-57
        still = still + 1 if ok else 0
        if still >= 21 and settled_at is None:
            settled_at = t + 1
        prev = cur
    return o.get_state()[2][:, :7].copy(), settled_at


VARIANTS = [
    ("default", {}),
    ("solver_iterations 5 (setPhysicsEngineParameter survives resetSimulation, App. C-1)", {"solver_iterations": 5}),
    ("erp2 (contact ERP) 0.9 (App. C-1)", {"erp2": 0.9}),
    ("URDF inertias for the robot instead of AABB boxes (App. C-2)", {"use_box_inertia": 0}),
    ("motor max force 1e3 instead of 1e5", {"motor_max_force": 1000.0}),
    ("motor kp 0.2 instead of 0.1 (App. B.3)", {"motor_kp": 0.2}),
    ("no linear / angular damping on free bodies", {"lin_damping": 0.0, "ang_damping": 0.0}),
    ("warm starting factor 0.85 instead of 0.1", {"warmstart": 0.85}),
]


def main():
    acts = actions(1000)
    base_q = free_rollout({}, acts)
    base_pose, base_t = settle({})
    print("| changed field | max abs dq over 1000 contact-free steps (rad) | max dq at step 1 | rest position shift (m) | steps to settle |")
    print("|---|---|---|---|---|")
    for name, cfg in VARIANTS:
        q = free_rollout(cfg, acts)
        pose, t = settle(cfg)
        print("| %s | %.3g | %.3g | %.3g | %s |" % (name, np.abs(q - base_q).max(), np.abs(q[0] - base_q[0]).max(),
                                                   np.abs(pose[:, :3] - base_pose[:, :3]).max(), t))


if __name__ == "__main__":
    main()
