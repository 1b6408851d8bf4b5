"""Goal generation on the batched simulator (reference real_robots/generate_goals.py:21-365, 395-439).

The reference rejection-samples object placements one env at a time: draw start poses with a minimum separation
(`drawPosition` :133-226, `generatePosition` :79-108), let the objects settle by stepping the env with a zero action
until they are still for 21 consecutive steps (`runEnv` :21-68), snapshot retina / mask, and keep (initial, final)
pairs that satisfy the challenge's constraints (`generateGoalREAL2020` :275-365).  Here every candidate runs in its own
env of a VecREALRobotEnv, so a whole batch of candidates settles in parallel; accepted pairs are written in the
reference's file format: `np.savez_compressed(path, list_of_Goal)` (:435), loadable by `REALRobotEnv.load_goals`.
"""
import math

import numpy as np

from .envs.env import Goal
from .util import euler_to_quat

OBJECTS = ["cube", "tomato", "mustard"]
BASE_ORIENTATION = {"cube": [0.0, 0.0, 0.0], "tomato": [0.0, 0.0, 0.0], "mustard": [1.5708, 3.14159, 0.0]}   # robot.py:19-24
# isOnShelf / isOnTable thresholds (generate_goals.py:245-272)
SHELF_Z = {"cube": 0.40, "tomato": 0.40, "mustard": 0.395}
TABLE_Z = {"cube": 0.33, "tomato": 0.34, "mustard": 0.33}


def generate_position(rng, obj, fixed, table_plane, base_quat=None):
    """generatePosition (:79-108): x range by table plane, y in [-.45,.45], z = 0.40 on the table / 0.50 above the shelf.
    A fixed orientation is the one the object SETTLES in after a reset (basePosition[obj][3:], :100, :418), passed in
    base_quat; without it (host-only tests) the nominal Euler angles of Kuka.object_poses are used."""
    if table_plane is None:
        min_x, max_x = -.25, .25
    elif table_plane:
        min_x, max_x = -.25, .05
    else:
        min_x, max_x = .10, .25
    x = rng.random() * (max_x - min_x) + min_x
    y = rng.random() * 0.9 - 0.45
    z = 0.40 if x <= 0.05 else 0.50
    if fixed:
        quat = list(base_quat[obj]) if base_quat is not None else euler_to_quat(BASE_ORIENTATION[obj])
    else:
        quat = euler_to_quat((rng.random(3) * math.pi * 2).tolist())
    return np.array([x, y, z] + list(quat))


def draw_start_poses(rng, objects, fixed_orientation, table_plane, min_separation, base_quat=None):
    """The sampling half of drawPosition (:133-165): permute objects, redraw an object until it clears the others."""
    poses = {}
    for obj in rng.permutation(objects):
        while True:
            p = generate_position(rng, obj, fixed_orientation, table_plane, base_quat)
            others = [q[:3] for q in poses.values()]
            if not others or min(np.linalg.norm(p[:3] - o) for o in others) >= min_separation:
                poses[obj] = p
                break
    return poses


def settle(env, max_t=1000):
    """runEnv (:21-68), batched: zero action until every object moved < 1e-4 m and < 1e-3 in quaternion for 21
    consecutive steps (and t > 10).  runEnv returns at the step an env meets the test, so every env's poses are LATCHED at
    that step (the batch keeps stepping for the slower envs) and written back before returning, velocities zeroed, so that a
    render after settle() shows the latched state.  Returns (still[N] bool, poses[N, n_obj, 7])."""
    import torch
    N = env.n_envs
    zero = torch.zeros(N, 9, dtype=torch.float64, device=env.device)
    pos = env.sim.read_state()[2].clone()
    stable = torch.zeros(N, dtype=torch.int64, device=env.device)
    still = torch.zeros(N, dtype=torch.bool, device=env.device)
    latched = pos.clone()
    for t in range(max_t):
        env.sim.step(zero)
        new = env.sim.read_state()[2].clone()
        dpos = (new[:, :, :3] - pos[:, :, :3]).norm(dim=2).max(dim=1).values
        dq = torch.minimum((new[:, :, 3:] - pos[:, :, 3:]).norm(dim=2), (new[:, :, 3:] + pos[:, :, 3:]).norm(dim=2)).max(dim=1).values
        ok = (dpos < 0.0001) & (dq < 0.001) & (t > 10)
        stable = torch.where(ok, stable + 1, torch.zeros_like(stable))
        now = (stable > 20) & ~still
        latched = torch.where(now[:, None, None], new, latched)
        still |= now
        pos = new
        if bool(still.all()):
            break
    latched = torch.where(still[:, None, None], latched, pos)
    env.set_object_poses(latched.contiguous(), mask=still)
    return still, latched


def _min_separation(poses):
    p = poses[:, :, :3]
    if p.shape[1] < 2:
        return np.full(p.shape[0], np.inf)
    d = np.linalg.norm(p[:, :, None] - p[:, None], axis=3)
    d[d == 0] = np.inf
    return d.min(axis=(1, 2))


def generate_goals(n_goals, n_objects=1, challenge='2D', seed=2020, device=0, batch=None, min_start_goal_dist=0.10,
                   min_objects_dist=0.05, max_rounds=20):
    """Returns a list of `Goal` (initial_state / final_state: dict name -> pose7, retina, retina_before, mask, challenge).

    challenge '2D': fixed (base) orientations, every object drawn on the table plane (objOnTable, :280-285); '2.5D': fixed
    orientations, x drawn over table and shelf, at least one object on the shelf in the initial or the final state
    (:314-319, 345-351); '3D': the same with free orientations.  Start and goal of every object are at least
    `min_start_goal_dist` apart in the table plane (:353-357, xy only).  (generateGoalREAL2020 :275-365; its
    max_objects_dist = 2 m test is always true on this table and is not reproduced.)"""
    import torch
    from .envs.vec_env import VecREALRobotEnv
    rng = np.random.default_rng(seed) if not hasattr(seed, "random") else seed
    objects = OBJECTS[:n_objects]
    fixed = challenge != '3D'
    plane = True if challenge == '2D' else None
    N = batch or max(8, 2 * n_goals)
    env = VecREALRobotEnv(N, device=device, objects=n_objects, action_type='joints', additional_obs=True)
    # basePosition (:415-418): the poses the objects settle in after a reset; fixed-orientation goals use THOSE orientations
    env.reset(render=False)
    _, base = settle(env, 200)
    base = base[0].cpu().numpy()
    base_quat = {name: base[o, 3:7].copy() for o, name in enumerate(objects)}
    goals = []
    for _ in range(max_rounds):
        if len(goals) >= n_goals:
            break
        states, images = [], []
        for phase in range(2):                       # 0 = initial placement, 1 = goal placement
            start = np.zeros((N, n_objects, 7))
            for e in range(N):
                poses = draw_start_poses(rng, objects, fixed, plane, min_objects_dist, base_quat)
                for o, name in enumerate(objects):
                    start[e, o] = poses[name]
            env.reset(render=False)
            settle(env, 200)                          # generateRealPosition (:111-121): reset, settle, then place
            env.set_object_poses(torch.from_numpy(start).to(env.device))
            still, poses = settle(env)
            rgb, depth, seg = env.sim.render()
            states.append((still.cpu().numpy(), poses.cpu().numpy(), start))
            images.append((rgb.cpu().numpy(), seg.cpu().numpy()))
        (s0, p0, st0), (s1, p1, st1) = states
        ok = s0 & s1 & (_min_separation(p0) >= min_objects_dist) & (_min_separation(p1) >= min_objects_dist)
        ok &= (np.linalg.norm(p0[:, :, :2] - p1[:, :, :2], axis=2) >= min_start_goal_dist).all(axis=1)
        for e in range(N):
            if not ok[e] or len(goals) >= n_goals:
                continue
            good = True
            if fixed:                                 # objects must keep their orientation (:186-200)
                for o in range(n_objects):
                    for p, st in ((p0, st0), (p1, st1)):
                        dq = min(np.linalg.norm(p[e, o, 3:] - st[e, o, 3:]), np.linalg.norm(p[e, o, 3:] + st[e, o, 3:]))
                        good &= dq <= 0.041
            if challenge != '2D':                     # at least one object on the shelf at one end of the move
                good &= any(p[e, o, 2] > SHELF_Z[name] for p in (p0, p1) for o, name in enumerate(objects))
            if not good:
                continue
            g = Goal(initial_state={n: p0[e, o].copy() for o, n in enumerate(objects)},
                     final_state={n: p1[e, o].copy() for o, n in enumerate(objects)},
                     retina_before=images[0][0][e].copy(), retina=images[1][0][e].copy(),
                     mask=images[1][1][e].copy(), challenge=challenge)
            g.subtype = str(n_objects)                # :359
            goals.append(g)
    env.close()
    return goals


def save_goals(goals, path):
    """The reference's file format (generate_goals.py:435): np.savez_compressed(name, list_of_goals)."""
    arr = np.empty(len(goals), dtype=object)
    for i, g in enumerate(goals):
        arr[i] = g
    np.savez_compressed(path, arr)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Generate a REALRobot2020 goals dataset on the GPU")
    ap.add_argument("--seed", type=int, default=None, help="seed of the goal draws (default: none, as in the reference)")
    ap.add_argument("--n_2d_goals", type=int, default=25)
    ap.add_argument("--n_25d_goals", type=int, default=15)
    ap.add_argument("--n_3d_goals", type=int, default=10)
    ap.add_argument("--n_obj", type=int, default=3)
    ap.add_argument("--file", default=None)
    a = ap.parse_args(argv)
    goals = []
    # generate_goals.py:425-433: 2D and 2.5D goals keep objects 0.25 m apart, 3D goals may touch; start and goal 0.2 m apart
    for k, (ch, n, sep) in enumerate((('2D', a.n_2d_goals, 0.25), ('2.5D', a.n_25d_goals, 0.25), ('3D', a.n_3d_goals, 0.0))):
        if n > 0:
            goals += generate_goals(n, a.n_obj, ch, seed=None if a.seed is None else a.seed + k, min_start_goal_dist=0.2, min_objects_dist=sep)
    name = a.file or "goals-REAL2020-s{}-{}-{}-{}-{}.npy".format(a.seed, a.n_2d_goals, a.n_25d_goals, a.n_3d_goals, a.n_obj)
    save_goals(goals, name)
    print("wrote %d goals to %s.npz" % (len(goals), name))


if __name__ == "__main__":
    main()
