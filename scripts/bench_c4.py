"""BASELINE.json config 4 on one GPU: R-macro_action (cartesian IK + trajectory) with 3 objects, the extrinsic phase of
evaluate() on synthetic goals (initial / final object xy ~ U([-0.25, 0.05] x [-0.45, 0.45]), z = 0.40 then
settle, seed 2020, one 1000-step trial per goal = one full macro plan), trial-parallel through evaluate_vectorised.
Usage: python scripts/bench_c4.py [G ...]   (default 50 4096)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from real_robots_b200.evaluate import evaluate_vectorised
from real_robots_b200.envs.env import MACRO_BOUNDS
from real_robots_b200.envs.vec_env import VecREALRobotEnv

N_OBJ, TRIAL = 3, 1000


def synthetic_goals(G, seed=2020, settle_steps=300):
    """Drop the objects from z = 0.40 at random table positions and let them settle (generate_goals.py:79-108 with
    tablePlane=True, without the rejection rules): the settled poses are the goals' initial / final states."""
    rng = np.random.default_rng(seed)
    out = []
    env = VecREALRobotEnv(G, objects=N_OBJ, action_type='joints', additional_obs=True)
    for _ in range(2):
        pq = np.zeros((G, N_OBJ, 7)); pq[..., 6] = 1.0
        pq[..., 0] = rng.uniform(-0.25, 0.05, (G, N_OBJ)); pq[..., 1] = rng.uniform(-0.45, 0.45, (G, N_OBJ)); pq[..., 2] = 0.40
        env.reset(render=False)
        env.set_object_poses(torch.from_numpy(pq).to(env.device))
        for _ in range(settle_steps):
            env.step({'joint_command': None, 'render': False})
        out.append(env.sim.read_state()[2].cpu().numpy())
    env.close()
    return dict(initial=out[0], final=out[1][:, :, :3].copy())


class RandomMacroPolicy:
    """One U(macro_space) draw per trial, repeated every step (a plan is 1000 rows long: env.py:447-454)."""
    def __init__(self, env, seed=2020):
        g = torch.Generator(device=env.device); g.manual_seed(seed)
        lo = torch.tensor(MACRO_BOUNDS[0], dtype=torch.float64, device=env.device)
        hi = torch.tensor(MACRO_BOUNDS[1], dtype=torch.float64, device=env.device)
        self.a = lo + (hi - lo) * torch.rand(env.n_envs, 2, 2, dtype=torch.float64, device=env.device, generator=g)

    def step(self, obs, reward, done):
        return {'macro_action': self.a}


for G in [int(a) for a in sys.argv[1:]] or [50, 4096]:
    goals = synthetic_goals(G)
    for render in (False, True):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        score, scores, vec = evaluate_vectorised(lambda env: RandomMacroPolicy(env), goals, n_objects=N_OBJ,
                                                 action_type='macro_action', extrinsic_timesteps=TRIAL, render=render)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("C4  macro_action, %5d goals x %d steps, render %-5s: %7.2f s wall  %10.0f env-steps/s  %7.1f trials/s  mean score %.4f"
              % (G, TRIAL, render, dt, G * TRIAL / dt, G / dt, float(np.mean(vec))), flush=True)
