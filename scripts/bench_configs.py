"""Throughput of the other BASELINE.json configs on one GPU (documentation numbers; bench.py measures the headline C3).
C2: joints action, 4096 envs, render off.  C3: render on.  C5: cartesian action, full R1 observation, N/GPU sweep."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from real_robots_b200.sim import BatchedSim
from real_robots_b200.envs.vec_env import VecREALRobotEnv


def timed(fn, steps):
    for t in range(5):                  # the call pattern's own warm-up (first render: output buffers; the solver schedule adapts to the pattern)
        fn(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        fn(t)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


dev = torch.device("cuda")

# C1: the reference's own configuration through the drop-in gym surface — one env, joints action, random policy
# (cli.RandomPolicy: p = 0.05 resample), 1000 steps with the 240x320 retina, numpy observations on the host
import numpy as np
import real_robots
env = real_robots.make('REALRobot2020-R2J3-v0')
obs = env.reset()
rng = np.random.default_rng(1234)
lo_a, hi_a = env.action_space['joint_command'].low, env.action_space['joint_command'].high
a = rng.uniform(lo_a, hi_a)          # (not action_space.sample(): gym's own generator is not seeded, and one env's step time depends on what its arm touches)
for render in (True, False):
    for t in range(50):
        env.step({'joint_command': a, 'render': render})
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(1000):
        if rng.random() < 0.05:
            a = rng.uniform(lo_a, hi_a)
        obs, _, done, _ = env.step({'joint_command': a, 'render': render})
    dt = time.perf_counter() - t0
    print("C1  REALRobot2020-R2J3-v0, 1 env, render %-5s: %7.3f ms/step  %10.0f env-steps/s  (gym facade, numpy obs, wall clock)" % (render, dt, 1000 / dt))
assert obs['retina'].shape == (240, 320, 3)
env.close()

N = 4096
sim = BatchedSim(N, 3)
acts = bench.random_policy_actions(torch, 200, N, dev, 1234)
for t in range(100):
    sim.step(acts[t])
ms = timed(lambda t: (sim.step(acts[100 + t % 100]), sim.read_state()), 100)
print("C2  joints, %5d envs, render off : %7.3f ms/step  %10.0f env-steps/s" % (N, ms, N / ms * 1e3))
ms = timed(lambda t: (sim.step(acts[100 + t % 100]), sim.render(), sim.read_state()), 50)
print("C3  joints, %5d envs, render on  : %7.3f ms/step  %10.0f env-steps/s" % (N, ms, N / ms * 1e3))
ms = timed(lambda t: (sim.step(acts[100 + t % 100]), sim.read_state(), sim.render()), 50)
print("C3' the same, state read BEFORE the render (the solver's heavy-env kernel cannot hide behind the rasteriser): %7.3f ms/step  %10.0f env-steps/s" % (ms, N / ms * 1e3))
sim.close()
del sim
for n in (1024, 2048, 4096, 8192):
    env = VecREALRobotEnv(n, objects=3, action_type='cartesian', additional_obs=True)
    env.reset(render=False)
    g = torch.Generator(device=dev); g.manual_seed(5)
    lo = torch.tensor([-0.25, -0.5, 0.40, -1, -1, -1, -1], dtype=torch.float64, device=dev)
    hi = torch.tensor([0.25, 0.5, 0.60, 1, 1, 1, 1], dtype=torch.float64, device=dev)
    cart = lo + (hi - lo) * torch.rand(n, 7, dtype=torch.float64, device=dev, generator=g)
    grip = torch.rand(n, 2, dtype=torch.float64, device=dev, generator=g) * 1.5707963

    def step(t, cart=cart, grip=grip, env=env, n=n):
        if t % 20 == 0:                          # p = 0.05 resample, here every 20th step for all envs: a new IK solve
            cart.copy_(lo + (hi - lo) * torch.rand(n, 7, dtype=torch.float64, device=dev, generator=g))
        env.step({'cartesian_command': cart, 'gripper_command': grip, 'render': True})
    for t in range(60):
        step(t)
    ms = timed(step, 40)
    print("C5  cartesian, %5d envs, R1 obs  : %7.3f ms/step  %10.0f env-steps/s" % (n, ms, n / ms * 1e3))
    env.close()
    del env
    torch.cuda.empty_cache()
