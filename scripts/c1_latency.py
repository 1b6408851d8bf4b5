"""N = 1 latency of the gym facade (deterministic random policy), for the library knobs in the environment."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import real_robots
env = real_robots.make('REALRobot2020-R2J3-v0')
env.reset()
rng = np.random.default_rng(1234)
lo, hi = env.action_space['joint_command'].low, env.action_space['joint_command'].high
a = rng.uniform(lo, hi)
out = []
for render in (True, False):
    for t in range(50):
        env.step({'joint_command': a, 'render': render})
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(1000):
        if rng.random() < 0.05:
            a = rng.uniform(lo, hi)
        env.step({'joint_command': a, 'render': render})
    out.append((time.perf_counter() - t0))
print("heavy %s eager %s: render True %.3f ms/step, render False %.3f ms/step" % (os.environ.get("RR_HEAVY_SOLVER", "1"), os.environ.get("RR_EAGER_COLLIDE", "1"), out[0], out[1]))
