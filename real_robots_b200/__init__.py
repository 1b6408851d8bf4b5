"""real_robots_b200 — B200-native batched simulator for the REALRobot2020 environment.

Host-side mirror of the reference's Python interface (AIcrowd/real_robots: real_robots/__init__.py,
real_robots/envs/env.py, real_robots/envs/robot.py, real_robots/evaluate.py, real_robots/policy.py) over the
C ABI in include/real_b200.h.  There is no CPU fallback: the CUDA library must be built and a GPU present.
"""
__version__ = "0.1.0"
