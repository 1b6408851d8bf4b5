from real_robots_b200.envs.robot import *  # noqa: F401,F403
from real_robots_b200.envs.robot import Kuka  # noqa: F401
