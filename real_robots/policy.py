from real_robots_b200.policy import *  # noqa: F401,F403
from real_robots_b200.policy import BasePolicy  # noqa: F401
