from real_robots_b200.envs.env import REALRobotEnv, Goal  # noqa: F401
from real_robots_b200.videomaker import EnvCamera  # noqa: F401
from . import env, robot  # noqa: F401
