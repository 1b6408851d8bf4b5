from .env import REALRobotEnv, Goal  # noqa: F401
from ..videomaker import EnvCamera  # noqa: F401,E402  (reference: real_robots.envs.EnvCamera, env.py:470)
