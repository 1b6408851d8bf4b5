from .env import REALRobotEnv, Goal  # noqa: F401
