from real_robots_b200.envs.env import REALRobotEnv, Goal, EyeCamera, DefaultRewardFunc  # noqa: F401
