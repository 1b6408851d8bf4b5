from real_robots_b200.generate_goals import *  # noqa: F401,F403
from real_robots_b200.generate_goals import generate_goals, save_goals, main  # noqa: F401
