from real_robots_b200.videomaker import *  # noqa: F401,F403
from real_robots_b200.videomaker import VideoMaker, EnvCamera, Frames  # noqa: F401
