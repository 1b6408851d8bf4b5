"""Drop-in alias: `import real_robots` resolves to the B200-native implementation (real_robots_b200), so user code
and pickled goal files (class path real_robots.envs.env.Goal, reference env.py:15-24 / generate_goals.py:435) keep working."""
from real_robots_b200 import *  # noqa: F401,F403
from real_robots_b200 import evaluate, make, getPackageDataPath, BasePolicy, REGISTRY, __version__  # noqa: F401
from . import envs, policy, videomaker, generate_goals  # noqa: F401


def copy_over_data_into_pybullet(force_copy=False):
    """Kept for source compatibility (reference __init__.py:42-64).  There is no pybullet_data directory to fill: the
    URDF / OBJ / PNG assets are compiled into real_robots_b200/assets/scene.rrb (tools/compile_assets.py)."""
    return None
