"""real_robots_b200 — B200-native batched simulator for the REALRobot2020 environment.

Host-side mirror of the reference's Python interface (AIcrowd/real_robots: real_robots/__init__.py,
real_robots/envs/env.py, real_robots/envs/robot.py, real_robots/evaluate.py, real_robots/policy.py) over the
C ABI in include/real_b200.h.  There is no CPU fallback: the CUDA library must be built and a GPU present to
create an environment (importing the package and inspecting spaces works anywhere).
"""
import os

__version__ = "0.1.0"

from .policy import BasePolicy  # noqa: F401,E402
from .evaluate import evaluate  # noqa: F401,E402

# real_robots/__init__.py:16-28 — 18 ids REALRobot2020-{R1,R2}{J,C,M}{1,2,3}-v0
_ACTION_TYPES = {"J": "joints", "C": "cartesian", "M": "macro_action"}
REGISTRY = {}
for _rnd, _additional in (("R1", True), ("R2", False)):
    for _a, _atype in _ACTION_TYPES.items():
        for _n in (1, 2, 3):
            REGISTRY["REALRobot2020-{}{}{}-v0".format(_rnd, _a, _n)] = dict(
                additional_obs=_additional, objects=_n, action_type=_atype)


def make(env_id, **kwargs):
    """gym.make for the REALRobot2020 ids (works with or without `gym` installed)."""
    if env_id not in REGISTRY:
        raise KeyError("No registered env with id: {}".format(env_id))
    from .envs.env import REALRobotEnv
    kw = dict(REGISTRY[env_id])
    kw.update(kwargs)
    return REALRobotEnv(**kw)


def getPackageDataPath():                 # real_robots/__init__.py:34-39
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def _register_with_gym():                 # pragma: no cover - gym is absent in the build image
    try:
        from gym.envs.registration import register
    except Exception:
        return
    for env_id, kw in REGISTRY.items():
        try:
            register(id=env_id, entry_point='real_robots_b200.envs:REALRobotEnv', kwargs=kw)
        except Exception:
            pass


_register_with_gym()
