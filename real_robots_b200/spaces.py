"""gym.spaces stand-ins (Box, Dict, MultiBinary) with the subset of the API the reference uses
(real_robots/envs/robot.py:69-112, real_robots/envs/env.py:44-76): .sample(), .contains(), .low/.high/.shape/.dtype,
Dict.spaces.  When `gym` is importable its own classes are used instead, so isinstance checks of user code hold."""
import numpy as np

try:                                    # pragma: no cover - gym is not installed in the build image
    from gym import spaces as _gym_spaces
    Box, Dict, MultiBinary = _gym_spaces.Box, _gym_spaces.Dict, _gym_spaces.MultiBinary
    HAVE_GYM = True
except Exception:
    HAVE_GYM = False

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def __contains__(self, x):
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                low = np.asarray(low)
                high = np.asarray(high)
                shape = low.shape if low.shape else high.shape
            shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape).astype(dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), shape).astype(dtype)
            super().__init__(shape, dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e3).astype(np.float64)
            hi = np.where(np.isfinite(self.high), self.high, 1e3).astype(np.float64)
            if self.dtype.kind in "iu":
                return self._rng.randint(lo.astype(np.int64), hi.astype(np.int64) + 1, size=self.shape).astype(self.dtype)
            return self._rng.uniform(lo, hi, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return "Box%s" % (self.shape,)

    class MultiBinary(Space):
        def __init__(self, n):
            self.n = n
            super().__init__((n,), np.int8)

        def sample(self):
            return self._rng.randint(0, 2, size=self.n).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x == 0) | (x == 1)).all())

    class Dict(Space):
        def __init__(self, spaces):
            self.spaces = dict(spaces)
            super().__init__(None, None)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def __getitem__(self, k):
            return self.spaces[k]
