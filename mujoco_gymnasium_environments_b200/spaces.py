"""``gymnasium.spaces.Box`` when Gymnasium is importable, else a minimal stand-in with the same attributes.

Gymnasium is absent from the authoring container and the GPU box; the env classes only need ``low``, ``high``,
``shape``, ``dtype``, ``sample()`` and ``contains()``.
"""
import numpy as np

try:  # pragma: no cover - exercised only where gymnasium is installed
    from gymnasium.spaces import Box  # type: ignore
    from gymnasium import Env as _GymEnv  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # ModuleNotFoundError here
    HAVE_GYMNASIUM = False

    class Box:  # noqa: D401
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            low = np.asarray(low, dtype=dtype); high = np.asarray(high, dtype=dtype)
            if shape is not None:
                low = np.broadcast_to(low, shape).copy(); high = np.broadcast_to(high, shape).copy()
            self.low, self.high, self.shape, self.dtype = low, high, low.shape, np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6); hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class _GymEnv:  # minimal base so the class API works without gymnasium
        metadata = {}


def batch_box(space, n):
    return Box(np.repeat(space.low[None], n, 0), np.repeat(space.high[None], n, 0), dtype=space.dtype)
