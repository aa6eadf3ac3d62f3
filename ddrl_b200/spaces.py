"""Minimal stand-ins for gym.spaces (gym is not a dependency of the learner path).  When gym is importable the
real classes are re-exported so objects interoperate with RLlib."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gym is absent in the build container
    from gym.spaces import Box, MultiDiscrete, Tuple  # type: ignore
except Exception:
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def __repr__(self):
            return f"Box{self.shape}"

    class MultiDiscrete:
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.shape, self.dtype = self.nvec.shape, np.dtype(np.int64)

        def __repr__(self):
            return f"MultiDiscrete{self.shape}"

    class Tuple:
        def __init__(self, spaces):
            self.spaces = tuple(spaces)
            self.shape = None

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def __getitem__(self, i):
            return self.spaces[i]

        def __repr__(self):
            return f"Tuple{self.spaces}"
