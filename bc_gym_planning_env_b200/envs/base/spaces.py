"""The slice of the reference's gym-like spaces the step path uses (envs/base/spaces.py):
`Box` with `sample()` returning an Action drawn from a module-level RandomState(0) (:9-10,
:134-141), plus `Dict`."""
from collections import OrderedDict

import numpy as np

from bc_gym_planning_env_b200.envs.base.action import Action

SPACE_LOCAL_RANDOM_STATE = np.random.RandomState()
SPACE_LOCAL_RANDOM_STATE.seed(0)


class Box(object):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        if shape is None:
            assert low.shape == high.shape
            shape = low.shape
        else:
            assert np.isscalar(low) and np.isscalar(high)
            low = low + np.zeros(shape)
            high = high + np.zeros(shape)
        if dtype is None:
            dtype = 'uint8' if (high == 255).all() else 'float32'
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        self.shape = tuple(shape)
        self.dtype = dtype

    def sample(self):
        integer = np.dtype(self.dtype).kind != 'f'
        v, w = SPACE_LOCAL_RANDOM_STATE.uniform(low=self.low, high=self.high + (1 if integer else 0),
                                                size=self.low.shape).astype(self.dtype)
        return Action(command=np.array([v, w]))

    def contains(self, x):
        return x.shape == self.shape and (x >= self.low).all() and (x <= self.high).all()

    __contains__ = contains


class Dict(object):
    def __init__(self, spaces):
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(spaces.items()))
        self.spaces = OrderedDict(spaces)
