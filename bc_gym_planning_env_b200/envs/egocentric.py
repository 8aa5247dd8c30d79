"""EgocentricCostmap observation wrapper (reference envs/egocentric.py:17-160): turns the rich
Observation of a (single-env) PlanEnv-like object into {'env': uint8 (H, W, 1) egocentric crop,
'goal_n_state': float32 (9, 1)}.  The crop and the vector come from the CUDA egocentric kernel
(`bcg_observe_ego`); for batches use VecPlanEnv(with_ego=True) directly."""
from collections import OrderedDict

import numpy as np

from bc_gym_planning_env_b200.envs.base import spaces


class ObservationWrapper(object):
    def __init__(self, env):
        self.env = env
        self.action_space = self.env.action_space

    def unwrapped(self):
        return self.env

    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        return self.observation(observation), reward, done, info

    def reset(self):
        return self.observation(self.env.reset())

    def observation(self, observation):
        raise NotImplementedError

    def render(self, mode='human'):
        return self.env.render(mode)

    def close(self):
        if self.env:
            self.env.close()

    def seed(self, seed=None):
        self.env.seed(seed)

    def get_state(self):
        return self.env.get_state().copy()

    def set_state(self, state):
        self.env.set_state(state)


def _plan_env_of(env):
    """The PlanEnv at the bottom of RandomAisleTurnEnv / RandomMiniEnv / PlanEnv."""
    return getattr(env, '_env', env)


class EgocentricCostmap(ObservationWrapper):
    def __init__(self, env):
        super(EgocentricCostmap, self).__init__(env)
        cp = _plan_env_of(env)._vec._c_params
        self.observation_space = spaces.Dict(OrderedDict((
            ('env', spaces.Box(low=0, high=255, shape=(cp.ego_h, cp.ego_w, 1), dtype=np.uint8)),
            ('goal', spaces.Box(low=-1., high=1., shape=(3, 1), dtype=np.float64)))))

    def observation(self, observation):
        vec = _plan_env_of(self.env)._vec
        image, goal_n_state = vec.observe_ego()
        return OrderedDict((('env', image[0].cpu().numpy()), ('goal_n_state', goal_n_state[0].cpu().numpy())))
