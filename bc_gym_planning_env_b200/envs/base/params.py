"""EnvParams / RewardParams: same fields, defaults and wire format as the reference
(envs/base/params.py:15-43, envs/base/reward.py:162-171)."""
import attr
import numpy as np

from bc_gym_planning_env_b200.robot_models.robot_dimensions import INDUSTRIAL_TRICYCLE_V1

CONTINUOUS_REWARD = 'continuous_reward'   # reward_provider_examples.py
CONTINUOUS_REWARD_PURE_PURSUIT = 'continuous_reward_pure_pursuit'


@attr.s
class RewardParams(object):
    VERSION = 1
    spatial_precision = attr.ib(type=float)
    angular_precision = attr.ib(type=float)
    spatial_progress_multiplier = attr.ib(type=float, default=0.0)

    def serialize(self):
        state = attr.asdict(self)
        state['version'] = self.VERSION
        return state

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        return cls(**state)


@attr.s(frozen=True)
class EnvParams(object):
    VERSION = 1
    dt = attr.ib(type=float, default=0.05)
    goal_ang_dist = attr.ib(type=float, default=np.pi / 2)
    goal_spat_dist = attr.ib(type=float, default=1.0)
    initial_wheel_angle = attr.ib(type=float, default=0.0)
    iteration_timeout = attr.ib(type=int, default=1200)
    path_limiter_max_dist = attr.ib(type=float, default=5.0)
    robot_name = attr.ib(default=INDUSTRIAL_TRICYCLE_V1)
    resolution = attr.ib(type=float, default=0.03)
    refine_path = attr.ib(type=bool, default=True)
    path_delta = attr.ib(type=float, default=0.05)
    pose_delay = attr.ib(type=int, default=0)
    control_delay = attr.ib(type=int, default=0)
    state_delay = attr.ib(type=int, default=0)
    reward_provider_name = attr.ib(default=CONTINUOUS_REWARD)
    reward_provider_params = attr.ib(default=attr.Factory(
        lambda self: RewardParams(spatial_precision=self.goal_spat_dist, angular_precision=self.goal_ang_dist),
        takes_self=True))

    def serialize(self):
        state = attr.asdict(self)
        state['version'] = self.VERSION
        state['reward_provider_params'] = self.reward_provider_params.serialize()
        return state

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        state['reward_provider_params'] = RewardParams.deserialize(state['reward_provider_params'])
        return cls(**state)
