"""PlanEnv: the reference's single-environment API (envs/base/env.py:217-439) over the batched
CUDA path.  One PlanEnv is a VecPlanEnv of one env; every step is the same `bcg_step` call the batch
makes, followed by a device->host read to build reference-shaped `Observation` / `State` objects.
Use VecPlanEnv directly for throughput; this facade exists so reference code runs unchanged.
"""
import copy

import attr
import numpy as np
import torch

from bc_gym_planning_env_b200 import _native as nat
from bc_gym_planning_env_b200.envs.base import spaces
from bc_gym_planning_env_b200.envs.base.action import Action
from bc_gym_planning_env_b200.envs.base.obs import Observation, robot_state_type
from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
from bc_gym_planning_env_b200.vec_env import DEFAULT_NOISE, VecPlanEnv, VecState


@attr.s(eq=False)
class ContinuousRewardProviderState(object):
    """reference envs/base/reward.py:12-75"""
    min_spat_dist_so_far = attr.ib(type=float)
    path = attr.ib(type=np.ndarray)
    target_idx = attr.ib(type=int)
    VERSION = 1

    def __eq__(self, other):
        return (isinstance(other, ContinuousRewardProviderState) and not (self.path != other.path).any()
                and self.min_spat_dist_so_far == other.min_spat_dist_so_far and self.target_idx == other.target_idx)

    def __ne__(self, other):
        return not self.__eq__(other)

    def copy(self):
        return attr.evolve(self, path=np.copy(self.path))

    def current_goal_pose(self):
        if self.target_idx < len(self.path):
            return self.path[self.target_idx]
        raise ValueError("No path left to follow.")

    def current_path(self):
        return self.path[self.target_idx:]

    def done(self):
        return self.target_idx > len(self.path) - 1

    def get_reward_provider_state_type_name(self):
        return self.reward_provider_state_type_name

    def serialize(self):
        return dict(min_spat_dist_so_far=self.min_spat_dist_so_far, path=self.path, target_idx=self.target_idx,
                    version=self.VERSION)

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        return cls(**state)


CONTINUOUS_REWARD_STATE = 'continuous_reward_state'      # reference reward_provider_examples.py
CONTINUOUS_REWARD_PURE_PURSUIT_STATE = 'continuous_reward_pure_pursuit_state'
ContinuousRewardProviderState.reward_provider_state_type_name = CONTINUOUS_REWARD_STATE


@attr.s(eq=False)
class ContinuousRewardPurePursuitProviderState(ContinuousRewardProviderState):
    """reference envs/base/reward.py:78-159: the goal is always the last path point, `target_idx` is the
    look-ahead way point and the current path is everything up to it."""
    reward_provider_state_type_name = CONTINUOUS_REWARD_PURE_PURSUIT_STATE

    def current_goal_pose(self):
        return self.path[-1]

    def current_path(self):
        return self.path[:self.target_idx + 1]

    def done(self, state=None):
        pose = state.pose if state is not None else None
        if pose is None:
            raise ValueError("the pure-pursuit provider needs the env state to decide done")
        goal = self.current_goal_pose()
        return bool(np.hypot(goal[0] - pose[0], goal[1] - pose[1]) < 1.0)


_REWARD_STATE_TYPES = {CONTINUOUS_REWARD_STATE: ContinuousRewardProviderState,
                       CONTINUOUS_REWARD_PURE_PURSUIT_STATE: ContinuousRewardPurePursuitProviderState}


@attr.s(eq=False)
class State(object):
    """Snapshot a PlanEnv can be reset to (reference envs/base/env.py:52-68): same fields."""
    reward_provider_state = attr.ib(type=object)
    path = attr.ib(type=np.ndarray)
    original_path = attr.ib(type=np.ndarray)
    costmap = attr.ib(type=CostMap2D)
    iter_timeout = attr.ib(type=int)
    current_time = attr.ib(type=float)
    current_iter = attr.ib(type=int)
    robot_collided = attr.ib(type=bool)
    poses_queue = attr.ib(type=list)
    robot_state_queue = attr.ib(type=list)
    control_queue = attr.ib(type=list)
    pose = attr.ib(type=np.ndarray)
    robot_state = attr.ib(type=object)
    VERSION = 1

    def copy(self):
        return attr.evolve(
            self, reward_provider_state=self.reward_provider_state.copy(), path=np.copy(self.path),
            pose=np.copy(self.pose), original_path=np.copy(self.original_path), costmap=self.costmap.copy(),
            poses_queue=copy.deepcopy(self.poses_queue), robot_state_queue=copy.deepcopy(self.robot_state_queue),
            control_queue=copy.deepcopy(self.control_queue), robot_state=self.robot_state.copy())

    def __eq__(self, other):
        if not isinstance(other, State):
            return False
        return (self.reward_provider_state == other.reward_provider_state
                and self.path.shape == other.path.shape and not (self.path != other.path).any()
                and not (self.original_path != other.original_path).any()
                and self.costmap == other.costmap and self.iter_timeout == other.iter_timeout
                and self.current_time == other.current_time and self.current_iter == other.current_iter
                and self.robot_collided == other.robot_collided
                and len(self.poses_queue) == len(other.poses_queue)
                and all((a == b).all() for a, b in zip(self.poses_queue, other.poses_queue))
                and self.robot_state_queue == other.robot_state_queue and self.control_queue == other.control_queue
                and not (self.pose != other.pose).any() and self.robot_state == other.robot_state)

    def __ne__(self, other):
        return not self.__eq__(other)

    # Wire format: a dict of basic types that pickles, with the reference's keys (envs/base/env.py:139-176).
    # (The reference writes 'reward_provider_state_type_name' but reads 'reward_provider_state_name' and
    # leaves queue items without a version; both spellings are accepted here and every item is versioned.)
    def serialize(self):
        return dict(
            version=self.VERSION,
            reward_provider_state_type_name=self.reward_provider_state.get_reward_provider_state_type_name(),
            reward_provider_state=self.reward_provider_state.serialize(),
            path=self.path, original_path=self.original_path, costmap=self.costmap.get_state(),
            iter_timeout=self.iter_timeout, current_time=self.current_time, current_iter=self.current_iter,
            robot_collided=self.robot_collided,
            poses_queue=[np.array(p) for p in self.poses_queue],
            robot_state_queue=[s.serialize() for s in self.robot_state_queue],
            control_queue=[a.serialize() for a in self.control_queue],
            pose=self.pose, robot_type_name=self.robot_state.get_robot_type_name(),
            robot_state=self.robot_state.serialize())

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        name = state.pop('reward_provider_state_type_name', None) or state.pop('reward_provider_state_name', None)
        if name not in _REWARD_STATE_TYPES:
            raise Exception('No reward provider state name "{}" exists'.format(name))
        state.pop('reward_provider_state_name', None)
        state['reward_provider_state'] = _REWARD_STATE_TYPES[name].deserialize(state['reward_provider_state'])
        state['costmap'] = CostMap2D.from_state(state['costmap'])
        rs_type = robot_state_type(state.pop('robot_type_name'))
        state['robot_state'] = rs_type.deserialize(state['robot_state'])
        state['robot_state_queue'] = [rs_type.deserialize(s) for s in state['robot_state_queue']]
        state['control_queue'] = [Action.deserialize(a) for a in state['control_queue']]
        return cls(**state)


class _RobotView(object):
    """The slice of TricycleRobot's interface that scripts written for the reference poke at
    (`env._robot.set_noise_parameters(None)`, `get_pose`, `get_footprint`)."""

    def __init__(self, env):
        self._env = env

    def set_noise_parameters(self, noise_parameters):
        self._env._set_noise(noise_parameters)

    def get_pose(self):
        return self._env._vec.state_f[nat.F_ROBOT:nat.F_ROBOT + 3, 0].cpu().numpy()

    def get_footprint(self):
        return self._env._vec.dims.footprint()

    def get_max_front_wheel_speed(self):
        return self._env._vec.dims.max_front_wheel_speed

    def get_dimensions(self):
        return self._env._vec.dims


class PlanEnv(object):
    """Poses planning problem as OpenAI gym task (reference envs/base/env.py:217)."""

    def __init__(self, costmap, path, params, noise_parameters=DEFAULT_NOISE, seed=0, device=None):
        """
        :param costmap CostMap2D: costmap denoting obstacles
        :param path array(N, 3): oriented path, presented as way points
        :param params EnvParams: parametrization of the environment
        :param noise_parameters: odometry noise alphas (default: PlanEnv's, reference :226-232) or None
        :param seed int: Philox key of the noise stream (the reference uses the global np.random)
        """
        self._params = params
        self._noise_kwargs = dict(noise_parameters=noise_parameters, seed=seed)
        self._vec = VecPlanEnv([costmap], [path], params, n_envs=1, noise_parameters=noise_parameters, seed=seed,
                               device=device)
        self._robot = _RobotView(self)
        low, high = self._vec.action_bounds()
        self.action_space = spaces.Box(low=low, high=high, dtype=np.float32)
        self.reward_range = (0.0, 1.0)
        self._costmap = costmap
        self._state_type = robot_state_type(params.robot_name)
        self._reward_state_type = (ContinuousRewardPurePursuitProviderState
                                   if self._vec._c_params.reward_kind == nat.REWARD_PURE_PURSUIT
                                   else ContinuousRewardProviderState)

    # ---- reference API -----------------------------------------------------------------------------
    def reset(self):
        self._vec.reset()
        return self._extract_obs()

    def step(self, action):
        cmd = np.asarray(action.command if isinstance(action, Action) else action)
        if cmd.dtype not in (np.float32, np.float64):
            cmd = cmd.astype(np.float64)
        _, reward, done, _ = self._vec.step(cmd.reshape(1, 2))
        return self._extract_obs(), float(reward[0]), bool(done[0]), {}

    def get_state(self):
        return self._to_state(self._vec.get_state())

    def set_state(self, state):
        self._vec.set_state(self._from_state(state), load_delayed_robot=True)

    def seed(self, seed=None):
        """No-op like the reference (env.py:325-332): the base env is deterministic given its noise key."""
        pass

    VERSION = 1

    def serialize(self):
        """Dict of basic types, reference keys (envs/base/env.py:251-261): pickle it to checkpoint the env."""
        state = self.get_state()
        return dict(version=self.VERSION, state=state.serialize(), params=self._params.serialize(),
                    path=state.original_path, costmap=state.costmap.get_state(),
                    noise=self._noise_kwargs)             # extension: the odometry-noise setting and Philox key

    @classmethod
    def deserialize(cls, state):
        """Rebuild an env and put it in the serialized state (envs/base/env.py:263-276).  The path in the
        dict is the already refined one, so it is not refined again."""
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        params = EnvParams.deserialize(state['params'])
        noise = state.get('noise') or dict(noise_parameters=DEFAULT_NOISE, seed=0)
        env = cls(CostMap2D.from_state(state['costmap']), state['path'], attr.evolve(params, refine_path=False), **noise)
        env._params = params
        env.set_state(State.deserialize(state['state']))
        return env

    def render(self, mode='human'):
        raise NotImplementedError("rendering is outside the B200 step path (reference envs/base/draw.py)")

    def close(self):
        pass

    # ---- conversions -----------------------------------------------------------------------------
    def _set_noise(self, noise_parameters):
        self._noise_kwargs = dict(self._noise_kwargs, noise_parameters=noise_parameters)
        v = self._vec
        v._c_params.noise_on = 0 if noise_parameters is None else 1
        if noise_parameters is not None:
            for k in range(6):
                v._c_params.alpha[k] = float(noise_parameters['alpha%d' % (k + 1)])

    def _extract_obs(self):
        v = self._vec
        f = v.state_f[:, 0].cpu().numpy()
        target = int(v.state_i[nat.I_TARGET, 0])
        return Observation(pose=f[nat.F_DPOSE:nat.F_DPOSE + 3].copy(), path=self._reward_state_type(0.0, v.full_path(0), target).current_path(),
                           costmap=self._costmap, robot_state=self._state_type.from_row(f[nat.F_DROBOT:nat.F_DROBOT + 7]),
                           time=float(f[nat.F_TIME]), dt=self._params.dt)

    def _queue(self, f, i, which):
        row0, slots, ncomp, irow = self._vec.queue_rows(which)
        if slots == 0:
            return []
        head, length = int(i[irow]) & 0xffff, int(i[irow]) >> 16
        return [f[row0 + ((head + k) % slots) * ncomp: row0 + ((head + k) % slots) * ncomp + ncomp].copy()
                for k in range(length)]

    def _to_state(self, vs):
        f, i = vs.f[:, 0].cpu().numpy(), vs.i[:, 0].cpu().numpy()
        full = self._vec.full_path(0)
        target = int(i[nat.I_TARGET])
        rps = self._reward_state_type(min_spat_dist_so_far=float(f[nat.F_MIN_DIST]), path=full.copy(), target_idx=target)
        return State(
            reward_provider_state=rps, path=rps.current_path().copy(), original_path=full.copy(), costmap=self._costmap.copy(),
            iter_timeout=self._params.iteration_timeout, current_time=float(f[nat.F_TIME]), current_iter=int(i[nat.I_ITER]),
            robot_collided=bool(i[nat.I_COLLIDED]),
            poses_queue=self._queue(f, i, 'pose'),
            robot_state_queue=[self._state_type.from_row(r) for r in self._queue(f, i, 'state')],
            control_queue=[Action(command=r) for r in self._queue(f, i, 'control')],
            pose=f[nat.F_DPOSE:nat.F_DPOSE + 3].copy(),
            robot_state=self._state_type.from_row(f[nat.F_DROBOT:nat.F_DROBOT + 7]))

    def _from_state(self, state):
        v = self._vec
        L = v.layout
        # The reference's set_state adopts the State wholesale, map and path included (envs/base/env.py:278-285).  Here the
        # map and the path live in device arenas that a snapshot does not carry, so a State taken from ANOTHER world is
        # refused instead of being applied to this env's map and path silently.
        full = v.full_path(0)
        if state.original_path is not None and (np.shape(state.original_path) != np.shape(full) or
                                                not np.array_equal(np.asarray(state.original_path), full)):
            raise ValueError("set_state: the State belongs to an env with another path; build a new env for it")
        if state.costmap is not None and (state.costmap.get_data().shape != self._costmap.get_data().shape or
                                          not np.array_equal(state.costmap.get_data(), self._costmap.get_data()) or
                                          not np.array_equal(state.costmap.get_origin(), self._costmap.get_origin())):
            raise ValueError("set_state: the State belongs to an env with another costmap; build a new env for it")
        if state.iter_timeout is not None and int(state.iter_timeout) != int(self._params.iteration_timeout):
            raise ValueError("set_state: the State was taken with iteration_timeout %s, this env has %s"
                             % (state.iter_timeout, self._params.iteration_timeout))
        f = np.zeros(L.n_frows, dtype=np.float64)
        i = np.zeros(L.n_irows, dtype=np.int32)
        row = state.robot_state.as_row()
        f[nat.F_ROBOT:nat.F_ROBOT + 7] = row           # env.py:284: the robot restarts from the delayed state
        f[nat.F_DROBOT:nat.F_DROBOT + 7] = row
        f[nat.F_DPOSE:nat.F_DPOSE + 3] = state.pose
        f[nat.F_TIME] = state.current_time
        f[nat.F_MIN_DIST] = state.reward_provider_state.min_spat_dist_so_far
        i[nat.I_ITER] = state.current_iter
        i[nat.I_TARGET] = state.reward_provider_state.target_idx
        i[nat.I_COLLIDED] = 1 if state.robot_collided else 0
        for which, items in (('control', [np.asarray(a.command, dtype=np.float64) for a in state.control_queue]),
                             ('pose', [np.asarray(p, dtype=np.float64) for p in state.poses_queue]),
                             ('state', [np.asarray(s.as_row()) for s in state.robot_state_queue])):
            row0, slots, ncomp, irow = v.queue_rows(which)
            if len(items) > slots:
                raise ValueError("%s queue holds %d items but the delay is %d" % (which, len(items), slots))
            for k, item in enumerate(items):
                f[row0 + k * ncomp: row0 + (k + 1) * ncomp] = item
            i[irow] = len(items) << 16
        return VecState(torch.from_numpy(f).reshape(-1, 1), torch.from_numpy(i).reshape(-1, 1))
