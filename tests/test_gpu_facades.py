"""-m gpu: the reference-shaped single-env API (PlanEnv / RandomMiniEnv / RandomAisleTurnEnv /
EgocentricCostmap) driven the way the reference's own scripts drive it, against the golden fixtures."""
import numpy as np
import pytest

from bc_gym_planning_env_b200.envs.base import spaces
from bc_gym_planning_env_b200.envs.base.action import Action
from bc_gym_planning_env_b200.envs.base.env import PlanEnv, State
from bc_gym_planning_env_b200.envs.base.obs import Observation
from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.envs.egocentric import EgocentricCostmap
from bc_gym_planning_env_b200.envs.mini_env import RandomMiniEnv
from bc_gym_planning_env_b200.envs.synth_turn_env import RandomAisleTurnEnv
from tests import common

pytestmark = pytest.mark.gpu


def test_random_mini_env_replays_config_1():
    """BASELINE config 1: RandomMiniEnv(seed), noise disabled, action_space.sample() actions."""
    d = common.load("mini_noise_off")
    spaces.SPACE_LOCAL_RANDOM_STATE.seed(0)
    for s in range(3):
        env = RandomMiniEnv(seed=s)
        env._env._robot.set_noise_parameters(None)
        assert np.array_equal(env._env._costmap.get_data(), d["costmap_%d" % s])
        for t in range(d["actions"].shape[1]):
            a = env.action_space.sample()
            assert np.array_equal(a.command, d["actions"][s, t])
            obs, r, done, info = env.step(a)
            assert isinstance(obs, Observation) and isinstance(r, float) and isinstance(done, bool) and info == {}
            np.testing.assert_allclose(obs.pose, d["ref_pose"][s, t], rtol=0, atol=1e-9)
            rs = obs.robot_state
            np.testing.assert_allclose([rs.x, rs.y, rs.angle, rs.v, rs.w, rs.steering_motor_command, rs.wheel_angle],
                                       d["ref_robot_state"][s, t], rtol=0, atol=1e-9)
            assert r == d["ref_reward"][s, t] and done == d["ref_done"][s, t]
            assert len(obs.path) == d["ref_path_len"][s, t] and abs(obs.time - d["ref_time"][s, t]) < 1e-12


def test_random_aisle_env_get_set_state():
    d = common.load("aisle_delays_211")
    p = d["params"]
    params = EnvParams(control_delay=2, pose_delay=1, state_delay=1, iteration_timeout=p["timeout"])
    env = RandomAisleTurnEnv(params=params, seed=3, noise_parameters=None)
    e = 3
    for t in range(40):
        obs, r, done, _ = env.step(Action(command=d["actions"][e, t]))
        np.testing.assert_allclose(obs.pose, d["ref_pose"][e, t], rtol=0, atol=1e-9)
    state = env.get_state()
    assert isinstance(state, State) and state.current_iter == 40 and len(state.control_queue) == 2
    assert len(state.poses_queue) == 1 and len(state.robot_state_queue) == 1
    assert state == state.copy() and state.costmap == env._env._costmap
    trace = [env.step(Action(command=d["actions"][e, t])) for t in range(40, 60)]
    env.set_state(state)
    assert env.get_state() == state
    # after set_state the robot restarts from the delayed robot state (env.py:284): the continuation is
    # self-consistent and repeatable
    again = [env.step(Action(command=d["actions"][e, t])) for t in range(40, 60)]
    env.set_state(state)
    third = [env.step(Action(command=d["actions"][e, t])) for t in range(40, 60)]
    for a, b in zip(again, third):
        assert np.array_equal(a[0].pose, b[0].pose) and a[1] == b[1] and a[2] == b[2]
    assert len(trace) == len(again)
    # reset rebuilds a fresh random turn like the reference
    obs0 = env.reset()
    assert isinstance(obs0, Observation) and obs0.time == 0.0


def test_egocentric_wrapper():
    d = common.load("aisle_ego")
    base = RandomAisleTurnEnv(params=EnvParams(pose_delay=1, state_delay=1), seed=100, noise_parameters=None)
    env = EgocentricCostmap(base)
    every = int(d["every"])
    k = 0
    for t in range(d["actions"].shape[1]):
        obs, r, done, _ = env.step(Action(command=d["actions"][0, t]))
        assert obs["env"].shape == (133, 117, 1) and obs["env"].dtype == np.uint8
        assert obs["goal_n_state"].shape == (9, 1) and obs["goal_n_state"].dtype == np.float32
        if t % every == every - 1:
            assert np.array_equal(obs["env"][..., 0], d["ref_ego_image"][0, k])
            np.testing.assert_allclose(obs["goal_n_state"][:, 0], d["ref_goal_n_state"][0, k], rtol=1e-6, atol=1e-7)
            k += 1


def test_plan_env_raises_like_the_reference():
    from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
    cm = CostMap2D(np.zeros((50, 50), np.uint8), 0.03, np.array([0., 0.]))
    with pytest.raises(ValueError, match="Goal pose too close to initial pose"):
        PlanEnv(cm, np.array([[0.5, 0.5, 0.], [0.6, 0.5, 0.]]), EnvParams())          # reward.py:275-277


def test_save_n_load_identity():
    """The reference's serialization contract (scripts/env_runners/save_n_load_checks.py:16-71): serialize to a
    dict of basic types, pickle, unpickle, deserialize -- state, observation, reward and done of the clone stay
    equal to the original's for 100 steps."""
    import pickle
    from bc_gym_planning_env_b200.envs.synth_turn_env import AisleTurnEnv, AisleTurnEnvParams, TurnParams

    def roundtrip(thing, the_class):
        return the_class.deserialize(pickle.loads(pickle.dumps(thing.serialize(), protocol=pickle.HIGHEST_PROTOCOL)))

    # state_delay stays 0: a State carries only the *delayed* robot state (env.py:52-68, :284), so with a state
    # delay neither the reference nor this port can resume the true robot from a saved State
    cfg = AisleTurnEnvParams(turn_params=TurnParams(), env_params=EnvParams(control_delay=2, pose_delay=1, state_delay=0))
    env = AisleTurnEnv(cfg, noise_parameters=None)
    env.reset()
    for _ in range(7):                       # so that the queues are not empty when the env is saved
        env.step(env.action_space.sample())
    env_two = roundtrip(env, PlanEnv)
    for action in [env.action_space.sample() for _ in range(100)]:
        assert env.get_state() == env_two.get_state()
        obs1, r1, done1, _ = env.step(action)
        obs2, r2, done2, _ = env_two.step(action)
        obs3 = roundtrip(obs2, Observation)
        assert env.get_state() == env_two.get_state()
        assert obs1 == obs2 and obs1 == obs3 and r1 == r2 and done1 == done2
    state = env.get_state()
    assert roundtrip(state, State) == state


def test_colored_observation_envs():
    from bc_gym_planning_env_b200.envs.synth_turn_env import (ColoredCostmapRandomAisleTurnEnv,
                                                             ColoredEgoCostmapRandomAisleTurnEnv)
    d = common.load("aisle_colored_ego")
    env = ColoredEgoCostmapRandomAisleTurnEnv(seed=None, noise_parameters=None)
    env.seed(400)
    obs = env.reset()                                   # same construction order as the fixture generator
    assert obs['environment'].shape == (133, 133, 1) and obs['goal'].shape == (5, 1) and obs['goal'].dtype == np.float64
    every, k = int(d["every"]), 0
    for t in range(d["actions"].shape[1]):
        obs, r, done, _ = env.step(Action(command=d["actions"][0, t]))
        if t % every == every - 1:
            assert np.array_equal(obs['environment'][..., 0], d["ref_environment"][0, k])
            np.testing.assert_allclose(obs['goal'][:, 0], d["ref_goal"][0, k], rtol=2e-7, atol=1e-7)
            k += 1
    full = ColoredCostmapRandomAisleTurnEnv(seed=5, noise_parameters=None)
    o = full.reset()
    assert o.ndim == 3 and o.shape[-1] == 1 and o.dtype == np.uint8 and (o == 254).any()
    o2, r, done, info = full.step(full.action_space.sample())
    assert o2.shape == o.shape


def test_pure_pursuit_provider_through_the_facade():
    d = common.load("aisle_pure_pursuit")
    p = d["params"]
    params = EnvParams(control_delay=1, pose_delay=1, state_delay=1, iteration_timeout=p["timeout"],
                       reward_provider_name='continuous_reward_pure_pursuit')
    env = RandomAisleTurnEnv(params=params, seed=90, noise_parameters=None)
    for t in range(120):
        obs, r, done, _ = env.step(Action(command=d["actions"][0, t]))
        assert abs(r - d["ref_reward"][0, t]) < 1e-9 and done == d["ref_done"][0, t]
        assert len(obs.path) == d["ref_path_len"][0, t]
    state = env.get_state()
    assert state.reward_provider_state.get_reward_provider_state_type_name() == 'continuous_reward_pure_pursuit_state'
    assert State.deserialize(state.serialize()) == state
    with pytest.raises(AssertionError):
        RandomAisleTurnEnv(params=EnvParams(reward_provider_name='nope'), seed=1)
