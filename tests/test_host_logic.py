"""CPU: host-side mirrors of the reference interface (value types, map/path builders)."""
import numpy as np
import pytest

from bc_gym_planning_env_b200.envs.base import spaces
from bc_gym_planning_env_b200.envs.base.action import Action
from bc_gym_planning_env_b200.envs.base.params import EnvParams, RewardParams
from bc_gym_planning_env_b200.envs.synth_turn_env import AisleTurnEnvParams, TurnParams, path_and_costmap_from_config, random_aisle_pool
from bc_gym_planning_env_b200.robot_models.robot_dimensions import get_dimensions_example
from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
from bc_gym_planning_env_b200.utilities.path_tools import refine_path
from oracle import plan_env_oracle as O
from tests import common


def test_env_params_defaults_and_roundtrip():
    p = EnvParams()
    assert (p.dt, p.resolution, p.iteration_timeout, p.path_delta) == (0.05, 0.03, 1200, 0.05)   # params.py:17-27
    assert (p.pose_delay, p.control_delay, p.state_delay) == (0, 0, 0)
    assert p.reward_provider_params.spatial_precision == 1.0 and p.reward_provider_params.angular_precision == np.pi / 2
    assert p.reward_provider_params.spatial_progress_multiplier == 0.0
    q = EnvParams(goal_spat_dist=0.2, goal_ang_dist=np.pi / 8, control_delay=2)
    assert q.reward_provider_params.spatial_precision == 0.2
    assert EnvParams.deserialize(q.serialize()) == q
    with pytest.raises(Exception):
        q.dt = 1.0                                                   # frozen like the reference


def test_costmap_2d_contract():
    cm = CostMap2D.create_empty((10, 6), 0.05, (-1, -3))
    assert cm.get_data().shape == (120, 200) and cm.get_data().dtype == np.uint8
    assert not cm.get_origin().flags.writeable
    np.testing.assert_allclose(cm.world_size(), (10, 6))
    np.testing.assert_allclose(cm.world_center(), (4, 0))
    assert cm.world_to_pixel(np.array([0., 0.])).tolist() == [20, 60]
    assert CostMap2D.from_state(cm.get_state()) == cm
    with pytest.raises(AssertionError):
        CostMap2D(np.zeros((2, 2), np.uint8), 0.05, np.array([0, 0]))     # integer origin is rejected


def test_robot_tables():
    tri = get_dimensions_example('industrial_tricycle_v1')
    assert tri.footprint().shape == (16, 2) and tri.footprint()[0, 1] == 0
    assert np.array_equal(tri.footprint(), O.TRICYCLE_FOOTPRINT)
    diff = get_dimensions_example('industrial_diffdrive_v1')
    assert np.array_equal(diff.footprint(), O.DIFFDRIVE_FOOTPRINT)
    with pytest.raises(AssertionError):
        get_dimensions_example('nope')


def test_action_space_sampling_sequence():
    """Box.sample draws from a module RandomState(0) like the reference (spaces.py:9-10, :134-141)."""
    s = 60. * np.pi / 180.
    box = spaces.Box(low=np.array([s / 10, -np.pi / 2]), high=np.array([s / 2, np.pi / 2]), dtype=np.float32)
    spaces.SPACE_LOCAL_RANDOM_STATE.seed(0)
    a = box.sample()
    rng = np.random.RandomState(0)
    want = rng.uniform(low=box.low, high=box.high, size=(2,)).astype(np.float32)
    assert isinstance(a, Action) and a.command.dtype == np.float32 and np.array_equal(a.command, want)
    d = common.load("mini_noise_off")         # the fixture drew its actions the same way
    assert np.array_equal(d["actions"][0, 0], want)


def test_refine_path_matches_oracle():
    rng = np.random.RandomState(0)
    for _ in range(50):
        p = rng.uniform(-5, 5, (rng.randint(2, 6), 3))
        assert np.array_equal(refine_path(p, 0.05), O.refine_path(p, 0.05))
    with pytest.raises(Exception):
        refine_path(np.zeros((3, 4)), 0.05)


def test_aisle_generator_reproduces_reference_maps():
    """random_aisle_pool(seed) must rasterise the very map RandomAisleTurnEnv(seed=seed) builds in the
    reference (fixture aisle_delays_211 holds maps of seeds 0..11, paths already refined)."""
    d = common.load("aisle_delays_211")
    for s in range(int(d["n_envs"])):
        costmaps, paths = random_aisle_pool(1, s)
        assert np.array_equal(costmaps[0].get_data(), d["costmap_%d" % s])
        assert np.array_equal(costmaps[0].get_origin(), d["origin_%d" % s])
        assert np.array_equal(refine_path(paths[0], 0.05), d["path_%d" % s])
    path, cm = path_and_costmap_from_config(AisleTurnEnvParams(turn_params=TurnParams()))
    assert path.shape == (4, 3) and (cm.get_data() == 254).sum() > 500
