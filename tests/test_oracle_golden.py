"""CPU: the oracle against trajectories recorded from the UNMODIFIED reference (tests/golden/*.npz,
regenerate with `python -m oracle.gen_golden` in the build container).  Bit-exact: the oracle runs the
same NumPy primitives as the reference."""
import numpy as np
import pytest

from oracle import plan_env_oracle as O
from tests import common


def _replay(d, oracles):
    T = d["actions"].shape[1]
    for e, o in enumerate(oracles):
        for t in range(T):
            obs, r, done, _ = o.step(d["actions"][e, t])
            assert np.array_equal(obs["pose"], d["ref_pose"][e, t]), (e, t)
            assert np.array_equal(np.array(obs["robot_state"]), d["ref_robot_state"][e, t]), (e, t)
            assert np.array_equal(np.array(o.robot_state[:3]), d["ref_true_pose"][e, t]), (e, t)
            assert r == d["ref_reward"][e, t] and done == d["ref_done"][e, t], (e, t)
            assert o.collided == d["ref_collided"][e, t] and o.target_idx == d["ref_target_idx"][e, t], (e, t)
            assert o.min_dist == d["ref_min_dist"][e, t] and obs["time"] == d["ref_time"][e, t], (e, t)
            assert len(obs["path"]) == d["ref_path_len"][e, t]


@pytest.mark.parametrize("name", ["mini_noise_off", "aisle_delays_211", "aisle_delays_120", "aisle_pure_pursuit", "edge_worlds",
                                  "aisle_goal_reached", "aisle_res005_scale125"])
def test_oracle_matches_reference_rollouts(name):
    d = common.load(name)
    _replay(d, common.make_oracles(d))


def test_oracle_matches_reference_with_noise():
    """The reference draws from the global np.random; the fixture recorded its N(0,1) draws in order
    (slot 1 = angular, slot 2 = final rotation with PlanEnv's alphas), the oracle replays them."""
    d = common.load("aisle_noise_on")
    draws = d["normal_draws"]
    assert np.isnan(draws[..., 2]).all() and not np.isnan(draws[..., :2]).any()   # exactly two draws per step
    sources = []
    for e in range(int(d["n_envs"])):
        def source(step, e=e):
            return lambda slot: draws[e, step, slot - 1]
        sources.append(source)
    _replay(d, common.make_oracles(d, alphas=O.DEFAULT_NOISE, normal_sources=sources))


def test_oracle_collision_and_pixel_counts():
    d = common.load("aisle_collision")
    maps = common.fixture_envs(d)
    res = float(d["resolution"])
    for e, (cm, origin, _) in enumerate(maps):
        for k in range(0, d["poses"].shape[1], 5):
            x, y, th = d["poses"][e, k]
            assert O.pose_collides(x, y, th, O.TRICYCLE_FOOTPRINT, cm, origin, res) == d["ref_flags"][e, k]
            assert O.footprint_pixels_in_map(x, y, th, O.TRICYCLE_FOOTPRINT, cm.shape, origin, res) == d["ref_pixels"][e, k]


@pytest.mark.parametrize("name", ["aisle_ego", "edge_worlds", "aisle_goal_reached"])
def test_oracle_ego_observation(name):
    """Closed form of cv2.warpAffine and the literal cv2 call, both against the reference's
    EgocentricCostmap wrapper -- on aisles, and on the tiny worlds whose crops lie partly or wholly outside the map."""
    d = common.load(name)
    oracles = common.make_oracles(d)
    every = int(d["every"])
    for e, o in enumerate(oracles):
        k = 0
        for t in range(d["actions"].shape[1]):
            obs, _, _, _ = o.step(d["actions"][e, t])
            if "ref_goal_n_state_all" in d:          # the goal vector at every step, across the goal-reached transition
                g = O.goal_n_state(obs["path"], obs["pose"], obs["robot_state"], o.resolution)
                assert np.array_equal(g, d["ref_goal_n_state_all"][e, t]), (e, t)
            if t % every == every - 1:
                want = d["ref_ego_image"][e, k]
                assert np.array_equal(O.ego_costmap(o.costmap, obs["pose"], o.origin, o.resolution), want)
                assert np.array_equal(O.ego_costmap_cv2(o.costmap, obs["pose"], o.origin, o.resolution), want)
                g = O.goal_n_state(obs["path"], obs["pose"], obs["robot_state"], o.resolution)
                assert np.array_equal(g, d["ref_goal_n_state"][e, k])
                k += 1


def test_oracle_diffdrive_steps():
    d = common.load("diffdrive_steps")
    res = float(d["resolution"])
    for e, (cm, origin, path) in enumerate(common.fixture_envs(d)):
        state = [path[0, 0], path[0, 1], path[0, 2], 0., 0., 0., 0.]
        for t in range(d["actions"].shape[1]):
            new = O.diffdrive_step(state, d["actions"][e, t], 0.05)
            hit = O.pose_collides(new[0], new[1], new[2], O.DIFFDRIVE_FOOTPRINT, cm, origin, res)
            if hit:
                new = [state[0], state[1], state[2], 0., 0., 0., 0.]
            state = new
            assert hit == d["ref_hit"][e, t]
            assert np.array_equal(np.array(state[:5]), d["ref_robot_state"][e, t]), (e, t)
