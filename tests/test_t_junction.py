"""T-junction worlds: the host-side builder against worlds of the unmodified reference
(tests/golden/t_junction.npz), and -- on the GPU -- stepping through one against the oracle."""
import json

import numpy as np
import pytest

from bc_gym_planning_env_b200.envs.t_junction_env import Bearing, TJunction
from tests import common

MOUTHS = ("left", "right", "bottom")


def _cases(d):
    return json.loads(str(d["cases"]))


def test_t_junction_worlds_match_the_reference():
    pytest.importorskip("cv2")
    d = common.load("t_junction")
    for i, kw in enumerate(_cases(d)):
        tj = TJunction(**kw)
        assert np.array_equal(tj.wall_corners, d["corners_%d" % i])
        cm = tj.get_costmap(0.03)
        assert np.array_equal(cm.get_data(), d["costmap_%d" % i]), i
        assert np.array_equal(cm.get_origin(), d["origin_%d" % i]), i
        for a in MOUTHS:
            for b in MOUTHS:
                if a != b:
                    np.random.seed(12 + i)
                    path = tj.get_path(a, b)
                    assert path.shape == (300, 3)
                    assert np.array_equal(path, d["path_%d_%s_%s" % (i, a, b)]), (i, a, b)


def test_t_junction_rejects_what_the_reference_rejects():
    for kw in (dict(column_width=0.0), dict(beam_width=-1.0), dict(window_height=0.0), dict(window_width=0.0),
               dict(column_width=11.0), dict(beam_width=10.5)):
        with pytest.raises(ValueError):
            TJunction(**kw)
    tj = TJunction()
    with pytest.raises(TypeError):
        tj.get_path(1, "left")
    with pytest.raises(ValueError):
        tj.get_path("left", "left")
    with pytest.raises(ValueError):
        tj.get_path("top", "left")
    assert Bearing.starting["bottom"] == np.pi / 2 and Bearing.ending["right"] == 0.0
    assert len(tj.obstacles) == 5


@pytest.mark.gpu
def test_stepping_through_a_t_junction_matches_the_oracle():
    import torch
    from bc_gym_planning_env_b200.envs.base.params import EnvParams
    from bc_gym_planning_env_b200.vec_env import VecPlanEnv
    from oracle import plan_env_oracle as O
    tj = TJunction(window_height=8.0, window_width=12.5, column_width=2.1, beam_width=1.3)
    cm = tj.get_costmap(0.03)
    paths = [tj.get_path(a, b) for a in MOUTHS for b in MOUTHS if a != b]
    ep = EnvParams(pose_delay=1, state_delay=1)
    env = VecPlanEnv([cm], paths, ep, noise_parameters=None, with_ego=True)
    oracles = [O.OraclePlanEnv(cm.get_data(), cm.get_origin(), 0.03, p, delays=(0, 1, 1)) for p in paths]
    rng = np.random.RandomState(2)
    low, high = env.action_bounds()
    hits = 0
    for t in range(150):
        a = rng.uniform(low, high, size=(env.n_envs, 2)).astype(np.float32)
        obs, r, done, _ = env.step(a)
        pose, rew, dn = obs.pose.cpu().numpy(), r.cpu().numpy(), done.cpu().numpy()
        for e, o in enumerate(oracles):
            oo, r2, d2, _ = o.step(a[e])
            np.testing.assert_allclose(pose[e], oo["pose"], rtol=0, atol=1e-9)
            assert rew[e] == r2 and bool(dn[e]) == d2, (t, e)
        hits += int(env.hit.sum())
    img = env.ego_image.cpu().numpy()[..., 0]
    for e, o in enumerate(oracles):
        assert np.array_equal(img[e], O.ego_costmap(o.costmap, o.pose, o.origin, o.resolution)), e
    assert hits > 0                       # random driving in a 2 m column does hit the walls
    env.check_status()
