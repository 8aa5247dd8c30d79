"""-m gpu: egocentric observation (crop image + goal_n_state) against the reference's
EgocentricCostmap wrapper (fixture aisle_ego) -- images bit-exact, vector to one fp32 ulp."""
import numpy as np
import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_tma", [True, False])
def test_ego_observation_matches_reference(use_tma):
    """Both ways of staging the source window (TMA box loads, plain span loads) against cv2."""
    d = common.load("aisle_ego")
    env = common.make_vec_env(d, with_ego=True, use_tma=use_tma)
    actions = torch.from_numpy(d["actions"]).cuda()
    every = int(d["every"])
    k = 0
    for t in range(actions.shape[1]):
        env.step(actions[:, t].contiguous())
        if t % every == every - 1:
            img = env.ego_image.cpu().numpy()[..., 0]
            vec = env.goal_n_state.cpu().numpy()[..., 0]
            assert np.array_equal(img, d["ref_ego_image"][:, k]), "ego image differs at step %d" % t
            ref = d["ref_goal_n_state"][:, k]
            assert np.all(np.abs(vec - ref) <= np.spacing(np.abs(ref).astype(np.float32)) + 1e-12), t
            k += 1
    assert k == d["ref_ego_image"].shape[1]
    assert (d["ref_ego_image"] != 0).any()
    env.check_status()


def test_observe_ego_standalone_equals_step_output():
    d = common.load("aisle_ego")
    env = common.make_vec_env(d, with_ego=True)
    actions = torch.from_numpy(d["actions"]).cuda()
    for t in range(10):
        env.step(actions[:, t].contiguous())
    img_step = env.ego_image.clone()
    vec_step = env.goal_n_state.clone()
    img, vec = env.observe_ego()
    assert torch.equal(img, img_step) and torch.equal(vec, vec_step)
