"""-m gpu: egocentric crops at arbitrary poses -- deep inside, straddling every map edge and far
outside the map -- against the oracle's closed form of cv2.warpAffine (itself pinned against cv2 in the
CPU suite), for every staging path of the kernel."""
import numpy as np
import pytest
import torch

from bc_gym_planning_env_b200 import _native as nat
from oracle import plan_env_oracle as O
from tests import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("staging", ["tiles", "tma", "spans"])
def test_ego_crops_at_random_poses(staging):
    d = common.load("aisle_collision")
    env = common.make_vec_env(d, with_ego=True, ego_staging=staging)
    maps = common.fixture_envs(d)
    res = float(d["resolution"])
    rng = np.random.RandomState(3)
    n = env.n_envs
    nonzero = 0
    for trial in range(40):
        poses = np.zeros((n, 3))
        for e, (cm, origin, path) in enumerate(maps):
            size = np.array([cm.shape[1], cm.shape[0]]) * res
            kind = trial % 4
            if kind == 0:      # anywhere inside
                poses[e, :2] = origin + rng.uniform(0, 1, 2) * size
            elif kind == 1:    # hugging an edge
                poses[e, :2] = origin + rng.choice([0.0, 1.0], 2) * size + rng.uniform(-1.5, 1.5, 2)
            elif kind == 2:    # on the path
                poses[e, :2] = path[rng.randint(len(path)), :2]
            else:              # well outside
                poses[e, :2] = origin + size / 2 + rng.choice([-1, 1], 2) * (size / 2 + rng.uniform(0, 6, 2))
            poses[e, 2] = rng.uniform(-np.pi, np.pi) if trial % 5 else rng.choice([0, np.pi / 2, -np.pi / 2, -np.pi])
        env.state_f[nat.F_DPOSE:nat.F_DPOSE + 3] = torch.from_numpy(poses.T.copy()).cuda()
        img, _ = env.observe_ego()
        img = img.cpu().numpy()[..., 0]
        for e, (cm, origin, _) in enumerate(maps):
            want = O.ego_costmap(cm, poses[e], origin, res)
            assert np.array_equal(img[e], want), (trial, e, poses[e])
            nonzero += int((want != 0).sum())
    assert nonzero > 1000
    env.check_status()
