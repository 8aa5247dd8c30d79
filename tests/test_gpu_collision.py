"""-m gpu: batched pose_collides (lethal tile plane and raw uint8 rows) against the reference's
verdicts and in-map footprint pixel counts."""
import numpy as np
import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu


def test_collision_flags_and_pixel_counts_match_reference():
    d = common.load("aisle_collision")
    env = common.make_vec_env(d)
    poses = d["poses"]                                    # [E, K, 3]
    for k in range(poses.shape[1]):
        p = torch.from_numpy(poses[:, k]).cuda()
        flags, pixels = env.pose_collides(p, count_pixels=True)
        flags_u8 = env.pose_collides(p, use_u8=True)
        assert np.array_equal(flags.cpu().numpy(), d["ref_flags"][:, k])
        assert np.array_equal(flags_u8.cpu().numpy(), d["ref_flags"][:, k])
        assert np.array_equal(pixels.cpu().numpy(), d["ref_pixels"][:, k])
    env.check_status()
    assert d["ref_flags"].any() and not d["ref_flags"].all()
