"""-m gpu: the diff-drive robot model (reference robot_models/differential_drive.py:236-265 stepped
through envs/base/env.py:442-461) and collision checks with an arbitrary footprint polygon."""
import numpy as np
import pytest
import torch

from bc_gym_planning_env_b200 import _native as nat
from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
from bc_gym_planning_env_b200.vec_env import VecPlanEnv
from tests import common

pytestmark = pytest.mark.gpu


def test_diffdrive_steps_match_reference():
    d = common.load("diffdrive_steps")
    params = EnvParams(robot_name='industrial_diffdrive_v1', refine_path=False, iteration_timeout=100000)
    env = common.make_vec_env(d, params=params)
    actions = torch.from_numpy(d["actions"]).cuda()            # float64 (v, w)
    for t in range(actions.shape[1]):
        env.step(actions[:, t].contiguous())
        got = env.state_f[nat.F_ROBOT:nat.F_ROBOT + 5].t().cpu().numpy()
        np.testing.assert_allclose(got, d["ref_robot_state"][:, t], rtol=0, atol=1e-9)
        assert np.array_equal(env.hit.cpu().numpy(), d["ref_hit"][:, t]), t
    assert d["ref_hit"].any()
    env.check_status()


def test_custom_footprint_collision_kat():
    """The reference's 20 golden poses + 1000 random ones (utilities/test_costmap_utils.py:251-325) with its
    rectangular test footprint at 0.05 m: `pose_collides` verdicts, and `is_robot_colliding`'s extra
    robot-centre-in-bounds rule applied on the host."""
    d = common.load("kat_is_robot_colliding")
    poses = d["poses"]
    n = len(poses)
    cm = CostMap2D(d["costmap"], 0.05, d["origin"])
    path = np.array([[0., 0., 0.], [100., 0., 0.]])
    env = VecPlanEnv([cm], [path], EnvParams(refine_path=False), n_envs=n, noise_parameters=None, footprint=d["footprint"])
    for use_u8 in (False, True):
        flags = env.pose_collides(poses, use_u8=use_u8).cpu().numpy()
        assert np.array_equal(flags, d["ref_pose_collides"])
        px = cm.world_to_pixel(poses[:, :2])
        centre_in = (px[:, 0] >= 0) & (px[:, 0] < cm.get_data().shape[1]) & (px[:, 1] >= 0) & (px[:, 1] < cm.get_data().shape[0])
        assert np.array_equal(flags & centre_in, d["ref_is_robot_colliding"])
    env.check_status()
