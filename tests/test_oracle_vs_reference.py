"""CPU, build container only: the oracle stepped side by side with the UNMODIFIED reference imported
read-only from /root/reference (skipped where that tree does not exist, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle import plan_env_oracle as O
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


def _oracle_for(pe, alphas):
    st, ep = pe._state, pe._params
    rp = ep.reward_provider_params
    return O.OraclePlanEnv(st.costmap.get_data(), st.costmap.get_origin(), st.costmap.get_resolution(), st.original_path,
                           dt=ep.dt, sp=rp.spatial_precision, ap=rp.angular_precision, multiplier=rp.spatial_progress_multiplier,
                           timeout=ep.iteration_timeout, delays=(ep.control_delay, ep.pose_delay, ep.state_delay),
                           alphas=alphas, refine=False)


@pytest.mark.parametrize("noise", [False, True])
def test_oracle_tracks_live_reference(noise):
    load_reference()
    from bc_gym_planning_env.envs.base.action import Action
    from bc_gym_planning_env.envs.base.params import EnvParams
    from bc_gym_planning_env.envs.synth_turn_env import RandomAisleTurnEnv
    env = RandomAisleTurnEnv(params=EnvParams(control_delay=1, pose_delay=2, state_delay=1, iteration_timeout=150), seed=99)
    pe = env._env
    if not noise:
        pe._robot.set_noise_parameters(None)
    o = _oracle_for(pe, O.DEFAULT_NOISE if noise else None)
    rng = np.random.RandomState(5)
    for t in range(250):
        a = rng.uniform(pe.action_space.low, pe.action_space.high).astype(np.float32)
        np.random.seed(t)
        obs, r, d, _ = env.step(Action(command=a))
        np.random.seed(t)                       # same global-RNG draws for the oracle's default normal source
        oo, r2, d2, _ = o.step(a)
        rs = obs.robot_state
        assert np.array_equal(obs.pose, oo["pose"]) and r == r2 and d == d2
        assert [rs.x, rs.y, rs.angle, rs.v, rs.w, rs.steering_motor_command, rs.wheel_angle] == oo["robot_state"]
        assert pe._state.reward_provider_state.target_idx == o.target_idx and pe._state.robot_collided == o.collided


def test_reference_kat_suites_still_pass_under_the_shim():
    """The reference's own helper KATs the oracle is pinned to run green here."""
    load_reference()
    from bc_gym_planning_env.utilities import test_coordinate_transformations as tct
    from bc_gym_planning_env.utilities import test_path_tools as tpt
    tct.test_world_to_pixel()
    tct.test_normalize_angles()
    tpt.test_path_velocity()
    tpt.test_pose_distances()
    tpt.test_compute_robot_area()
