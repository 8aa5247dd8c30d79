"""-m gpu: the odometry-noise path.  The reference draws from the global np.random; the CUDA path draws
Philox4x32-10 normals keyed by (seed; env id, step).  (1) with the oracle fed the *same* Philox draws
the trajectories agree to 1e-9 and flags bit-exactly -- the noise model itself is pinned against the
reference in tests/test_oracle_golden.py; (2) the draws are N(0,1): moments checked statistically."""
import numpy as np
import pytest
import torch

from bc_gym_planning_env_b200 import _native as nat
from bc_gym_planning_env_b200.vec_env import DEFAULT_NOISE
from oracle import plan_env_oracle as O
from tests import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("alphas", [O.DEFAULT_NOISE, (2e-2, 1e-2, 1e-2, 3e-2, 2e-3, 1e-3)])
def test_noisy_steps_match_oracle_with_same_philox_draws(alphas):
    d = common.load("aisle_delays_211")
    seed, base = 0x1234567812345678, 1000
    noise = {"alpha%d" % (k + 1): a for k, a in enumerate(alphas)}
    env = common.make_vec_env(d, noise_parameters=noise, seed=seed, env_id_base=base)
    sources = [(lambda e: (lambda step: O.philox_normal_source(seed, base + e, step)))(e) for e in range(env.n_envs)]
    oracles = common.make_oracles(d, alphas=alphas, normal_sources=sources)
    actions = torch.from_numpy(d["actions"]).cuda()
    for t in range(150):
        obs, r, done, _ = env.step(actions[:, t].contiguous())
        pose, rs = obs.pose.cpu().numpy(), obs.robot_state.cpu().numpy()
        rew, dn = r.cpu().numpy(), done.cpu().numpy()
        for e, o in enumerate(oracles):
            oo, r2, d2, _ = o.step(d["actions"][e, t])
            np.testing.assert_allclose(pose[e], oo["pose"], rtol=0, atol=1e-9)
            np.testing.assert_allclose(rs[e], oo["robot_state"], rtol=0, atol=1e-9)
            assert rew[e] == r2 and bool(dn[e]) == d2, (t, e)
    env.check_status()


def test_noise_moments():
    """A diff-drive robot commanded (v, w) = (0.5, 0) from rest: per differential_drive.py:65-73 its yaw
    increment is (N(0, a3 v^2) + N(0, a5 v^2 + a6 w'^2)) dt -- zero-mean Gaussian of known variance
    (the a6 term is 1e-5 of it) -- independently per env and per step."""
    from bc_gym_planning_env_b200.envs.base.params import EnvParams
    from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
    from bc_gym_planning_env_b200.vec_env import VecPlanEnv
    n = 20000
    cm = CostMap2D(np.zeros((64, 64), np.uint8), 0.03, np.array([-1000., -1000.]))      # nothing to hit
    path = np.array([[0., 0., 0.], [500., 0., 0.]])
    params = EnvParams(robot_name='industrial_diffdrive_v1', refine_path=False, iteration_timeout=100000)
    env = VecPlanEnv([cm], [path], params, n_envs=n, noise_parameters=DEFAULT_NOISE, seed=7)
    a = torch.tensor([[0.5, 0.0]], dtype=torch.float64).repeat(n, 1).cuda()
    std = np.sqrt(1e-2 * 0.5 ** 2 + 1e-3 * 0.5 ** 2) * 0.05
    samples = []
    for _ in range(3):
        th0 = env.state_f[nat.F_ROBOT + 2].clone()
        env.step(a)
        dth = (env.state_f[nat.F_ROBOT + 2] - th0).cpu().numpy()
        samples.append((dth + np.pi) % (2 * np.pi) - np.pi)
    for dth in samples:
        assert abs(dth.mean()) < 5 * std / np.sqrt(n)
        assert abs(dth.std() / std - 1) < 0.03
        k = ((dth - dth.mean()) ** 4).mean() / dth.var() ** 2
        assert abs(k - 3) < 0.25                 # Gaussian kurtosis
        assert len(np.unique(np.round(dth, 14))) > 0.999 * n      # every env draws its own numbers
    # successive steps of one env are uncorrelated
    assert abs(np.corrcoef(samples[0], samples[1])[0, 1]) < 0.05
    # a different seed gives a different stream, the same seed the same stream
    env2 = VecPlanEnv([cm], [path], params, n_envs=n, noise_parameters=DEFAULT_NOISE, seed=7)
    env3 = VecPlanEnv([cm], [path], params, n_envs=n, noise_parameters=DEFAULT_NOISE, seed=8)
    env2.step(a)
    env3.step(a)
    first = samples[0]
    d2 = env2.state_f[nat.F_ROBOT + 2].cpu().numpy()
    d3 = env3.state_f[nat.F_ROBOT + 2].cpu().numpy()
    assert np.allclose(d2, first, rtol=0, atol=1e-15) and not np.allclose(d3, first, rtol=0, atol=1e-6)
