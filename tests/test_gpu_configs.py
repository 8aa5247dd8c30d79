"""-m gpu: the other BASELINE.json configurations as parity cases (bench.py measures configs[2] only):
  configs[1]  RandomMiniEnv batched 4096 envs, per-env randomised costmap and obstacle, noise off
  configs[3]  corridor with 3 boxes (synthetic stand-in for the S3 data), industrial tricycle, 16 384 envs
  configs[4]  Monte-Carlo rollouts: get_state / set_state fan-out with per-start-state statistics
Each checks a sample of envs against the oracle and the whole batch through kernel-vs-kernel properties."""
import numpy as np
import pytest
import torch

from bc_gym_planning_env_b200 import _native as nat
from bc_gym_planning_env_b200 import parallel
from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.envs.mini_env import RandomMiniEnvParams, random_mini_pool
from bc_gym_planning_env_b200.envs.rw_corridors.tdwa_test_environments import \
    get_random_maps_squeeze_between_obstacle_in_corridor_on_path
from bc_gym_planning_env_b200.vec_env import VecPlanEnv
from oracle import plan_env_oracle as O
from tests import common

pytestmark = pytest.mark.gpu


def _uniform_actions(env, n_steps, seed):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    low, high = env.action_bounds()
    lo, hi = torch.from_numpy(low).cuda(), torch.from_numpy(high).cuda()
    return [(lo + (hi - lo) * torch.rand((env.n_envs, 2), generator=gen, device="cuda")).contiguous() for _ in range(n_steps)]


def _check_sample_against_oracle(env, oracles, sample, actions):
    for a in actions:
        obs, r, done, _ = env.step(a)
        cand = env._cand[:3].t().contiguous()
        assert torch.equal(env.pose_collides(cand), env.pose_collides(cand, use_u8=True))
        a_host, pose = a[sample].cpu().numpy(), obs.pose[sample].cpu().numpy()
        for k, o in enumerate(oracles):
            oo, r2, d2, _ = o.step(a_host[k])
            np.testing.assert_allclose(pose[k], oo["pose"], rtol=0, atol=1e-9)
            assert float(r[sample[k]]) == r2 and bool(done[sample[k]]) == d2
    env.check_status()


def test_config_1_random_mini_env_batched_4096():
    n = 4096
    ep = EnvParams(goal_ang_dist=np.pi / 8., goal_spat_dist=0.2)
    costmaps, paths = random_mini_pool(n, 0, RandomMiniEnvParams(env_params=ep))
    # the first 12 are the very envs the reference's RandomMiniEnv(seed=0..11) builds (fixture)
    d = common.load("mini_noise_off")
    for s in range(12):
        assert np.array_equal(costmaps[s].get_data(), d["costmap_%d" % s])
    assert all(c.get_data().shape == (183, 183) for c in costmaps)
    env = VecPlanEnv(costmaps, paths, ep, noise_parameters=None)
    sample = np.arange(0, n, n // 16)
    oracles = [O.OraclePlanEnv(costmaps[e].get_data(), costmaps[e].get_origin(), 0.03, paths[e], sp=0.2, ap=np.pi / 8)
               for e in sample]
    _check_sample_against_oracle(env, oracles, sample, _uniform_actions(env, 80, 2))


def test_config_3_corridor_stand_in_16384():
    n = 16384
    original, path, variants = get_random_maps_squeeze_between_obstacle_in_corridor_on_path(n_variants=64, seed=1)
    assert set(np.unique(original.get_data())) == {0, 253, 254, 255} and original.get_data().shape == (267, 667)
    ep = EnvParams(iteration_timeout=1200, pose_delay=1, control_delay=0, state_delay=1, goal_spat_dist=1.0,
                   goal_ang_dist=np.pi / 2, dt=0.05)          # the reference runner's params (rw_randomized_corridor_3_boxes.py:20-29)
    env = VecPlanEnv(list(variants), [path], ep, n_envs=n, map_ids=np.arange(n) % len(variants),
                     path_ids=np.zeros(n, dtype=np.int64), noise_parameters=None, with_ego=True)
    sample = np.arange(0, n, n // 8)
    oracles = [O.OraclePlanEnv(variants[e % len(variants)].get_data(), variants[e % len(variants)].get_origin(), 0.03, path,
                               delays=(0, 1, 1)) for e in sample]
    _check_sample_against_oracle(env, oracles, sample, _uniform_actions(env, 120, 3))
    # the egocentric crop copies every cost value through (255 / 253 / 254 / 0), not only lethal cells
    img = env.ego_image[sample].cpu().numpy()[..., 0]
    for k, o in enumerate(oracles):
        assert np.array_equal(img[k], O.ego_costmap(o.costmap, o.pose, o.origin, o.resolution))
    assert {253, 255}.issubset(set(np.unique(img)))


def test_config_4_monte_carlo_fan_out():
    """README.md:45-61 of the reference, batched: K start states taken mid-episode, each fanned out to R
    noisy rollouts of a fixed action sequence; per-start-state statistics are summed (all-reduced when
    there are several ranks; here world size 1 -- the multi-rank plumbing is covered on gloo)."""
    K, R, H = 8, 64, 40
    params = EnvParams(control_delay=2, pose_delay=1, state_delay=1)
    d = common.load("aisle_delays_211")
    maps = common.fixture_envs(d)[:K]
    from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
    costmaps = [CostMap2D(cm, 0.03, np.array(o)) for cm, o, _ in maps]
    paths = [p for _, _, p in maps]
    params = common.env_params(d["params"], iteration_timeout=1200)
    # start states: step the K envs without noise for a while, snapshot
    src = VecPlanEnv(costmaps, paths, params, noise_parameters=None)
    acts = torch.from_numpy(d["actions"][:K]).cuda()
    for t in range(30):
        src.step(acts[:, t].contiguous())
    start = src.get_state()
    # fan-out env: K * R envs, env g uses the map/path of start state g % K
    n = K * R
    ids = np.arange(n) % K
    fan = VecPlanEnv(costmaps, paths, params, n_envs=n, map_ids=ids, path_ids=ids, seed=99)
    plan = acts[:, 30:30 + H].permute(1, 0, 2).contiguous()          # [H, K, 2]
    out = parallel.monte_carlo_rollouts(fan, start, plan)
    out = out.cpu().numpy()
    assert out.shape == (K, 4) and np.all(out[:, 3] == R)
    # every rollout restarted from its start state: cross-check one rollout per start state with the oracle
    # fed the same Philox stream
    oracles = common.make_oracles(dict(d, n_envs=np.int64(K)), alphas=O.DEFAULT_NOISE,
                                  normal_sources=[(lambda e: (lambda step: O.philox_normal_source(99, e, step)))(e) for e in range(K)])
    rets = np.zeros(K)
    for e, o in enumerate(oracles):
        for t in range(30):
            o.alphas = None
            o.step(d["actions"][e, t])
        o.alphas = O.DEFAULT_NOISE
        o.set_state(o.get_state())
        for h in range(H):
            # the fan-out env's step counter started at 0 when it was created
            o.normal_source = (lambda e: (lambda step, h=h: O.philox_normal_source(99, e, h)))(e)
            _, r, _, _ = o.step(d["actions"][e, 30 + h])
            rets[e] += r
    got_first = np.zeros(K)
    # rerun the fan-out deterministically and read rollout 0 of each start state (env ids 0..K-1)
    fan2 = VecPlanEnv(costmaps, paths, params, n_envs=n, map_ids=ids, path_ids=ids, seed=99)
    cols = parallel.fan_out_columns(K, n).cuda()
    fan2.set_state(type(start)(start.f[:, cols].contiguous(), start.i[:, cols].contiguous()))
    for h in range(H):
        _, r, _, _ = fan2.step(plan[h][cols].contiguous())
        got_first += r[:K].cpu().numpy()
    assert np.array_equal(got_first, rets)
    # noise makes the rollouts of one start state differ from each other
    assert fan.state_f[nat.F_ROBOT, :n].reshape(R, K).std(dim=0).max() > 1e-6
    assert np.all(out[:, 0] >= 0) and np.all(out[:, 1] <= R)
