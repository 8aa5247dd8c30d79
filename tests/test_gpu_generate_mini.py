"""-m gpu: device-side RandomMiniEnv generation (bcg_generate_minis) against the worlds the
unmodified reference sampled (tests/golden/mini_worlds.npz) and against the mini oracle driven by the device's
Philox stream."""
import numpy as np
import pytest
import torch

from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.envs.mini_env import RandomMiniEnvParams
from bc_gym_planning_env_b200.vec_aisle_env import MINI_DTYPE, VecRandomMiniEnv
from oracle import mini_oracle as M
from oracle import plan_env_oracle as O
from tests import common
from tests.test_oracle_mini import FIELDS, unpack

pytestmark = pytest.mark.gpu


def _golden_params(d):
    n = int(d["n_envs"])
    arr = np.zeros(n, dtype=MINI_DTYPE)
    for s in range(n):
        mp = unpack(d["params_%d" % s])
        for name, _ in FIELDS:
            arr[name][s] = mp[name]
    return arr


def _check_world(env, e, costmap, origin, path, target_idx, min_dist):
    cm = env.costmap(e)
    assert cm.get_data().shape == costmap.shape, e
    assert np.array_equal(cm.get_data(), costmap), e
    np.testing.assert_allclose(cm.get_origin(), origin, rtol=0, atol=1e-12)
    got = env.full_path(e)
    assert got.shape == path.shape, (e, got.shape, path.shape)
    np.testing.assert_allclose(got, path, rtol=0, atol=1e-9)
    assert int(env.state_i[1, e]) == target_idx
    np.testing.assert_allclose(float(env.state_f[18, e]), min_dist, rtol=0, atol=1e-9)
    np.testing.assert_allclose(env.state_f[0:3, e].cpu().numpy(), path[0], rtol=0, atol=1e-9)


def test_device_mini_worlds_match_the_reference():
    d = common.load("mini_worlds")
    n = int(d["n_envs"])
    params = _golden_params(d)
    env = VecRandomMiniEnv(n, mini_params=params, noise_parameters=None, with_ego=True)
    for e in range(n):
        _check_world(env, e, d["costmap_%d" % e], d["origin_%d" % e], d["path_%d" % e], int(d["target_idx_%d" % e]),
                     float(d["min_dist_%d" % e]))
    assert np.array_equal(env.mini_params().tobytes(), params.tobytes())
    # derived planes follow the clipped walls: collisions on the tile plane == on the bytes == the oracle; crops too
    rng = np.random.RandomState(3)
    for _ in range(6):
        poses = np.c_[rng.uniform(-2.5, 2.5, (n, 2)), rng.uniform(-np.pi, np.pi, n)]
        flags = env.pose_collides(poses).cpu().numpy()
        assert np.array_equal(flags, env.pose_collides(poses, use_u8=True).cpu().numpy())
        for e in range(n):
            assert flags[e] == O.pose_collides(poses[e, 0], poses[e, 1], poses[e, 2], O.TRICYCLE_FOOTPRINT,
                                               d["costmap_%d" % e], d["origin_%d" % e], 0.03), e
    img, _ = env.observe_ego()
    img = img.cpu().numpy()[..., 0]
    for e in range(n):
        assert np.array_equal(img[e], O.ego_costmap(d["costmap_%d" % e], d["path_%d" % e][0], d["origin_%d" % e], 0.03)), e
    # a second set of worlds into the same slots: the first walls are erased everywhere
    env.generate(mini_params=np.roll(params, 1))
    for e in range(n):
        k = (e - 1) % n
        _check_world(env, e, d["costmap_%d" % k], d["origin_%d" % k], d["path_%d" % k], int(d["target_idx_%d" % k]),
                     float(d["min_dist_%d" % k]))
    img, _ = env.observe_ego()
    img = img.cpu().numpy()[..., 0]
    for e in range(n):
        k = (e - 1) % n
        assert np.array_equal(img[e], O.ego_costmap(d["costmap_%d" % k], d["path_%d" % k][0], d["origin_%d" % k], 0.03)), e
    env.check_status()


def test_device_sampler_follows_the_philox_oracle():
    n, seed, base = 64, 777, 3
    env = VecRandomMiniEnv(n, seed=seed, env_id_base=base, noise_parameters=None)
    for draw in range(2):
        got = env.mini_params()
        attempts = []
        for e in range(n):
            want, k = M.sample_mini_params(M.philox_source(seed, base + e, draw))
            attempts.append(k)
            for name, _ in FIELDS:
                np.testing.assert_allclose(np.asarray(got[name][e]), np.asarray(want[name]), rtol=0, atol=1e-12, err_msg="%d %d %s" % (draw, e, name))
            # the world is built from the parameters the device accepted: CUDA and NumPy sin / cos differ in the last
            # bit, and the circle method's chord (2 x 1.75 m = 70.0 path steps) sits exactly on a refine_path boundary
            mine = {name: (float(got[name][e]) if size == 1 else np.array(got[name][e])) for name, size in FIELDS}
            coarse, costmap, origin = M.mini_world(mine, 0.03)
            path = O.refine_path(coarse, 0.05)
            target, min_dist = O.initial_reward_state(path, 0.2, np.pi / 8.)
            _check_world(env, e, costmap, origin, path, target, min_dist)
        assert max(attempts) > 1                  # the rejection loop was exercised
        env.reset()                               # draw_new_turn_on_reset: draw index 1
    env.check_status()


def test_stepping_generated_mini_worlds_matches_the_oracle():
    n = 32
    gen = RandomMiniEnvParams(env_params=EnvParams(goal_ang_dist=np.pi / 8., goal_spat_dist=0.2, pose_delay=1))
    env = VecRandomMiniEnv(n, gen, seed=21, noise_parameters=None, with_ego=True)
    worlds = [(env.costmap(e), env.full_path(e)) for e in range(n)]
    oracles = [O.OraclePlanEnv(c.get_data(), c.get_origin(), 0.03, p, sp=0.2, ap=np.pi / 8., delays=(0, 1, 0), refine=False)
               for c, p in worlds]
    rng = np.random.RandomState(1)
    low, high = env.action_bounds()
    for t in range(100):
        a = rng.uniform(low, high, size=(n, 2)).astype(np.float32)
        obs, r, done, _ = env.step(a)
        pose, rew, dn = obs.pose.cpu().numpy(), r.cpu().numpy(), done.cpu().numpy()
        for e, o in enumerate(oracles):
            oo, r2, d2, _ = o.step(a[e])
            np.testing.assert_allclose(pose[e], oo["pose"], rtol=0, atol=1e-9)
            assert rew[e] == r2 and bool(dn[e]) == d2, (t, e)
    img = env.ego_image.cpu().numpy()[..., 0]
    for e, o in enumerate(oracles):
        assert np.array_equal(img[e], O.ego_costmap(o.costmap, o.pose, o.origin, o.resolution)), e
    env.check_status()


def test_mini_reset_storm():
    n = 4096                                       # BASELINE.json configs[1]: 4096 envs, per-env randomised costmap
    env = VecRandomMiniEnv(n, seed=5, auto_reset=True)
    first = env.mini_params().copy()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    start.record()
    env.reset()
    stop.record()
    torch.cuda.synchronize()
    second = env.mini_params()
    assert (second["o"] != first["o"]).any(axis=1).all()
    # every accepted world has collision-free end poses (the property the sampler enforces)
    starts = torch.from_numpy(np.ascontiguousarray(second["start"])).cuda()
    ends = torch.from_numpy(np.ascontiguousarray(second["end"])).cuda()
    assert not bool(env.pose_collides(starts).any()) and not bool(env.pose_collides(ends).any())
    assert start.elapsed_time(stop) < 100.0        # ms for 4096 worlds; the host path takes ~40 s
    env.check_status()
