"""-m gpu: device-side AisleTurnEnv generation (bcg_generate_aisles) against the worlds the
unmodified reference built (tests/golden/aisle_worlds.npz) and against the aisle oracle."""
import numpy as np
import pytest
import torch

from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.vec_aisle_env import TURN_DTYPE, VecRandomAisleTurnEnv, turn_params_array
from oracle import aisle_oracle as A
from oracle import plan_env_oracle as O
from tests import common

pytestmark = pytest.mark.gpu


def _golden_turns(d):
    out = []
    for row in d["turn_params"]:
        tp = dict(zip(A.TURN_FIELDS, row))
        tp["flip_arnd_oy"], tp["flip_arnd_ox"] = bool(tp["flip_arnd_oy"]), bool(tp["flip_arnd_ox"])
        out.append(tp)
    return out


def _check_world(env, e, costmap, origin, path, target_idx, min_dist):
    cm = env.costmap(e)
    assert cm.get_data().shape == costmap.shape, e
    assert np.array_equal(cm.get_data(), costmap), e
    np.testing.assert_allclose(cm.get_origin(), origin, rtol=0, atol=1e-12)
    got = env.full_path(e)
    assert got.shape == path.shape, e
    np.testing.assert_allclose(got, path, rtol=0, atol=1e-9)
    assert int(env.state_i[1, e]) == target_idx
    np.testing.assert_allclose(float(env.state_f[18, e]), min_dist, rtol=0, atol=1e-9)
    np.testing.assert_allclose(env.state_f[0:3, e].cpu().numpy(), path[0], rtol=0, atol=1e-9)


def test_device_worlds_match_the_reference():
    d = common.load("aisle_worlds")
    n = int(d["n_envs"])
    env = VecRandomAisleTurnEnv(n, EnvParams(), turn_params=_golden_turns(d), noise_parameters=None, with_ego=True)
    for e in range(n):
        _check_world(env, e, d["costmap_%d" % e], d["origin_%d" % e], d["path_%d" % e], int(d["target_idx_%d" % e]),
                     float(d["min_dist_%d" % e]))
    back = env.turn_params()
    assert np.array_equal(back, turn_params_array(_golden_turns(d)))
    # the derived planes follow the drawn pixels: collision (lethal tile plane) == collision on the uint8 rows ==
    # the oracle on the reference's map; egocentric crop (cell tiles) == the oracle's crop
    rng = np.random.RandomState(3)
    for _ in range(6):
        poses = np.zeros((n, 3))
        for e in range(n):
            path = d["path_%d" % e]
            k = rng.randint(len(path))
            poses[e] = path[k] + np.r_[rng.uniform(-1.5, 1.5, 2), rng.uniform(-np.pi, np.pi)]
        flags = env.pose_collides(poses).cpu().numpy()
        assert np.array_equal(flags, env.pose_collides(poses, use_u8=True).cpu().numpy())
        for e in range(n):
            assert flags[e] == O.pose_collides(poses[e, 0], poses[e, 1], poses[e, 2], O.TRICYCLE_FOOTPRINT,
                                               d["costmap_%d" % e], d["origin_%d" % e], 0.03), e
    img, _ = env.observe_ego()
    img = img.cpu().numpy()[..., 0]
    for e in range(n):
        assert np.array_equal(img[e], O.ego_costmap(d["costmap_%d" % e], d["path_%d" % e][0], d["origin_%d" % e], 0.03)), e
    env.check_status()


def test_regeneration_rewrites_slots_in_place():
    d = common.load("aisle_worlds")
    n = int(d["n_envs"])
    turns = _golden_turns(d)
    env = VecRandomAisleTurnEnv(n, EnvParams(), seed=11, noise_parameters=None, with_ego=True)   # drawn worlds first
    first = [env.costmap(e).get_data() for e in range(n)]
    mask = np.zeros(n, dtype=bool)
    mask[::2] = True
    shuffled = turns[1:] + turns[:1]                         # env e now gets the world of fixture e + 1
    env.generate(mask, turn_params=shuffled)
    for e in range(n):
        if mask[e]:
            k = (e + 1) % n
            _check_world(env, e, d["costmap_%d" % k], d["origin_%d" % k], d["path_%d" % k], int(d["target_idx_%d" % k]),
                         float(d["min_dist_%d" % k]))
        else:
            assert np.array_equal(env.costmap(e).get_data(), first[e]), e
    # erased pixels are gone from the derived planes too: a third world on top, then check crops and collisions
    env.generate(turn_params=turns)
    img, _ = env.observe_ego()
    img = img.cpu().numpy()[..., 0]
    poses = np.stack([d["path_%d" % e][len(d["path_%d" % e]) // 2] + np.r_[0.6, -0.4, 0.3] for e in range(n)])
    flags = env.pose_collides(poses).cpu().numpy()
    for e in range(n):
        _check_world(env, e, d["costmap_%d" % e], d["origin_%d" % e], d["path_%d" % e], int(d["target_idx_%d" % e]),
                     float(d["min_dist_%d" % e]))
        assert np.array_equal(img[e], O.ego_costmap(d["costmap_%d" % e], d["path_%d" % e][0], d["origin_%d" % e], 0.03)), e
        assert flags[e] == O.pose_collides(poses[e, 0], poses[e, 1], poses[e, 2], O.TRICYCLE_FOOTPRINT,
                                           d["costmap_%d" % e], d["origin_%d" % e], 0.03), e
    env.check_status()


def test_device_draws_follow_the_philox_oracle():
    n, seed, base = 96, 20250917, 5
    env = VecRandomAisleTurnEnv(n, EnvParams(), seed=seed, env_id_base=base, noise_parameters=None)
    for draw in range(2):
        got = env.turn_params()
        for e in range(n):
            want = A.philox_turn_params(seed, base + e, draw)
            for f in TURN_DTYPE.names:
                assert float(got[f][e]) == float(want[f]), (draw, e, f)
        if draw == 0:
            # the drawn worlds are the oracle's worlds for those parameters
            for e in range(0, n, 12):
                tp = A.philox_turn_params(seed, base + e, 0)
                coarse, costmap, origin = A.aisle_world(tp, 0.03)
                path = O.refine_path(coarse, 0.05)
                target, min_dist = O.initial_reward_state(path, 1.0, np.pi / 2)
                _check_world(env, e, costmap, origin, path, target, min_dist)
            env.reset()                                        # draw_new_turn_on_reset: draw index 1
    env.check_status()


def test_stepping_generated_worlds_matches_the_oracle():
    n = 32
    ep = EnvParams(control_delay=2, pose_delay=1, state_delay=1)
    env = VecRandomAisleTurnEnv(n, ep, seed=4, noise_parameters=None, with_ego=True)
    worlds = [(env.costmap(e), env.full_path(e)) for e in range(n)]
    oracles = [O.OraclePlanEnv(c.get_data(), c.get_origin(), 0.03, p, delays=(2, 1, 1), refine=False) for c, p in worlds]
    rng = np.random.RandomState(1)
    low, high = env.action_bounds()
    for t in range(120):
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


def test_reset_storm_draws_new_worlds_for_done_envs_only():
    n = 512
    env = VecRandomAisleTurnEnv(n, EnvParams(iteration_timeout=40), seed=9, auto_reset=True)
    before = env.turn_params().copy()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    low, high = env.action_bounds()
    lo, hi = torch.from_numpy(low).cuda(), torch.from_numpy(high).cuda()
    changed = np.zeros(n, dtype=bool)
    for t in range(60):
        a = (lo + (hi - lo) * torch.rand((n, 2), generator=gen, device="cuda")).contiguous()
        _, _, done, _ = env.step(a)
        if bool(done.any()):
            dn = done.clone()
            env.reset(dn)
            changed |= dn.cpu().numpy()
    after = env.turn_params()
    assert changed.all()                                       # the 40-step timeout ends every episode
    assert (after["rot_theta"] != before["rot_theta"]).all()
    # a regenerated env is a consistent fresh episode: iter 0, at the start of its own path
    e = int(np.flatnonzero(changed)[0])
    path = env.full_path(e)
    coarse, costmap, origin = A.aisle_world({f: (bool(after[f][e]) if f.startswith("flip") else float(after[f][e]))
                                             for f in TURN_DTYPE.names}, 0.03)
    assert np.array_equal(env.costmap(e).get_data(), costmap)
    np.testing.assert_allclose(path, O.refine_path(coarse, 0.05), rtol=0, atol=1e-9)
    env.check_status()


def test_other_resolution_and_footprint_scale_step_parity():
    """Nothing in the kernels is specialised to 0.03 m cells or the 117 x 133 crop: a 0.05 m world (70 x 80 crop)."""
    from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
    from bc_gym_planning_env_b200.vec_env import VecPlanEnv
    res = 0.05
    rng = np.random.RandomState(8)
    worlds = []
    for _ in range(6):
        coarse, costmap, origin = A.aisle_world(A.draw_turn_params(rng), res)
        worlds.append((CostMap2D(costmap, res, origin), coarse))
    ep = EnvParams(control_delay=1, pose_delay=1, state_delay=0)
    env = VecPlanEnv([c for c, _ in worlds], [p for _, p in worlds], ep, noise_parameters=None, with_ego=True)
    assert tuple(env.ego_image.shape[1:]) == (80, 70, 1)
    oracles = [O.OraclePlanEnv(c.get_data(), c.get_origin(), res, p, delays=(1, 1, 0)) for c, p in worlds]
    low, high = env.action_bounds()
    for t in range(100):
        a = rng.uniform(low, high, size=(env.n_envs, 2)).astype(np.float32)
        obs, r, done, _ = env.step(a)
        pose, rew, dn = obs.pose.cpu().numpy(), r.cpu().numpy(), done.cpu().numpy()
        for e, o in enumerate(oracles):
            oo, r2, d2, _ = o.step(a[e])
            np.testing.assert_allclose(pose[e], oo["pose"], rtol=0, atol=1e-9)
            assert rew[e] == r2 and bool(dn[e]) == d2, (t, e)
        if t % 10 == 9:
            img = env.ego_image.cpu().numpy()[..., 0]
            for e, o in enumerate(oracles):
                assert np.array_equal(img[e], O.ego_costmap(o.costmap, o.pose, o.origin, res)), (t, e)
    env.check_status()


def test_device_worlds_at_another_resolution():
    n, seed = 12, 77
    env = VecRandomAisleTurnEnv(n, EnvParams(), seed=seed, noise_parameters=None, resolution=0.05)
    for e in range(n):
        coarse, costmap, origin = A.aisle_world(A.philox_turn_params(seed, e, 0), 0.05)
        path = O.refine_path(coarse, 0.05)
        target, min_dist = O.initial_reward_state(path, 1.0, np.pi / 2)
        _check_world(env, e, costmap, origin, path, target, min_dist)
    env.check_status()
