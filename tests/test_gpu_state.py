"""-m gpu: get_state / set_state / reset / auto-reset / episode statistics (reference
envs/base/env.py:278-303 and the Monte-Carlo use of README.md:45-61)."""
import numpy as np
import pytest
import torch

from bc_gym_planning_env_b200 import _native as nat
from tests import common

pytestmark = pytest.mark.gpu


def _run(env, actions, t0, t1):
    out = []
    for t in range(t0, t1):
        obs, r, d, _ = env.step(actions[:, t].contiguous())
        out.append((obs.pose.clone(), obs.robot_state.clone(), r.clone(), d.clone(), obs.target_idx.clone()))
    return out


def test_snapshot_restore_continues_bit_identically():
    d = common.load("aisle_delays_211")
    actions = torch.from_numpy(d["actions"]).cuda()
    env = common.make_vec_env(d)
    _run(env, actions, 0, 60)
    snap = env.get_state()
    first = _run(env, actions, 60, 120)
    env.set_state(snap, load_delayed_robot=False)          # exact device-side snapshot
    second = _run(env, actions, 60, 120)
    for a, b in zip(first, second):
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_set_state_follows_the_reference_quirk():
    """env.py:284: set_state loads the *delayed* robot state into the robot.  The oracle does the same."""
    d = common.load("aisle_delays_211")
    actions = torch.from_numpy(d["actions"]).cuda()
    env = common.make_vec_env(d)
    oracles = common.make_oracles(d)
    for t in range(50):
        env.step(actions[:, t].contiguous())
        for e, o in enumerate(oracles):
            o.step(d["actions"][e, t])
    env.set_state(env.get_state())                         # default: reference semantics
    for o in oracles:
        o.set_state(o.get_state())
    true_pose = env.state_f[nat.F_ROBOT:nat.F_ROBOT + 3].t().cpu().numpy()
    delayed = env.state_f[nat.F_DROBOT:nat.F_DROBOT + 3].t().cpu().numpy()
    assert np.array_equal(true_pose, delayed)
    for t in range(50, 90):
        obs, r, done, _ = env.step(actions[:, t].contiguous())
        pose = obs.pose.cpu().numpy()
        for e, o in enumerate(oracles):
            oo, r2, d2, _ = o.step(d["actions"][e, t])
            np.testing.assert_allclose(pose[e], oo["pose"], rtol=0, atol=1e-9)
            assert float(r[e]) == r2 and bool(done[e]) == d2


def test_partial_snapshot_and_masked_reset():
    d = common.load("aisle_delays_211")
    actions = torch.from_numpy(d["actions"]).cuda()
    env = common.make_vec_env(d)
    init = env.get_state()
    _run(env, actions, 0, 30)
    moved = env.get_state()
    env.set_state(env.get_state([1, 4]), [4, 1], load_delayed_robot=False)   # swap two columns
    now = env.get_state()
    assert torch.equal(now.f[:, 4], moved.f[:, 1]) and torch.equal(now.f[:, 1], moved.f[:, 4])
    assert torch.equal(now.f[:, 0], moved.f[:, 0])
    mask = torch.zeros(env.n_envs, dtype=torch.bool)
    mask[[0, 3]] = True
    env.reset(mask)
    after = env.get_state()
    assert torch.equal(after.f[:, 0], init.f[:, 0]) and torch.equal(after.i[:, 3], init.i[:, 3])
    assert torch.equal(after.f[:, 2], moved.f[:, 2])
    env.reset()
    assert torch.equal(env.get_state().f, init.f) and torch.equal(env.get_state().i, init.i)
    with pytest.raises(IndexError):
        env.get_state([env.n_envs])


def test_auto_reset_and_episode_statistics():
    d = common.load("aisle_delays_211")
    actions = torch.from_numpy(d["actions"]).cuda()
    env = common.make_vec_env(d, auto_reset=True)
    init = env.get_state()
    ref_done = d["ref_done"]
    first_done = np.array([np.argmax(ref_done[e]) if ref_done[e].any() else -1 for e in range(env.n_envs)])
    episodes = 0
    for t in range(actions.shape[1]):
        _, r, done, _ = env.step(actions[:, t].contiguous())
        dn = done.cpu().numpy()
        episodes += int(dn.sum())
        for e in np.flatnonzero(dn):
            # the state of a finished env is its initial state again, and its first episode ended where
            # the reference's did
            assert torch.equal(env.state_f[:, e], init.f[:, e]) and torch.equal(env.state_i[:, e], init.i[:, e])
            if first_done[e] >= 0 and t <= first_done[e]:
                assert t == first_done[e]
    stats = dict(zip(nat.STAT_NAMES, env.episode_stats().cpu().numpy()))
    assert stats["episodes"] == episodes and episodes >= (first_done >= 0).sum()
    assert stats["collided"] + stats["goal"] + stats["timeout"] >= stats["episodes"]
    assert stats["length"] >= stats["episodes"]
    env.episode_stats(reset=True)
    assert float(env.episode_stats().abs().sum()) == 0.0


def test_step_host_returns_the_step_results_in_pinned_memory():
    """VecPlanEnv.step_host == step + copies: same rewards, dones and compact observation; the call returns when those are
    on the host (the egocentric kernel may still run), and the images in HBM are valid for what is enqueued next."""
    d = common.load("aisle_delays_211")
    a = common.make_vec_env(d, with_ego=True)
    b = common.make_vec_env(d, with_ego=True)
    actions = torch.from_numpy(d["actions"])                     # [E, T, 2] float32, host
    for t in range(40):
        host = actions[:, t].contiguous().pin_memory()
        dev_copy = host.cuda()
        reward, done, obs = a.step_host(host)
        host.zero_()                                             # the upload is over when the call returns: the buffer is the caller's again
        assert reward.is_pinned() and done.is_pinned() and obs.is_pinned()
        _, r2, d2, _ = b.step(dev_copy)
        assert torch.equal(reward, r2.cpu()) and torch.equal(done.bool(), d2.cpu())
        assert torch.equal(obs, b.obs_vec.cpu())
        assert torch.equal(a.ego_image, b.ego_image)
    host = actions[:, 40].contiguous().pin_memory()
    reward, done, obs, image, goal = a.step_host(host, images=True)   # crops and goal vectors to the host as well
    b.step(host.cuda())
    assert image.is_pinned() and goal.is_pinned()
    assert torch.equal(reward, b.reward.cpu()) and torch.equal(image, b.ego_image.cpu())
    assert torch.equal(goal, b.goal_n_state.cpu())
    with pytest.raises(ValueError):
        a.step_host(actions[:, 0].double())
