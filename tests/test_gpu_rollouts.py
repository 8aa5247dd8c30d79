"""-m gpu: the CUDA step (through the C-ABI) against trajectories recorded from the unmodified
reference (tests/golden, generator oracle/gen_golden.py).  Flags bit-exact, floats to 1e-9
(contract 1e-5 relative)."""
import numpy as np
import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu


def _rollout(d, **kw):
    env = common.make_vec_env(d, **kw)
    actions = torch.from_numpy(d["actions"]).cuda()      # [E, T, 2]
    E, T = actions.shape[:2]
    rec = {k: [] for k in ("pose", "robot_state", "true_pose", "reward", "done", "collided", "target_idx", "min_dist", "time")}
    from bc_gym_planning_env_b200 import _native as nat
    for t in range(T):
        obs, r, done, _ = env.step(actions[:, t].contiguous())
        rec["pose"].append(obs.pose.cpu().numpy().copy())
        rec["robot_state"].append(obs.robot_state.cpu().numpy().copy())
        rec["true_pose"].append(env.state_f[nat.F_ROBOT:nat.F_ROBOT + 3].t().cpu().numpy().copy())
        rec["reward"].append(r.cpu().numpy().copy())
        rec["done"].append(done.cpu().numpy().copy())
        rec["collided"].append(env.state_i[nat.I_COLLIDED].cpu().numpy().astype(bool))
        rec["target_idx"].append(obs.target_idx.cpu().numpy().copy())
        rec["min_dist"].append(env.state_f[nat.F_MIN_DIST].cpu().numpy().copy())
        rec["time"].append(obs.time.cpu().numpy().copy())
    env.check_status()
    return {k: np.swapaxes(np.array(v), 0, 1) for k, v in rec.items()}


@pytest.mark.parametrize("name", ["mini_noise_off", "aisle_delays_211", "aisle_delays_120", "aisle_pure_pursuit", "edge_worlds",
                                  "aisle_goal_reached", "aisle_res005_scale125"])
def test_step_matches_reference(name):
    d = common.load(name)
    got = _rollout(d)
    for k in ("done", "collided", "target_idx"):
        assert np.array_equal(got[k], d["ref_" + k]), "flag %s differs from the reference" % k
    for k in ("pose", "robot_state", "true_pose", "reward", "min_dist", "time"):
        ref = d["ref_" + k]
        np.testing.assert_allclose(got[k], ref, rtol=common.CONTRACT_RTOL, atol=common.CONTRACT_RTOL)
        np.testing.assert_allclose(got[k], ref, rtol=0, atol=common.TIGHT_ATOL, err_msg=k)
    assert d["ref_collided"].any() and d["ref_done"].any()
    assert (d["ref_reward"] == 1.0).any() or (d["ref_reward"] < -99).any()
    if name == "aisle_goal_reached":          # every env finishes its path and is stepped on past done (reward.py:223-224)
        assert (d["ref_path_len"][:, -1] == 0).all() and (got["reward"][:, -40:] == 0).all()


def test_f64_actions_and_shared_pools_agree():
    d = common.load("aisle_delays_211")
    d64 = dict(d)
    d64["actions"] = d["actions"].astype(np.float64)
    a = _rollout(d)
    b = _rollout(d64)
    for k in a:
        assert np.array_equal(a[k], b[k])
    c = _rollout(d, private_map_copies=True)
    for k in a:
        assert np.array_equal(a[k], c[k])
