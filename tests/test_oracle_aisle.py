"""The aisle-generation oracle (oracle/aisle_oracle.py) against cv2.line and against worlds built by the
unmodified reference (tests/golden/aisle_worlds.npz)."""
import numpy as np
import pytest

from oracle import aisle_oracle as A
from oracle import plan_env_oracle as O
from tests.common import load

cv2 = pytest.importorskip("cv2")


def _turn(row):
    tp = dict(zip(A.TURN_FIELDS, row))
    tp["flip_arnd_oy"], tp["flip_arnd_ox"] = bool(tp["flip_arnd_oy"]), bool(tp["flip_arnd_ox"])
    return tp


def test_line_pixels_match_cv2():
    rng = np.random.RandomState(5)
    cases = [(0, 0, 0, 0), (3, 4, 3, 40), (3, 40, 3, 4), (5, 7, 60, 7), (60, 7, 5, 7), (0, 0, 63, 63), (63, 0, 0, 63),
             (10, 10, 11, 50), (10, 50, 11, 10), (10, 10, 50, 11), (50, 10, 10, 11)]
    cases += [tuple(rng.randint(0, 64, size=4)) for _ in range(3000)]
    for x0, y0, x1, y1 in cases:
        img = np.zeros((64, 64), dtype=np.uint8)
        cv2.line(img, (int(x0), int(y0)), (int(x1), int(y1)), color=254, thickness=1)
        mine = np.zeros_like(img)
        px = A.line_pixels(int(x0), int(y0), int(x1), int(y1))
        mine[px[:, 1], px[:, 0]] = 254
        assert np.array_equal(img, mine), (x0, y0, x1, y1)
        assert len(px) == max(abs(x1 - x0), abs(y1 - y0)) + 1


def test_aisle_world_matches_reference():
    d = load("aisle_worlds")
    res = float(d["resolution"])
    for i in range(int(d["n_envs"])):
        coarse, costmap, origin = A.aisle_world(_turn(d["turn_params"][i]), res)
        assert np.array_equal(costmap, d["costmap_%d" % i]), i
        assert np.array_equal(origin, d["origin_%d" % i]), i
        assert np.array_equal(coarse, d["coarse_%d" % i]), i
        path = O.refine_path(coarse, 0.05)
        assert np.array_equal(path, d["path_%d" % i]), i
        target, min_dist = O.initial_reward_state(path, 1.0, np.pi / 2)
        assert target == int(d["target_idx_%d" % i]) and min_dist == float(d["min_dist_%d" % i])


def test_turn_param_draws_follow_the_reference_order():
    d = load("aisle_worlds")
    for i in range(int(d["n_envs"]) - 4):                  # the last four fixtures were reset twice
        tp = A.draw_turn_params(np.random.RandomState(300 + i))
        assert [float(tp[f]) for f in A.TURN_FIELDS] == list(d["turn_params"][i])
    rng = np.random.RandomState(300 + int(d["n_envs"]) - 1)
    for _ in range(3):
        tp = A.draw_turn_params(rng)
    assert [float(tp[f]) for f in A.TURN_FIELDS] == list(d["turn_params"][-1])


def test_philox_turn_params_are_in_range_and_distinct():
    seen = set()
    for env in range(64):
        tp = A.philox_turn_params(1234, env, 0)
        assert 10 <= tp["main_corridor_length"] < 16 and 4 <= tp["turn_corridor_length"] < 12
        assert abs(tp["turn_corridor_angle"]) <= 3. / 8. * np.pi
        assert 0.5 <= tp["main_corridor_width"] < 1.5 and 0.5 <= tp["turn_corridor_width"] < 1.5
        assert 0 <= tp["rot_theta"] < 2 * np.pi
        seen.add(tp["rot_theta"])
    assert len(seen) == 64
    assert A.philox_turn_params(1234, 3, 0) != A.philox_turn_params(1234, 3, 1)
