"""The mini-env generation oracle (oracle/mini_oracle.py) against cv2.line with far-away end points and against the
worlds the unmodified reference sampled (tests/golden/mini_worlds.npz)."""
import numpy as np
import pytest

from oracle import mini_oracle as M
from oracle import plan_env_oracle as O
from tests.common import load

cv2 = pytest.importorskip("cv2")

FIELDS = (("h", 1), ("w", 1), ("start", 3), ("end", 3), ("a", 2), ("o", 2), ("b", 2))


def unpack(row):
    out, k = {}, 0
    for name, n in FIELDS:
        out[name] = float(row[k]) if n == 1 else np.array(row[k:k + n])
        k += n
    return out


def test_clipped_lines_match_cv2():
    rng = np.random.RandomState(0)
    for t in range(4000):
        w, h = (183, 183) if t % 2 else (rng.randint(5, 300), rng.randint(5, 300))
        if t % 3 == 0:
            p = rng.randint(-1500, 1700, size=4)
        else:
            p = np.r_[rng.randint(0, w), rng.randint(0, h), rng.randint(-1500, 1700, size=2)]
        x1, y1, x2, y2 = [int(v) for v in p]
        img = np.zeros((h, w), np.uint8)
        cv2.line(img, (x1, y1), (x2, y2), color=254, thickness=1)
        mine = np.zeros_like(img)
        px = M.clipped_line_pixels(w, h, x1, y1, x2, y2)
        mine[px[:, 1], px[:, 0]] = 254
        assert np.array_equal(img, mine), (w, h, x1, y1, x2, y2)


def test_sampler_and_worlds_match_the_reference():
    d = load("mini_worlds")
    res = float(d["resolution"])
    for s in range(int(d["n_envs"])):
        mp, attempts = M.sample_mini_params(M.rng_source(np.random.RandomState(400 + s)))
        want = unpack(d["params_%d" % s])
        for name, _ in FIELDS:
            assert np.array_equal(np.asarray(mp[name]), np.asarray(want[name])), (s, name)
        coarse, costmap, origin = M.mini_world(mp, res)
        assert np.array_equal(costmap, d["costmap_%d" % s]) and np.array_equal(origin, d["origin_%d" % s])
        path = O.refine_path(coarse, 0.05)
        assert np.array_equal(path, d["path_%d" % s])
        target, min_dist = O.initial_reward_state(path, 0.2, np.pi / 8.)
        assert target == int(d["target_idx_%d" % s]) and min_dist == float(d["min_dist_%d" % s])
        assert 1 <= attempts < 1000


def test_philox_source_is_a_counter_stream():
    a, b = M.philox_source(3, 5, 0), M.philox_source(3, 5, 0)
    xs = [a() for _ in range(9)]
    assert xs == [b() for _ in range(9)] and len(set(xs)) == 9 and all(0 <= x < 1 for x in xs)
    assert M.philox_source(3, 5, 1)() != xs[0] and M.philox_source(3, 6, 0)() != xs[0]
