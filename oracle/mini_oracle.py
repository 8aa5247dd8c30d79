"""TEST INFRASTRUCTURE ONLY -- CPU oracle for device-side RandomMiniEnv generation.

NumPy fp64 restatement of what the reference does when `RandomMiniEnv` (re)builds its env (paths relative to
/root/reference/bc_gym_planning_env):

  envs/mini_env.py:269-320   _sample_mini_env_params_no_final_check   -> draw_mini_params
  envs/mini_env.py:113-240   _sample_pose, _sample_pose_circ, _pick_pts_square_method / _circle_method
  envs/mini_env.py:323-361   _sample_mini_env_params (collision / too-close rejection) -> sample_mini_params
  envs/mini_env.py:364-389   prepare_map_and_path                       -> mini_world
  OpenCV cv::clipLine + LineIterator (cv2.line of walls that leave the map) -> clip_line + aisle_oracle.line_pixels

The sampler takes its randomness from a `uniform01()` callable, so the same code runs on a numpy RandomState (to be
pinned against the reference: tests/test_oracle_mini.py, tests/golden/mini_worlds.npz) and on the device's Philox
stream (`philox_source`).  Only tests/ may import this module.
"""
import numpy as np

from oracle.aisle_oracle import line_pixels
from oracle.plan_env_oracle import (LETHAL, TRICYCLE_FOOTPRINT, TWO_PI, normalize_angle, philox4x32_10, pose_collides,
                                    world_to_pixel)

GEN_DEFAULTS = dict(inner_h=3., inner_w=3., mid_margin=0.25, out_margin=1., min_obstacle_angle=np.pi / 8.,
                    max_obstacle_angle=np.pi, lim_euc_dist=1000., lim_ang_dist=np.pi, angular_pose_noise_scale=np.pi / 2.0,
                    goal_spat_dist=0.2, goal_ang_dist=np.pi / 8.)       # RandomMiniEnv's own defaults (mini_env.py:424-430)


class SpaceSeemsEmpty(Exception):
    pass


def rng_source(rng):
    """uniform01 backed by a numpy RandomState: rng.uniform(a, b) == a + (b - a) * rng.random_sample()."""
    return lambda: rng.random_sample()


def philox_source(seed, env, draw):
    """uniform01 of the device: k-th call = 53 bits of one half of the Philox4x32-10 block with counter
    (env, draw lo, 'MINI' + k // 2, draw hi), key = seed."""
    state = {"k": 0}

    def uniform01():
        k = state["k"]
        state["k"] += 1
        ctr = (env & 0xffffffff, draw & 0xffffffff, (0x4d494e49 + (k >> 1)) & 0xffffffff, (draw >> 32) & 0xffffffff)
        r = philox4x32_10(ctr, (seed & 0xffffffff, (seed >> 32) & 0xffffffff))
        hi, lo = (r[2], r[3]) if (k & 1) else (r[0], r[1])
        return float((int(hi) << 21) | (int(lo) >> 11)) * 2.0 ** -53
    return uniform01


def _uniform(u01, a, b):
    return a + (b - a) * u01()


def _not_inside_obstacle(x, y, o, start_angle, angle):
    phi = float(normalize_angle(np.arctan2(y - o[1], x - o[0])))
    if start_angle <= phi <= start_angle + angle:
        return False
    if start_angle <= phi + TWO_PI <= start_angle + angle:
        return False
    return True


def _sample_pose(u01, g, constraints):
    """mini_env.py:113-138"""
    for _ in range(1000):
        x = _uniform(u01, -g["inner_w"] / 2 - g["mid_margin"], g["inner_w"] / 2 + g["mid_margin"])
        y = _uniform(u01, -g["inner_h"] / 2 - g["mid_margin"], g["inner_h"] / 2 + g["mid_margin"])
        theta = float(normalize_angle(_uniform(u01, 0, TWO_PI)))
        if all(f(x, y, theta) for f in constraints):
            return x, y, theta
    raise ValueError("Something went wrong, the sampling space looks empty.")


def draw_mini_params(u01, g):
    """_sample_mini_env_params_no_final_check (mini_env.py:269-320): dict with h, w, start (3), end (3), a, o, b."""
    o = (_uniform(u01, -g["inner_w"] / 2, g["inner_w"] / 2), _uniform(u01, -g["inner_h"] / 2, g["inner_h"] / 2))
    start_angle = _uniform(u01, 0, TWO_PI)
    angle = _uniform(u01, g["min_obstacle_angle"], g["max_obstacle_angle"])
    r = 3 * (g["inner_h"] + g["inner_w"] + g["mid_margin"] + g["out_margin"])
    a = (r * np.cos(start_angle) + o[0], r * np.sin(start_angle) + o[1])
    b = (r * np.cos(start_angle + angle) + o[0], r * np.sin(start_angle + angle) + o[1])
    h = g["inner_h"] + 2 * g["mid_margin"] + 2 * g["out_margin"]
    w = g["inner_w"] + 2 * g["mid_margin"] + 2 * g["out_margin"]
    outside = lambda x, y, th=None: _not_inside_obstacle(x, y, o, start_angle, angle)   # noqa: E731
    if u01() < 0.7:
        # circle method (mini_env.py:141-178, :243-266): two opposite points of a circle, heading along the chord
        for _ in range(1000):
            rc = min((g["inner_w"] + g["inner_h"]) / 4. + g["mid_margin"], g["lim_euc_dist"])
            phi = _uniform(u01, 0, TWO_PI)
            x, y = rc * np.cos(phi), rc * np.sin(phi)
            _uniform(u01, 0, TWO_PI)                                     # a heading the reference draws and discards
            theta = float(normalize_angle(np.arctan2(-y - y, -x - x)))
            if outside(x, y) and outside(-x, -y):
                start, end = (x, y, theta), (-x, -y, theta)
                break
        else:
            raise SpaceSeemsEmpty()
    else:
        # square method (:181-240)
        sx, sy, st = _sample_pose(u01, g, [outside])
        ex, ey, _ = _sample_pose(u01, g, [outside,
                                          lambda x, y, th: np.mod(st - th, TWO_PI) < g["lim_ang_dist"],
                                          lambda x, y, th: np.sqrt((sx - x) ** 2 + (sy - y) ** 2) < g["lim_euc_dist"]])
        theta = float(normalize_angle(np.arctan2(ey - sy, ex - sx)))
        start, end = (sx, sy, theta), (ex, ey, theta)
    n1 = _uniform(u01, -g["angular_pose_noise_scale"] / 2.0, g["angular_pose_noise_scale"] / 2.0)
    start = (start[0], start[1], float(normalize_angle(start[2] + n1)))
    n2 = _uniform(u01, -g["angular_pose_noise_scale"] / 2.0, g["angular_pose_noise_scale"] / 2.0)
    end = (end[0], end[1], float(normalize_angle(end[2] + n2)))
    return dict(h=h, w=w, start=np.array(start), end=np.array(end), a=np.array(a), o=np.array(o), b=np.array(b))


def clip_line(w, h, x1, y1, x2, y2):
    """cv::clipLine(Size(w, h), pt1, pt2) (OpenCV imgproc drawing.cpp; 4.13.0 here): the end points of the part of
    the segment inside the image, computed in integers with truncation, or None when nothing is inside."""
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return None
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    if (c1 | c2) != 0:
        return None
    return x1, y1, x2, y2


def clipped_line_pixels(w, h, x1, y1, x2, y2):
    """Pixels of cv2.line((x1, y1), (x2, y2), thickness=1) on a w x h image, end points anywhere."""
    c = clip_line(w, h, int(x1), int(y1), int(x2), int(y2))
    return np.zeros((0, 2), dtype=np.int64) if c is None else line_pixels(*c)


def mini_world(mp, resolution=0.03):
    """prepare_map_and_path (mini_env.py:364-389): (coarse path [2, 3], costmap uint8, origin)."""
    size = np.array([mp["h"], mp["w"]], dtype=np.float64)               # world_size=(h, w): x extent h, y extent w
    origin = np.array([-mp["h"] / 2., -mp["w"] / 2.])
    wpx, hpx = world_to_pixel(size, np.zeros(2), resolution)
    costmap = np.zeros((int(hpx), int(wpx)), dtype=np.uint8)
    assert max(1, int(0.05 / resolution)) == 1
    for far in (mp["a"], mp["b"]):
        q0, q1 = world_to_pixel(mp["o"], origin, resolution), world_to_pixel(far, origin, resolution)
        px = clipped_line_pixels(int(wpx), int(hpx), q0[0], q0[1], q1[0], q1[1])
        costmap[px[:, 1], px[:, 0]] = LETHAL
    return np.array([mp["start"], mp["end"]]), costmap, origin


def sample_mini_params(u01, g=None, resolution=0.03, footprint=TRICYCLE_FOOTPRINT):
    """_sample_mini_env_params (mini_env.py:323-361): draw until neither end pose collides and the two are not
    within the goal tolerances of each other.  Returns (params, attempts)."""
    g = dict(GEN_DEFAULTS) if g is None else g
    for attempt in range(1000):
        try:
            mp = draw_mini_params(u01, g)
        except SpaceSeemsEmpty:
            continue
        path, costmap, origin = mini_world(mp, resolution)
        hit = [pose_collides(p[0], p[1], p[2], footprint, costmap, origin, resolution) for p in path]
        cart = float(np.hypot(path[0, 0] - path[1, 0], path[0, 1] - path[1, 1]))
        ang = float(np.abs(normalize_angle(path[0, 2] - path[1, 2])))
        too_close = cart < g["goal_spat_dist"] and ang < g["goal_ang_dist"]
        if not (hit[0] or hit[1]) and not too_close:
            return mp, attempt + 1
    raise ValueError("Something went wrong, the sampling space looks empty.")
