"""Corridor-with-three-boxes environments.

The reference downloads a pickled real-world corridor costmap, the human-driven path and 1000
box-randomised variants from S3 (envs/rw_corridors/tdwa_test_environments.py:17-62).  That data cannot be
fetched here, so `get_random_maps_squeeze_between_obstacle_in_corridor_on_path` keeps the reference's name
and return shape -- (original costmap, path, tuple of randomised costmaps) -- but builds a SYNTHETIC
STAND-IN: a 20 m x 8 m corridor at 0.03 m/cell with unknown (255) outside, lethal
(254) walls with an inscribed-cost (253) lining, free (0) inside, three 0.6 m lethal boxes pasted with
+-0.5 m jitter per variant, and a gently weaving centre-line path.  Results obtained on it must be
labelled "synthetic stand-in".
"""
import numpy as np

from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D

_RES = 0.03
_WORLD = (20.0, 8.0)          # metres
_BOX = 0.6
_BOX_CENTRES = ((6.0, 4.9), (10.0, 3.2), (14.0, 4.8))


def _px(v):
    return int(round(v / _RES))


def _base_corridor():
    w, h = _px(_WORLD[0]), _px(_WORLD[1])
    data = np.full((h, w), CostMap2D.NO_INFORMATION, dtype=np.uint8)
    y0, y1 = _px(2.0), _px(6.0)                      # a 4 m wide corridor along x
    data[y0:y1, _px(0.5):w - _px(0.5)] = CostMap2D.FREE_SPACE
    lining = _px(0.15)
    for ya, yb in ((y0, y0 + lining), (y1 - lining, y1)):
        data[ya:yb, _px(0.5):w - _px(0.5)] = 253
    data[y0 - 2:y0, :] = CostMap2D.LETHAL_OBSTACLE
    data[y1:y1 + 2, :] = CostMap2D.LETHAL_OBSTACLE
    data[y0:y1, _px(0.5) - 2:_px(0.5)] = CostMap2D.LETHAL_OBSTACLE
    data[y0:y1, w - _px(0.5):w - _px(0.5) + 2] = CostMap2D.LETHAL_OBSTACLE
    return data


def _paste_boxes(data, centres):
    out = data.copy()
    half = _px(_BOX / 2)
    for cx, cy in centres:
        px, py = _px(cx), _px(cy)
        out[py - half - 3:py + half + 3, px - half - 3:px + half + 3] = 253
        out[py - half:py + half, px - half:px + half] = CostMap2D.LETHAL_OBSTACLE
    return out


def _centre_line_path():
    xs = np.arange(1.8, 18.2, 0.25)
    ys = 4.0 - 0.55 * np.sin((xs - 2.0) * (2 * np.pi / 8.0))     # weave between the boxes
    th = np.arctan2(np.gradient(ys), np.gradient(xs))
    return np.stack([xs, ys, th], axis=1)


def get_random_maps_squeeze_between_obstacle_in_corridor_on_path(n_variants=1000, seed=0):
    """:return: (original costmap with the 3 boxes, reference path array(n, 3), tuple of n_variants costmaps with
    the boxes re-pasted around their original places) -- synthetic stand-in, see the module docstring."""
    origin = np.array([0.0, 0.0])
    base = _base_corridor()
    original = CostMap2D(_paste_boxes(base, _BOX_CENTRES), _RES, origin)
    rng = np.random.RandomState(seed)
    variants = []
    for _ in range(n_variants):
        centres = [(cx + rng.uniform(-0.5, 0.5), cy + rng.uniform(-0.5, 0.5)) for cx, cy in _BOX_CENTRES]
        variants.append(CostMap2D(_paste_boxes(base, centres), _RES, origin))
    return original, _centre_line_path(), tuple(variants)
