"""CPU: the claim ego_sparse_kernel rests on, checked in NumPy against the oracle's gather (= cv2.warpAffine).

The kernel renders the egocentric crop as a scatter: every occupied source cell (X, Y) is mapped forward with the
float32 matrix cv2 is given (in 16.16 fixed point, relative to the window origin), and only the <= 4 crop pixels around its image are tested with the exact fixed-point
inverse rule `X == (adx[u] + bx[v]) >> 10 and Y == (ady[u] + by[v]) >> 10`.  That is bit-identical to
the gather iff every pixel that samples (X, Y) is among those four candidates.  Here: random maps, random poses (on
the map, at its border, outside it), every cost value -- scatter == gather."""
import numpy as np
import pytest

from oracle import plan_env_oracle as O


def _scatter(costmap, pose, origin, resolution):
    w, h = O.ego_crop_size(resolution)
    m32 = O.ego_affine_f32(pose, origin, resolution)
    m = m32.astype(np.float64)
    det = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1. / det if det != 0 else 0.
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    a12, a21 = -m[0, 1] * d, -m[1, 0] * d
    b1 = -a11 * m[0, 2] - a12 * m[1, 2]
    b2 = -a21 * m[0, 2] - a22 * m[1, 2]
    u = np.arange(w, dtype=np.float64)
    v = np.arange(h, dtype=np.float64)
    adx = np.rint(a11 * u * 1024).astype(np.int64)
    ady = np.rint(a21 * u * 1024).astype(np.int64)
    bdx = np.rint((a12 * v + b1) * 1024).astype(np.int64) + 512
    bdy = np.rint((a22 * v + b2) * 1024).astype(np.int64) + 512
    out = np.zeros((h, w), dtype=np.uint8)
    ys, xs = np.nonzero(costmap)
    if len(xs) == 0:
        return out, 0
    # the kernel's candidates: the forward image in 16.16 fixed point relative to the window origin (the tile-aligned
    # corner of the rotated crop's bounding box, write_ego_tile_record), floored, and the next pixel, per axis
    qx = np.array([b1, a11 * (w - 1) + b1, a11 * (w - 1) + a12 * (h - 1) + b1, a12 * (h - 1) + b1])
    qy = np.array([b2, a21 * (w - 1) + b2, a21 * (w - 1) + a22 * (h - 1) + b2, a22 * (h - 1) + b2])
    x0 = int(np.floor(qx.min() - 0.51)) & ~15
    y0 = int(np.floor(qy.min() - 0.51)) & ~7
    fx = [int(np.rint(m[0, 0] * 65536)), int(np.rint(m[0, 1] * 65536)),
          int(np.rint(np.clip(m[0, 0] * x0 + m[0, 1] * y0 + m[0, 2], -16000, 16000) * 65536))]
    fy = [int(np.rint(m[1, 0] * 65536)), int(np.rint(m[1, 1] * 65536)),
          int(np.rint(np.clip(m[1, 0] * x0 + m[1, 1] * y0 + m[1, 2], -16000, 16000) * 65536))]
    xr, yr = xs.astype(np.int64) - x0, ys.astype(np.int64) - y0
    fu = (fx[0] * xr + fx[1] * yr + fx[2]) >> 16
    fv = (fy[0] * xr + fy[1] * yr + fy[2]) >> 16
    inside = (fu >= -1) & (fu < w) & (fv >= -1) & (fv < h)
    xs, ys, fu, fv = xs[inside], ys[inside], fu[inside], fv[inside]
    vals = costmap[ys, xs]
    tested = 0
    for du in (0, 1):
        for dv in (0, 1):
            cu, cv = fu + du, fv + dv
            ok = (cu >= 0) & (cu < w) & (cv >= 0) & (cv < h)     # one step outside the crop: the kernel reads a sentinel that never hits
            cu, cv = np.clip(cu, 0, w - 1), np.clip(cv, 0, h - 1)
            hit = ok & (((adx[cu] + bdx[cv]) >> 10) == xs) & (((ady[cu] + bdy[cv]) >> 10) == ys)
            out[cv[hit], cu[hit]] = vals[hit]
            tested += len(cu)
    return out, tested


def _random_map(rng, kind):
    h, w = int(rng.randint(40, 400)), int(rng.randint(40, 400))
    m = np.zeros((h, w), dtype=np.uint8)
    if kind == "walls":                       # thin lines, like the aisle worlds
        import cv2
        for _ in range(rng.randint(2, 7)):
            p0 = (int(rng.randint(0, w)), int(rng.randint(0, h)))
            p1 = (int(rng.randint(0, w)), int(rng.randint(0, h)))
            cv2.line(m, p0, p1, 254, 1)
    elif kind == "noise":                     # every cost value, 5 % of the cells
        mask = rng.rand(h, w) < 0.05
        m[mask] = rng.randint(1, 256, size=int(mask.sum())).astype(np.uint8)
    else:                                     # filled blocks (what the dense kernel gets on the device)
        for _ in range(rng.randint(1, 5)):
            y0, x0 = rng.randint(0, h - 10), rng.randint(0, w - 10)
            m[y0:y0 + rng.randint(5, 60), x0:x0 + rng.randint(5, 60)] = rng.choice([100, 253, 254, 255])
    return m


@pytest.mark.parametrize("kind", ["walls", "noise", "blocks"])
@pytest.mark.parametrize("resolution", [0.03, 0.05])
def test_scatter_of_occupied_cells_equals_the_gather(kind, resolution):
    rng = np.random.RandomState({"walls": 1, "noise": 2, "blocks": 3}[kind] + int(resolution * 1000))
    hits = 0
    for trial in range(40):
        m = _random_map(rng, kind)
        h, w = m.shape
        origin = rng.uniform(-5, 5, size=2)
        # poses over the map, around its border and well outside it
        px = origin[0] + rng.uniform(-0.3, 1.3) * w * resolution
        py = origin[1] + rng.uniform(-0.3, 1.3) * h * resolution
        for th in (rng.uniform(-np.pi, np.pi), rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi / 4, -np.pi])):
            pose = np.array([px, py, th])
            want = O.ego_costmap(m, pose, origin, resolution)
            got, _ = _scatter(m, pose, origin, resolution)
            assert np.array_equal(got, want), (kind, trial, pose)
            hits += int((want != 0).sum())
    assert hits > 1000                        # the crops did see the maps
