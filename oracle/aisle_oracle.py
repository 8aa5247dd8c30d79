"""TEST INFRASTRUCTURE ONLY -- CPU oracle for device-side AisleTurnEnv generation.

NumPy fp64 restatement of what the reference does when `RandomAisleTurnEnv.reset` draws a new turn
(paths relative to /root/reference/bc_gym_planning_env):

  envs/synth_turn_env.py:317-332   _draw_random_turn_params      -> draw_turn_params (RandomState) /
                                                                      philox_turn_params (the device's RNG)
  envs/synth_turn_env.py:41-98     corner / way points           -> aisle_world
  envs/synth_turn_env.py:110-192   path_and_costmap_from_config  -> aisle_world
  envs/base/maps.py:27-42, utilities/map_drawing_utils.py:140-156  Wall.render = cv2.line -> line_pixels
  utilities/costmap_2d.py create_empty (world_to_pixel of the size) -> aisle_world

Third-party arithmetic: `cv2.line` (OpenCV 4.13.0 here) with thickness 1, 8-connected.  `line_pixels` is
the closed form of its LineIterator that the CUDA kernel implements; tests/test_oracle_aisle.py pins it
against the cv2 call itself and pins `aisle_world` against worlds built by the unmodified reference
(tests/golden/aisle_worlds.npz, made by oracle/gen_golden.py).  Only tests/ may import this module.
"""
import numpy as np

from oracle.plan_env_oracle import LETHAL, TWO_PI, philox4x32_10, world_to_pixel

TURN_FIELDS = ("main_corridor_length", "turn_corridor_length", "turn_corridor_angle", "main_corridor_width",
               "turn_corridor_width", "margin", "rot_theta", "flip_arnd_oy", "flip_arnd_ox")


def draw_turn_params(rng):
    """envs/synth_turn_env.py:317-332 with a numpy RandomState: same distributions, same draw order."""
    return dict(
        main_corridor_length=rng.uniform(10, 16),
        turn_corridor_length=rng.uniform(4, 12),
        turn_corridor_angle=rng.uniform(-3. / 8. * np.pi, 3. / 8. * np.pi),
        main_corridor_width=rng.uniform(0.5, 1.5),
        turn_corridor_width=rng.uniform(0.5, 1.5),
        margin=1.0,
        flip_arnd_oy=bool(rng.rand() < 0.5),
        flip_arnd_ox=bool(rng.rand() < 0.5),
        rot_theta=rng.uniform(0, 2 * np.pi))


def philox_uniform(seed, env, draw, k):
    """The device's uniform in [0, 1): 53 bits of one half of the Philox4x32-10 block with counter
    (env, draw lo, 'AISE' + k // 2, draw hi) and key = seed; draws 2j and 2j + 1 share a block."""
    ctr = (env & 0xffffffff, draw & 0xffffffff, (0x41495345 + (k >> 1)) & 0xffffffff, (draw >> 32) & 0xffffffff)
    r = philox4x32_10(ctr, (seed & 0xffffffff, (seed >> 32) & 0xffffffff))
    hi, lo = (r[2], r[3]) if (k & 1) else (r[0], r[1])
    return float((int(hi) << 21) | (int(lo) >> 11)) * 2.0 ** -53


def philox_turn_params(seed, env, draw):
    """_draw_random_turn_params with the device's counter-based RNG in place of MT19937."""
    u = [philox_uniform(seed, env, draw, k) for k in range(8)]
    lim = 3. / 8. * np.pi
    return dict(
        main_corridor_length=10.0 + 6.0 * u[0],
        turn_corridor_length=4.0 + 8.0 * u[1],
        turn_corridor_angle=-lim + (2 * lim) * u[2],
        main_corridor_width=0.5 + u[3],
        turn_corridor_width=0.5 + u[4],
        margin=1.0,
        flip_arnd_oy=bool(u[5] < 0.5),
        flip_arnd_ox=bool(u[6] < 0.5),
        rot_theta=TWO_PI * u[7])


def line_pixels(x0, y0, x1, y1):
    """Pixels (x, y) of cv2.line((x0, y0), (x1, y1), thickness=1, LINE_8) on an image that contains both
    ends: OpenCV's LineIterator run left to right.  The major axis advances every step; after step j the
    minor axis has moved ceil((2 d j - D) / (2 D)) times (0 while negative), D / d = major / minor extent."""
    if x1 < x0:
        x0, y0, x1, y1 = x1, y1, x0, y0
    dx, dy = x1 - x0, y1 - y0
    sy = 1 if dy >= 0 else -1
    ady = abs(dy)
    steep = ady > dx
    D, dm = (ady, dx) if steep else (dx, ady)
    j = np.arange(D + 1, dtype=np.int64)
    t = 2 * dm * j - D
    mv = np.where(t > 0, (t + 2 * D - 1) // max(2 * D, 1), 0)
    if steep:
        return np.stack([x0 + mv, y0 + j * sy], axis=1)
    return np.stack([x0 + j, y0 + mv * sy], axis=1)


def aisle_world(tp, resolution=0.03):
    """path_and_costmap_from_config (envs/synth_turn_env.py:110-192).
    Returns (coarse path fp64 [4, 3], costmap uint8 [H, W], origin fp64 [2])."""
    h, far = tp["main_corridor_length"] / 2, tp["turn_corridor_length"] / 2
    alpha, d, z, margin = tp["turn_corridor_angle"], tp["main_corridor_width"], tp["turn_corridor_width"], tp["margin"]
    ta, ca = np.tan(alpha), np.cos(alpha)
    lower, upper = -z / ca, z / ca
    corners = np.array([(-d, -h), (0, -h), (d, -h), (d, d * ta + lower), (far, far * ta + lower), (far, far * ta),
                        (d, d * ta + upper), (far, far * ta + upper), (-d, h), (d, h)])          # a .. j (:41-77)
    way = [(0, -h, np.pi / 2), (0, d * ta + lower, np.pi / 2), (d, d * ta, alpha), (far * ca, far * ca * ta, alpha)]
    c, s = np.cos(tp["rot_theta"]), np.sin(tp["rot_theta"])
    flip = np.array([[-1. if tp["flip_arnd_oy"] else 1., 0.], [0., -1. if tp["flip_arnd_ox"] else 1.]])
    transform = np.dot(np.array(((c, -s), (s, c))), flip)
    moved = np.array([np.dot(transform, pt) for pt in corners])
    path = []
    for x, y, t in way:
        nx, ny = np.dot(transform, np.array([x, y]))
        if tp["flip_arnd_ox"]:
            t = -t
        if tp["flip_arnd_oy"]:
            t = np.pi - t
        path.append([nx, ny, np.mod(t + tp["rot_theta"], 2 * np.pi)])
    min_x, max_x, min_y, max_y = moved[:, 0].min(), moved[:, 0].max(), moved[:, 1].min(), moved[:, 1].max()
    size = np.array([abs(max_x - min_x) + 2 * margin, abs(max_y - min_y) + 2 * margin])
    origin = np.array([min_x - margin, min_y - margin])
    w, hgt = world_to_pixel(size, np.zeros(2), resolution)          # CostMap2D.create_empty
    costmap = np.zeros((int(hgt), int(w)), dtype=np.uint8)
    assert max(1, int(0.05 / resolution)) == 1, "walls thicker than one pixel are not restated here"
    A, C_, D_, G, H, I, J = moved[0], moved[2], moved[3], moved[6], moved[7], moved[8], moved[9]
    E = moved[4]
    for p0, p1 in ((A, I), (C_, D_), (D_, E), (J, G), (G, H)):
        q0, q1 = world_to_pixel(p0, origin, resolution), world_to_pixel(p1, origin, resolution)
        px = line_pixels(int(q0[0]), int(q0[1]), int(q1[0]), int(q1[1]))
        keep = (px[:, 0] >= 0) & (px[:, 0] < costmap.shape[1]) & (px[:, 1] >= 0) & (px[:, 1] < costmap.shape[0])
        costmap[px[keep, 1], px[keep, 0]] = LETHAL
    return np.array(path), costmap, origin
