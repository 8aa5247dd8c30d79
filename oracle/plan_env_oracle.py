"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the batched B200 `PlanEnv.step` path.

A from-the-spec NumPy fp64 restatement of what braincorp/bc-gym-planning-env computes in
`PlanEnv.step`.  Every function cites the upstream file:line whose
*behaviour* it follows (paths relative to /root/reference/bc_gym_planning_env).  Only tests/,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of bench.py may import
this module; the product package never does (it fails loudly without its CUDA extension).

Third-party arithmetic the reference delegates to and that is therefore *not* under /root/reference:
  * OpenCV (`opencv-python`, unpinned by the reference's setup.py:23; 4.13.0 in this image):
    `cv2.fillPoly` defines the footprint mask, `cv2.warpAffine`/`getRotationMatrix2D` define the
    egocentric crop.  `pixel_footprint` below calls cv2.fillPoly itself (exactly as the reference
    does), `ego_costmap_cv2` calls warpAffine; `ego_costmap` is the closed-form sampling rule of
    SURVEY.md A.9 that the CUDA kernel implements, and tests pin it against the cv2 call.
  * NumPy (2.3.5 here): np.round half-to-even, np.sinc, floor-mod `%`, global Mersenne-Twister
    normals.  Restated with the same NumPy primitives.

Parity pinning: this oracle is checked (a) against the reference's own known-answer tests
(tests/test_oracle_kats.py restates the vectors of utilities/test_coordinate_transformations.py,
test_path_tools.py, test_costmap_utils.py) and (b) against trajectories produced by the unmodified
reference imported in the build container (tests/golden/*.npz, generator oracle/gen_golden.py).
"""
import math

import numpy as np

try:  # the reference's own raster dependency; the oracle refuses to guess without it
    import cv2
except ImportError:  # pragma: no cover
    cv2 = None

LETHAL = 254  # utilities/costmap_2d.py:21
TWO_PI = 2 * np.pi

# ----------------------------------------------------------------------------------------------
# robot constant tables (robot_models/robot_dimensions_examples.py:51-85 diff-drive, :109-190 tricycle)
# ----------------------------------------------------------------------------------------------
TRICYCLE_FOOTPRINT = np.array([
    [1348.35, 0.], [1338.56, 139.75], [1306.71, 280.12], [1224.36, 338.62], [1093.81, 374.64],
    [-214.37, 374.64], [-313.62, 308.56], [-366.36, 117.44], [-374.01, -135.75], [-227.96, -459.13],
    [-156.72, -458.78], [759.8, -442.96], [849.69, -426.4], [1171.05, -353.74], [1303.15, -286.54],
    [1341.34, -118.37]]) / 1000.

DIFFDRIVE_FOOTPRINT = np.array([
    [644.5, 0], [634.86, 61], [571.935, 130.54], [553.38, 161], [360.36, 186], [250, 186], [250, 186],
    [100, 186], [100, 186], [0, 196], [-119.21, 190.5], [-173.4, 146], [-193, 0], [-173.4, -143],
    [-111.65, -246], [-71.57, -246], [100, -246], [100, -246], [250, -246], [250, -246],
    [413.085, -223], [491.5, -204.5], [553, -161], [634.86, -62]]) / 1000.

TRICYCLE = dict(
    kind="tricycle",
    footprint=TRICYCLE_FOOTPRINT,
    front_wheel_from_axis=0.964,
    max_front_wheel_angle=0.5 * 170 * np.pi / 180.,
    max_front_wheel_speed=60. * np.pi / 180.,
    max_linear_acceleration=1. / 2.5,
    max_angular_acceleration=1. / 2.,
    front_column_p_gain=0.16,
)
DIFFDRIVE = dict(kind="diffdrive", footprint=DIFFDRIVE_FOOTPRINT)
ROBOTS = {"industrial_tricycle_v1": TRICYCLE, "industrial_diffdrive_v1": DIFFDRIVE}

# PlanEnv's hard-wired odometry-noise alphas (envs/base/env.py:226-232)
DEFAULT_NOISE = (0.0, 0.0, 1.e-2, 1.e-2, 1.e-3, 1.e-3)


# ----------------------------------------------------------------------------------------------
# scalar helpers
# ----------------------------------------------------------------------------------------------
def normalize_angle(z):
    """utilities/coordinate_transformations.py:28-36 -- floor-mod wrap into [-pi, pi)."""
    return (np.asarray(z, dtype=np.float64) + np.pi) % TWO_PI - np.pi


def world_to_pixel(world, origin, resolution):
    """utilities/coordinate_transformations.py:185-205 -- multiply by the reciprocal, round half
    to even, cast to int."""
    anti = 1. / resolution
    return np.round((np.asarray(world, dtype=np.float64) - np.asarray(origin, dtype=np.float64)) * anti).astype(np.int64)


def delay_line(queue, element, delay):
    """envs/base/env.py:27-49 -- push; pop-front once the list is longer than `delay`, else peek."""
    queue.append(element)
    if len(queue) > delay:
        return queue.pop(0)
    return queue[0]


# ----------------------------------------------------------------------------------------------
# kinematics
# ----------------------------------------------------------------------------------------------
def pose_motion_step(x, y, th, v, w, dt):
    """robot_models/differential_drive.py:21-40 -- exact-arc unicycle integration via np.sinc."""
    h = 0.5 * w * dt
    f = v * dt * np.sinc(h / np.pi)
    nx = x + f * np.cos(th + h)
    ny = y + f * np.sin(th + h)
    nth = float(normalize_angle(th + w * dt))
    return float(nx), float(ny), nth


def noisy_pose_motion_step(x, y, th, v, w, dt, alphas, normal):
    """robot_models/differential_drive.py:43-74.  `normal(slot)` returns one N(0,1) draw; a draw is
    only consumed when its variance is > 0 (:48-52).  Slots: 0 = linear, 1 = angular, 2 = final
    rotation -- drawn in that order by the reference."""
    a1, a2, a3, a4, a5, a6 = alphas
    var = a1 * v ** 2 + a2 * w ** 2
    if var > 0:
        v = v + np.sqrt(var) * normal(0)
    var = a3 * v ** 2 + a4 * w ** 2
    if var > 0:
        w = w + np.sqrt(var) * normal(1)
    var = a5 * v ** 2 + a6 * w ** 2
    gamma = np.sqrt(var) * normal(2) if var > 0 else 0.
    nx, ny, nth = pose_motion_step(x, y, th, v, w, dt)
    nth = float(normalize_angle(nth + gamma * dt))
    return nx, ny, nth


def measured_velocity(x0, y0, th0, x1, y1, th1, dt):
    """utilities/path_tools.py:298-323 (`path_velocity` on the 2-row path the robots build,
    tricycle_model.py:520-529): chord speed with a heading-projected sign, wrapped yaw rate."""
    dx, dy = x1 - x0, y1 - y0
    sign = np.sign(np.cos(th0) * dx + np.sin(th0) * dy)
    if sign == 0.:
        sign = np.sign(np.sin(th0) * dy)
    ds = np.sqrt(dx * dx + dy * dy) * sign
    dth = th1 - th0
    if dth < -np.pi:
        dth += TWO_PI
    if dth > np.pi:
        dth -= TWO_PI
    return float(ds / dt), float(dth / dt)


def tricycle_step(state, cmd, dt, robot=TRICYCLE, alphas=None, normal=None):
    """robot_models/tricycle_model.py:478-538 (TricycleRobot.step) through :71-188.

    state = [x, y, th, v, w, steer_cmd, wheel]; cmd = (wheel_v, wheel_angle).  Returns new state."""
    x, y, th, v, w, _, wheel = [float(s) for s in state]
    u_v, u_phi = float(cmd[0]), float(cmd[1])
    # front column P-controller, rate- and range-limited (:127-154)
    max_delta = robot["max_front_wheel_speed"] * dt
    delta = robot["front_column_p_gain"] * (u_phi - wheel)
    delta = min(max(delta, -max_delta), max_delta)
    new_wheel = wheel + delta
    new_wheel = min(max(new_wheel, -robot["max_front_wheel_angle"]), robot["max_front_wheel_angle"])
    # acceleration-limited body velocities (:157-188)
    v_des = u_v * np.cos(new_wheel)
    w_des = u_v * np.sin(new_wheel) / robot["front_wheel_from_axis"]
    a_lin = (v_des - v) / dt
    a_ang = (w_des - w) / dt
    a_lin = min(max(a_lin, -2 * robot["max_linear_acceleration"]), robot["max_linear_acceleration"])
    a_ang = min(max(a_ang, -robot["max_angular_acceleration"]), robot["max_angular_acceleration"])
    v_new = v + a_lin * dt
    w_new = w + a_ang * dt
    v_new = v_new if not (0.0 > v_new) else 0.0
    if alphas is None:
        nx, ny, nth = pose_motion_step(x, y, th, v_new, w_new, dt)
    else:
        nx, ny, nth = noisy_pose_motion_step(x, y, th, v_new, w_new, dt, alphas, normal)
    mv, mw = measured_velocity(x, y, th, nx, ny, nth, dt)
    steer_cmd = wheel - u_phi  # :532
    return [nx, ny, nth, mv, mw, float(steer_cmd), float(new_wheel)]


def diffdrive_step(state, cmd, dt, alphas=None, normal=None):
    """robot_models/differential_drive.py:236-265 (DiffDriveRobot.step); cmd = (v, w) applied
    directly.  state = [x, y, th, v, w, 0, 0] (the two trailing slots are unused padding)."""
    x, y, th = float(state[0]), float(state[1]), float(state[2])
    v, w = float(cmd[0]), float(cmd[1])
    if alphas is None:
        nx, ny, nth = pose_motion_step(x, y, th, v, w, dt)
    else:
        nx, ny, nth = noisy_pose_motion_step(x, y, th, v, w, dt, alphas, normal)
    mv, mw = measured_velocity(x, y, th, nx, ny, nth, dt)
    return [nx, ny, nth, mv, mw, 0., 0.]


# ----------------------------------------------------------------------------------------------
# footprint raster + collision
# ----------------------------------------------------------------------------------------------
def rotated_pixel_vertices(angle, footprint, resolution):
    """utilities/path_tools.py:140-150 -- rotate footprint/res by `angle`
    (x' = fx c - fy s, y' = fx s + fy c), return (float verts, round-half-even int verts)."""
    c, s = np.cos(angle), np.sin(angle)
    f = np.asarray(footprint, dtype=np.float64) / resolution
    # Same BLAS route as the reference's np.dot(footprint/res, m[2,2,1]) (:145): each coordinate is
    # a length-2 ddot, which OpenBLAS evaluates here as fma(fy, -s, fx*c) / fma(fy, c, fx*s)
    # [probe: 4800/4800 bitwise].  The CUDA kernels hard-code that fused form.
    m = np.array([[c, -s], [s, c]], dtype=np.float64).reshape(2, 2, 1)
    rot = np.ascontiguousarray(np.dot(f, m)[:, :, 0])
    return rot, np.round(rot).astype(np.int32)


def pixel_footprint(angle, footprint, resolution):
    """utilities/path_tools.py:122-162 with fill=True: uint8 canvas (2*hy+1, 2*hx+1), polygon of
    rounded vertices filled by cv2.fillPoly around the canvas centre."""
    if cv2 is None:
        raise RuntimeError("oracle needs cv2 (the reference's raster dependency)")
    rot, iv = rotated_pixel_vertices(angle, footprint, resolution)
    corner = np.maximum(rot.max(axis=0), -rot.min(axis=0))
    half = np.ceil(corner).astype(np.int32)
    canvas = np.zeros((2 * half[1] + 1, 2 * half[0] + 1), dtype=np.uint8)
    cv2.fillPoly(canvas, [iv + half], (255, 255, 255))
    return canvas


def pose_collides(x, y, angle, footprint, costmap, origin, resolution):
    """envs/base/env.py:464-489 -- any in-map footprint pixel equal to LETHAL (254)."""
    mask = pixel_footprint(angle, footprint, resolution)
    rows, cols = np.where(mask)
    px, py = world_to_pixel(np.array([x, y]), origin, resolution)
    rr = py + rows - mask.shape[0] // 2
    cc = px + cols - mask.shape[1] // 2
    good = (rr >= 0) & (rr < costmap.shape[0]) & (cc >= 0) & (cc < costmap.shape[1])
    return bool(np.any(costmap[rr[good], cc[good]] == LETHAL))


def footprint_pixels_in_map(x, y, angle, footprint, costmap_shape, origin, resolution):
    """Number of footprint pixels that fall inside the map: the ALGORITHMIC byte count of one
    collision check on the reference's uint8 costmap."""
    mask = pixel_footprint(angle, footprint, resolution)
    rows, cols = np.where(mask)
    px, py = world_to_pixel(np.array([x, y]), origin, resolution)
    rr = py + rows - mask.shape[0] // 2
    cc = px + cols - mask.shape[1] // 2
    good = (rr >= 0) & (rr < costmap_shape[0]) & (cc >= 0) & (cc < costmap_shape[1])
    return int(good.sum())


# ----------------------------------------------------------------------------------------------
# path + reward
# ----------------------------------------------------------------------------------------------
def refine_path(path, delta):
    """utilities/path_tools.py:178-240 (angle_delta=None): a segment longer than delta gets
    int(d/delta)+2 evenly spaced points (end excluded), angle copied from the segment start."""
    path = np.asarray(path, dtype=np.float64)
    out = []
    seg = np.linalg.norm(np.diff(path[:, :2], axis=0), axis=1)
    for i, d in enumerate(seg):
        if d > delta:
            n = int(d / delta) + 2
            xs = np.linspace(path[i, 0], path[i + 1, 0], num=n)
            ys = np.linspace(path[i, 1], path[i + 1, 1], num=n)
            piece = np.stack([xs, ys, np.ones(n) * path[i, 2]], axis=1)
            out.append(piece[:-1])
        else:
            out.append(path[i][None])
    out.append(path[-1][None])
    return np.ascontiguousarray(np.vstack(out))


def reached_mask(pose, path, sp, ap):
    """utilities/path_tools.py:408-429 with the default parallel threshold -sp/9 (:423-424)."""
    dist = np.hypot(path[:, 0] - pose[0], path[:, 1] - pose[1])
    ang = np.abs(normalize_angle(pose[2] - path[:, 2]))
    par = np.cos(path[:, 2]) * (pose[0] - path[:, 0]) + np.sin(path[:, 2]) * (pose[1] - path[:, 1])
    return (dist < sp) & (ang < ap) & (par >= -sp / 9)


def last_reached(pose, path, sp, ap):
    """utilities/path_tools.py:432-448 -- max reached index or None."""
    idx = np.where(reached_mask(pose, path, sp, ap))[0]
    return int(idx[-1]) if len(idx) else None


def initial_reward_state(path, sp, ap):
    """envs/base/reward.py:261-288 -- (target_idx, min_dist) for a fresh episode."""
    last = last_reached(path[0], path, sp, ap)
    if last == len(path) - 1:
        raise ValueError("Goal pose too close to initial pose")
    target = last + 1
    d = float(np.hypot(path[target, 0] - path[0, 0], path[target, 1] - path[0, 1]))
    return target, d


def reward_step(pose, path, target_idx, min_dist, sp, ap, multiplier):
    """envs/base/reward.py:214-259.  Returns (reward, target_idx, min_dist)."""
    n = len(path)
    if target_idx > n - 1:
        return 0.0, target_idx, min_dist
    last = last_reached(pose, path, sp, ap)
    if last is not None and last >= target_idx:
        target_idx = last + 1
        if target_idx > n - 1:
            min_dist = 0.0
        else:
            g = path[target_idx]
            min_dist = float(np.hypot(g[0] - pose[0], g[1] - pose[1]))
        return 1.0, target_idx, min_dist
    g = path[target_idx]
    d = float(np.hypot(g[0] - pose[0], g[1] - pose[1]))
    if d < min_dist:
        r = (min_dist - d) * multiplier
        return float(r), target_idx, d
    return 0.0, target_idx, min_dist


def pure_pursuit_initial_state(path):
    """envs/base/reward.py:352-371 -- (target_idx, min_dist) of ContinuousRewardPurePursuitProvider."""
    return 1, float(np.hypot(path[-1, 0] - path[0, 0], path[-1, 1] - path[0, 1]))


def pure_pursuit_reward_step(pose, path, target_idx, min_dist, collided, radius=2.):
    """envs/base/reward.py:126-137 (update_goal: first way point from target_idx on that is more than
    `radius` away, else the last one) and :331-350 (reward = -0.05 + progress towards the LAST path point,
    -100 while collided).  Returns (reward, target_idx, min_dist)."""
    new_target = len(path) - 1
    for i in range(target_idx, len(path)):
        if np.linalg.norm(path[i, :2] - pose[:2]) > radius:
            new_target = i
            break
    d = float(np.hypot(path[-1, 0] - pose[0], path[-1, 1] - pose[1]))
    reward = -0.05
    reward += min_dist - d
    if collided:
        reward -= 100
    return float(reward), new_target, d


def pure_pursuit_done(pose, path):
    """envs/base/reward.py:139-149 -- within 1 m of the last path point."""
    return bool(np.hypot(path[-1, 0] - pose[0], path[-1, 1] - pose[1]) < 1.0)


# ----------------------------------------------------------------------------------------------
# egocentric observation
# ----------------------------------------------------------------------------------------------
EGO_X_BOUNDS = (-0.5, 3.)   # envs/egocentric.py:113
EGO_Y_BOUNDS = (-2., 2.)    # envs/egocentric.py:114


def ego_crop_size(resolution=0.03):
    """envs/egocentric.py:115-119 sizes the gym space with a hard-wired 0.03; the crop itself is
    sized with the costmap resolution (utilities/costmap_utils.py:49).  Returns (W, H) pixels."""
    size = np.array([EGO_X_BOUNDS[1] - EGO_X_BOUNDS[0], EGO_Y_BOUNDS[1] - EGO_Y_BOUNDS[0]])
    wh = world_to_pixel(size, np.zeros(2), resolution)
    return int(wh[0]), int(wh[1])


def ego_affine_f32(pose, origin, resolution):
    """utilities/costmap_utils.py:42-65: cv2.getRotationMatrix2D about the robot pixel composed
    (in float32, as the reference does) with the shift that puts (-0.5, -2.0) at the crop origin."""
    px, py = world_to_pixel(np.array(pose[:2], dtype=np.float64), origin, resolution)
    deg = 180 * pose[2] / np.pi
    rad = deg * (np.pi / 180.)          # cv2.getRotationMatrix2D converts back with CV_PI/180
    a, b = math.cos(rad), math.sin(rad)
    cx, cy = float(px), float(py)
    rot = np.array([[a, b, (1 - a) * cx - b * cy], [-b, a, b * cx + (1 - a) * cy]], dtype=np.float64)
    shift_w = np.array([EGO_X_BOUNDS[0], EGO_Y_BOUNDS[0]]) - (np.asarray(origin, dtype=np.float64) - np.asarray(pose[:2], dtype=np.float64))
    ds = world_to_pixel(shift_w, np.zeros(2), resolution)
    r32 = rot.astype(np.float32)
    m = r32.copy()
    m[0, 2] = np.float32(r32[0, 2] - np.float32(ds[0]))
    m[1, 2] = np.float32(r32[1, 2] - np.float32(ds[1]))
    return m


def ego_costmap(costmap, pose, origin, resolution):
    """Closed-form restatement of extract_egocentric_costmap (utilities/costmap_utils.py:25-75)
    for uint8 maps, i.e. of cv2.warpAffine(INTER_NEAREST, borderValue=0) on the float32 matrix:
    fp64 inverse, 10-bit fixed-point source coordinates."""
    w, h = ego_crop_size(resolution)
    m = ego_affine_f32(pose, origin, resolution).astype(np.float64)
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
    xs = (adx[None, :] + bdx[:, None]) >> 10
    ys = (ady[None, :] + bdy[:, None]) >> 10
    hh, ww = costmap.shape
    ok = (xs >= 0) & (xs < ww) & (ys >= 0) & (ys < hh)
    out = np.zeros((h, w), dtype=np.uint8)
    out[ok] = costmap[ys[ok], xs[ok]]
    return out


def ego_costmap_cv2(costmap, pose, origin, resolution):
    """The literal cv2 route of utilities/costmap_utils.py:66-72, used to pin `ego_costmap`."""
    w, h = ego_crop_size(resolution)
    m = ego_affine_f32(pose, origin, resolution)
    return cv2.warpAffine(costmap, m, (w, h), flags=cv2.INTER_NEAREST, borderValue=0)


def ego_path(path, pose):
    """utilities/coordinate_transformations.py:341-362 -> :57-84 (inverse_transform) -> :310-328 (project_poses): the
    path in the robot frame, R(-th)(p - t), angles wrapped.  Like the reference, the points go through ONE
    np.dot(3x3 homogeneous matrix, [x; y; 1] columns) (:322-324): BLAS evaluates that 3-term dot product with fused
    multiply-adds, so it must be the same call on the same shapes to give the same last bits (a way point dead ahead
    of the robot has y ~ 1e-17 by cancellation)."""
    path = np.asarray(path, dtype=np.float64)
    c, s = np.cos(pose[2]), np.sin(pose[2])
    tx = -pose[0] * c - pose[1] * s
    ty = pose[0] * s - pose[1] * c
    tt = float(normalize_angle(-pose[2]))
    ct, st = np.cos(tt), np.sin(tt)
    h = np.identity(3)
    h[:2, :2] = np.array([[ct, -st], [st, ct]])
    h[:2, 2] = (tx, ty)
    ph = np.hstack((path[:, :2], np.ones((path.shape[0], 1))))
    out = np.dot(h, ph.T).T
    out[:, 2] = normalize_angle(path[:, 2] + tt)
    return out


def goal_n_state(remaining_path, pose, robot_state, resolution=0.03):
    """envs/egocentric.py:141-160: float32 (9,) = [clip(ego goal xy / crop world size, +-1),
    ego goal angle, x, y, th, v, w, wheel]; zeros when no path is left."""
    if len(remaining_path) == 0:
        return np.zeros(9, dtype=np.float32)
    w, h = ego_crop_size(resolution)
    ox, oy = EGO_X_BOUNDS[0], EGO_Y_BOUNDS[0]
    world = np.array([(ox + resolution * w) - ox, (oy + resolution * h) - oy])  # costmap_2d.py:106-121
    g = ego_path(np.asarray(remaining_path, dtype=np.float64), pose)[0]      # the whole remaining path, like the reference
    ng = np.clip(g[:2] / world, (-1., -1.), (1., 1.))
    x, y, th, v, w_, _, wheel = robot_state
    return np.hstack([ng, g[2], [x, y, th, v, w_, wheel]]).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# counter-based noise (the product's replacement for the reference's global np.random)
# ----------------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_M32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11), pure-Python integers; KAT-pinned in tests."""
    c0, c1, c2, c3 = [int(c) & _M32 for c in counter]
    k0, k1 = int(key[0]) & _M32, int(key[1]) & _M32
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
        k0, k1 = (k0 + _PHILOX_W0) & _M32, (k1 + _PHILOX_W1) & _M32
    return c0, c1, c2, c3


def philox_normals(seed, env_id, step, block):
    """Two N(0,1) draws for (seed; env_id, step, block): 53-bit uniforms from the four words,
    Box-Muller in fp64.  The device kernel uses the same construction (csrc/bcg_kernels.cu)."""
    r = philox4x32_10((env_id & _M32, step & _M32, block & _M32, (step >> 32) & _M32),
                      (seed & _M32, (seed >> 32) & _M32))
    u1 = (((r[0] << 21) | (r[1] >> 11)) + 1) * (2.0 ** -53)   # (0, 1]
    u2 = ((r[2] << 21) | (r[3] >> 11)) * (2.0 ** -53)         # [0, 1)
    rad = math.sqrt(-2.0 * math.log(u1))
    ang = TWO_PI * u2
    return rad * math.cos(ang), rad * math.sin(ang)


def philox_normal_source(seed, env_id, step):
    """normal(slot) for noisy_pose_motion_step: slot 1 (angular) and slot 2 (final rotation) are
    the two halves of block 0 -- the only draws PlanEnv's default alphas consume -- and slot 0
    (linear) is the first half of block 1."""
    cache = {}

    def normal(slot):
        block = 1 if slot == 0 else 0
        if block not in cache:
            cache[block] = philox_normals(seed, env_id, step, block)
        return cache[block][0] if slot in (0, 1) else cache[block][1]

    return normal


# ----------------------------------------------------------------------------------------------
# the env state machine
# ----------------------------------------------------------------------------------------------
class OraclePlanEnv(object):
    """One environment, stepped exactly like envs/base/env.py:334-461.

    costmap: uint8 [H, W]; origin (2,), resolution; path: fp64 [N, 3] (coarse; refined here when
    `refine`); delays = (control, pose, state); alphas None = noise off; `normal_source(step)`
    returns the normal(slot) callable for that step (defaults to np.random like the reference).
    """

    def __init__(self, costmap, origin, resolution, path, robot="industrial_tricycle_v1", dt=0.05,
                 sp=1.0, ap=np.pi / 2, multiplier=0.0, timeout=1200, delays=(0, 0, 0), alphas=None,
                 normal_source=None, refine=True, path_delta=0.05, initial_wheel_angle=0.0,
                 reward_provider="continuous_reward", footprint_scale=1.0):
        self.costmap = np.ascontiguousarray(costmap)
        self.origin = np.asarray(origin, dtype=np.float64)
        self.resolution = float(resolution)
        self.robot = dict(ROBOTS[robot])
        if footprint_scale != 1.0:          # robot_models/tricycle_model.py:371: dimensions.footprint() * footprint_scale
            self.robot["footprint"] = np.asarray(self.robot["footprint"]) * footprint_scale
        self.dt, self.sp, self.ap, self.multiplier = dt, sp, ap, multiplier
        self.timeout = int(timeout)
        self.delays = tuple(int(d) for d in delays)
        self.alphas = alphas
        self.normal_source = normal_source
        self.pure_pursuit = reward_provider == "continuous_reward_pure_pursuit"
        path = np.asarray(path, dtype=np.float64)
        self.path = refine_path(path, path_delta) if refine else np.ascontiguousarray(path)
        self.reset()

    # envs/base/env.py:179-214 make_initial_state (+ :293-303 reset)
    def reset(self):
        p0 = self.path[0]
        # NB: TricycleRobotState() starts with wheel_angle 0 (tricycle_model.py:234-244): PlanEnv
        # never applies EnvParams.initial_wheel_angle (env.py:192-197).
        self.robot_state = [float(p0[0]), float(p0[1]), float(p0[2]), 0., 0., 0., 0.]
        self.delayed_robot_state = list(self.robot_state)
        self.pose = np.array(p0, dtype=np.float64)
        if self.pure_pursuit:
            self.target_idx, self.min_dist = pure_pursuit_initial_state(self.path)
        else:
            self.target_idx, self.min_dist = initial_reward_state(self.path, self.sp, self.ap)
        self.time, self.iter, self.collided = 0.0, 0, False
        self.control_queue, self.pose_queue, self.state_queue = [], [], []
        return self.observation()

    def observation(self):
        """envs/base/env.py:421-433 -- (delayed pose, remaining path, delayed robot state, time)."""
        path = self.path[:self.target_idx + 1] if self.pure_pursuit else self.path[self.target_idx:]   # reward.py:120-124 / :59-64
        return dict(pose=np.array(self.pose), path=path, robot_state=list(self.delayed_robot_state), time=self.time)

    def done(self):
        """envs/base/env.py:400-419 + envs/base/reward.py:66-69 (:139-149 for pure pursuit)."""
        goal = pure_pursuit_done(self.pose, self.path) if self.pure_pursuit else self.target_idx > len(self.path) - 1
        return bool(goal or self.iter >= self.timeout or self.collided)

    def step(self, cmd):
        """envs/base/env.py:334-398, 442-461.  Returns (observation, reward, done, hit_this_step)."""
        dc, dp, ds = self.delays
        cmd = delay_line(self.control_queue, np.asarray(cmd, dtype=np.float64), dc)
        old = self.robot_state
        if self.alphas is None:
            normal = None
        elif self.normal_source is not None:
            normal = self.normal_source(self.iter)
        else:
            normal = lambda slot: np.random.normal(0, 1)  # noqa: E731 -- reference's global RNG
        if self.robot["kind"] == "tricycle":
            new = tricycle_step(old, cmd, self.dt, self.robot, self.alphas, normal)
        else:
            new = diffdrive_step(old, cmd, self.dt, self.alphas, normal)
        hit = pose_collides(new[0], new[1], new[2], self.robot["footprint"], self.costmap, self.origin,
                            self.resolution)
        if hit:  # env.py:458-459 + tricycle_model.py:471-476: pose restored, v = w = 0, wheel kept
            new = [old[0], old[1], old[2], 0., 0., new[5], new[6]]
        self.robot_state = new
        self.pose = np.array(delay_line(self.pose_queue, np.array(new[:3]), dp))
        self.time = self.time + self.dt
        self.iter += 1
        self.delayed_robot_state = list(delay_line(self.state_queue, list(new), ds))
        self.collided = self.collided or hit
        if self.pure_pursuit:
            r, self.target_idx, self.min_dist = pure_pursuit_reward_step(self.pose, self.path, self.target_idx,
                                                                         self.min_dist, self.collided)
        else:
            r, self.target_idx, self.min_dist = reward_step(self.pose, self.path, self.target_idx, self.min_dist,
                                                            self.sp, self.ap, self.multiplier)
        return self.observation(), float(r), self.done(), hit

    # envs/base/env.py:278-291 -- deep snapshot; set_state loads the *delayed* robot state into the robot
    def get_state(self):
        return dict(robot_state=list(self.robot_state), delayed_robot_state=list(self.delayed_robot_state),
                    pose=np.array(self.pose), target_idx=self.target_idx, min_dist=self.min_dist,
                    time=self.time, iter=self.iter, collided=self.collided,
                    control_queue=[np.array(c) for c in self.control_queue],
                    pose_queue=[np.array(p) for p in self.pose_queue],
                    state_queue=[list(s) for s in self.state_queue])

    def set_state(self, s):
        self.delayed_robot_state = list(s["delayed_robot_state"])
        self.robot_state = list(s["delayed_robot_state"])   # env.py:284
        self.pose = np.array(s["pose"])
        self.target_idx, self.min_dist = int(s["target_idx"]), float(s["min_dist"])
        self.time, self.iter, self.collided = float(s["time"]), int(s["iter"]), bool(s["collided"])
        self.control_queue = [np.array(c) for c in s["control_queue"]]
        self.pose_queue = [np.array(p) for p in s["pose_queue"]]
        self.state_queue = [list(q) for q in s["state_queue"]]
