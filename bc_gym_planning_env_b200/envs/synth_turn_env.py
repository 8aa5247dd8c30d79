"""Aisle-turn environments: geometry -> (path, costmap) on the host, stepping on the GPU.

Mirrors the reference's envs/synth_turn_env.py API: TurnParams, AisleTurnEnvParams,
path_and_costmap_from_config (:110-192), AisleTurnEnv (:195-216), RandomAisleTurnEnv (:219-332).
`draw_random_turn_params` consumes the RandomState exactly like `_draw_random_turn_params`
(:317-332), so `RandomAisleTurnEnv(seed=s)` builds the same map as the reference's.
"""
import attr
import numpy as np

from bc_gym_planning_env_b200.envs.base.env import PlanEnv
from bc_gym_planning_env_b200.envs.base.maps import Wall
from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D


@attr.s
class TurnParams(object):
    main_corridor_length = attr.ib(default=8, type=float)
    turn_corridor_length = attr.ib(default=5, type=float)
    turn_corridor_angle = attr.ib(default=2 * np.pi / 8, type=float)
    main_corridor_width = attr.ib(default=1.0, type=float)
    turn_corridor_width = attr.ib(default=1.0, type=float)
    margin = attr.ib(default=1.0, type=float)
    flip_arnd_oy = attr.ib(default=False, type=bool)
    flip_arnd_ox = attr.ib(default=False, type=bool)
    rot_theta = attr.ib(default=0, type=float)


@attr.s
class AisleTurnEnvParams(object):
    env_params = attr.ib(factory=EnvParams)
    turn_params = attr.ib(factory=TurnParams)


def draw_random_turn_params(rng):
    """Same draw order as the reference's _draw_random_turn_params (:317-332)."""
    return TurnParams(
        main_corridor_length=rng.uniform(10, 16),
        turn_corridor_length=rng.uniform(4, 12),
        turn_corridor_angle=rng.uniform(-3. / 8. * np.pi, 3. / 8. * np.pi),
        main_corridor_width=rng.uniform(0.5, 1.5),
        turn_corridor_width=rng.uniform(0.5, 1.5),
        flip_arnd_oy=bool(rng.rand() < 0.5),
        flip_arnd_ox=bool(rng.rand() < 0.5),
        rot_theta=rng.uniform(0, 2 * np.pi))


def path_and_costmap_from_config(params):
    """Aisle geometry of the reference (:41-192): a main corridor of half-width d along +y from
    -h to +h, a side corridor of half-width z leaving at angle alpha, both as 1-pixel walls; four
    oriented way points through the turn.  Optional mirror flips, then a rotation by rot_theta.
    Returns (coarse path array(4, 3), CostMap2D)."""
    tp = params.turn_params
    h, far = tp.main_corridor_length / 2, tp.turn_corridor_length / 2
    alpha, d, z = tp.turn_corridor_angle, tp.main_corridor_width, tp.turn_corridor_width
    ta, ca = np.tan(alpha), np.cos(alpha)
    lower, upper = -z / ca, z / ca
    corners = dict(                      # reference's lettering (:41-77)
        a=(-d, -h), b=(0, -h), c=(d, -h),
        d=(d, d * ta + lower), e=(far, far * ta + lower), f=(far, far * ta),
        g=(d, d * ta + upper), h=(far, far * ta + upper), i=(-d, h), j=(d, h))
    waypoints = [(0, -h, np.pi / 2), (0, d * ta + lower, np.pi / 2), (d, d * ta, alpha),
                 (far * ca, far * ca * ta, alpha)]

    c, s = np.cos(tp.rot_theta), np.sin(tp.rot_theta)
    flip = np.array([[-1. if tp.flip_arnd_oy else 1., 0.], [0., -1. if tp.flip_arnd_ox else 1.]])
    transform = np.dot(np.array(((c, -s), (s, c))), flip)
    moved = {k: np.dot(transform, np.array(v)) for k, v in corners.items()}

    path = []
    for x, y, t in waypoints:
        nx, ny = np.dot(transform, np.array([x, y]))
        if tp.flip_arnd_ox:
            t = -t
        if tp.flip_arnd_oy:
            t = np.pi - t
        path.append(np.array([nx, ny, np.mod(t + tp.rot_theta, 2 * np.pi)]))

    pts = np.array([moved[k] for k in 'abcdefghij'])
    min_x, max_x, min_y, max_y = pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max()
    world_size = abs(max_x - min_x) + 2 * tp.margin, abs(max_y - min_y) + 2 * tp.margin
    world_origin = min_x - tp.margin, min_y - tp.margin
    costmap = CostMap2D.create_empty(world_size=world_size, resolution=params.env_params.resolution,
                                     world_origin=world_origin)
    for p, q in ('ai', 'cd', 'de', 'jg', 'gh'):
        Wall(from_pt=moved[p], to_pt=moved[q]).render(costmap)
    return np.array(path), costmap


class AisleTurnEnv(PlanEnv):
    """Turn into an aisle whose geometry `config` (AisleTurnEnvParams) fixes."""

    def __init__(self, config, **kw):
        self._config = config
        path, costmap = path_and_costmap_from_config(config)
        super(AisleTurnEnv, self).__init__(costmap, path, config.env_params, **kw)


class RandomAisleTurnEnv(object):
    """AisleTurnEnv with the turn drawn at random on construction and (by default) on every reset."""

    def __init__(self, params=None, draw_new_turn_on_reset=True, seed=None, rng=None, **kw):
        self._rng = np.random.RandomState() if rng is None else rng
        self.seed(seed)
        self._draw_new_turn_on_reset = draw_new_turn_on_reset
        self._env_params = EnvParams() if params is None else params
        self._kw = kw
        self.config = AisleTurnEnvParams(turn_params=draw_random_turn_params(self._rng), env_params=self._env_params)
        self._env = AisleTurnEnv(self.config, **kw)
        self.action_space = self._env.action_space

    def seed(self, seed=None):
        if seed is not None:
            self._rng.seed(seed)

    def step(self, action):
        return self._env.step(action)

    def reset(self):
        if self._draw_new_turn_on_reset:
            self.config = AisleTurnEnvParams(turn_params=draw_random_turn_params(self._rng), env_params=self._env_params)
            self._env = AisleTurnEnv(self.config, **self._kw)
        return self._env.reset()

    def render(self, mode='human'):
        return self._env.render(mode)

    def close(self):
        self._env.close()

    def get_state(self):
        return self._env.get_state()

    def set_state(self, state):
        self._env.set_state(state)


class ColoredCostmapRandomAisleTurnEnv(RandomAisleTurnEnv):
    """RandomAisleTurnEnv whose observation is the whole costmap, uint8 (H, W, 1)
    (reference envs/synth_turn_env.py:335-377)."""

    def step(self, action):
        rich_obs, reward, done, info = super(ColoredCostmapRandomAisleTurnEnv, self).step(action)
        return np.expand_dims(rich_obs.costmap.get_data(), -1), reward, done, info

    def reset(self):
        rich_obs = super(ColoredCostmapRandomAisleTurnEnv, self).reset()
        return np.expand_dims(rich_obs.costmap.get_data(), -1)


class ColoredEgoCostmapRandomAisleTurnEnv(RandomAisleTurnEnv):
    """RandomAisleTurnEnv whose observation is {'environment': egocentric crop uint8 (133, 133, 1) about the
    true robot pose, 'goal': float64 (5, 1)} (reference envs/synth_turn_env.py:380-451).  Crop and goal vector
    come from the CUDA egocentric kernel (VecPlanEnv.observe_colored_ego); the goal vector is computed in fp64
    and stored in fp32 on the device."""

    def _extract_egocentric_observation(self, rich_observation):
        from collections import OrderedDict
        image, goal = self._env._vec.observe_colored_ego()
        return OrderedDict((('environment', image[0].cpu().numpy()),
                            ('goal', goal[0].cpu().numpy().astype(np.float64))))

    def step(self, action):
        rich_obs, reward, done, info = super(ColoredEgoCostmapRandomAisleTurnEnv, self).step(action)
        return self._extract_egocentric_observation(rich_obs), reward, done, info

    def reset(self):
        return self._extract_egocentric_observation(super(ColoredEgoCostmapRandomAisleTurnEnv, self).reset())


def random_aisle_pool(n, seed, env_params=None):
    """n random aisle turns for a batch: ([CostMap2D], [coarse path]) drawn like
    RandomAisleTurnEnv(seed=seed + i) would draw its first turn."""
    env_params = EnvParams() if env_params is None else env_params
    costmaps, paths = [], []
    for i in range(n):
        rng = np.random.RandomState(seed + i)
        cfg = AisleTurnEnvParams(turn_params=draw_random_turn_params(rng), env_params=env_params)
        path, costmap = path_and_costmap_from_config(cfg)
        costmaps.append(costmap)
        paths.append(path)
    return costmaps, paths
