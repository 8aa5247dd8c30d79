"""T-junction worlds: a column of width `column_width` meeting a beam of width `beam_width`, drawn as five
one-pixel walls, and the static paths between its three mouths.

Host-side mirror of the reference's envs/t_junction_env.py (`TJunction` :34-275, `Bearing` :14-30): same
constructor arguments, validation messages, corner lettering (O A B C D E F G), wall list, costmap sizing
(1 m margin, :141-171) and 2 x 150 way points per path (:173-275).  The resulting (costmap, path) pair feeds
`PlanEnv` / `VecPlanEnv` like any other world; stepping happens on the GPU.
"""
import numpy as np

from bc_gym_planning_env_b200.envs.base.maps import Wall
from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D

_MOUTHS = ("left", "right", "bottom")
WAY_POINTS_PER_LEG = 150


class Bearing(object):
    """Heading of the robot on the first (entering) and the second (leaving) leg of a path, by mouth."""
    starting = {"top": 3 * np.pi / 2, "bottom": np.pi / 2, "left": 0.0, "right": np.pi}
    ending = {"top": np.pi / 2, "bottom": 3 * np.pi / 2, "left": np.pi, "right": 0.0}


class TJunction(object):
    def __init__(self, start_noise_scale=0.0, window_height=10.0, window_width=10.0, column_width=1.5, beam_width=1.5):
        self.start_noise_scale = start_noise_scale
        self.window_height, self.window_width = window_height, window_width
        self.column_width, self.beam_width = column_width, beam_width
        self._validate()
        self.wall_corners = self.get_map_standard_coordinates()
        self.obstacles = self.get_map_walls()

    def _validate(self):
        checks = ((self.column_width > 0, "column_width_left must be a positive real number greater than 0"),
                  (self.beam_width > 0, "beam_width must be a positive real number greater than 0"),
                  (self.window_height > 0, "beam_width must be a positive real number greater than 0"),
                  (self.window_width > 0, "window_width must be a positive real number greater than 0"),
                  (self.column_width <= self.window_width, "column_width must be less than or equal to window_width"),
                  (self.beam_width <= self.window_height, "beam_width must be less than equal to window_height"))
        for ok, message in checks:
            if not ok:
                raise ValueError(message)

    def get_map_standard_coordinates(self):
        """Corners O A B C D E F G with O at the origin (reference :98-121)."""
        cw, top = self.column_width, self.window_height
        under_beam = top - self.beam_width
        overhang = (self.window_width - cw) / 2.0
        self._corner = dict(o=(0.0, 0.0), a=(0.0 + cw, 0.0), b=(0.0 + cw, 0.0 + under_beam),
                            c=(0.0 + cw + overhang, 0.0 + under_beam), d=(0.0 + cw + overhang, 0.0 + top),
                            e=(0.0 - overhang, 0.0 + top), f=(0.0 - overhang, 0.0 + under_beam), g=(0.0, 0.0 + under_beam))
        self._corner = {k: np.array(v) for k, v in self._corner.items()}
        return np.array([self._corner[k] for k in "oabcdefg"])

    def get_map_walls(self):
        c = self._corner
        return [Wall(from_pt=c[p], to_pt=c[q]) for p, q in ("ab", "bc", "de", "fg", "go")]

    def get_costmap(self, resolution=0.03):
        margin = 1.0
        lo, hi = self.wall_corners.min(axis=0), self.wall_corners.max(axis=0)
        costmap = CostMap2D.create_empty(world_size=(abs(hi[0] - lo[0]) + 2 * margin, abs(hi[1] - lo[1]) + 2 * margin),
                                         world_origin=(lo[0] - margin, lo[1] - margin), resolution=resolution)
        for wall in self.obstacles:
            wall.render(costmap)
        return costmap

    def _legs(self, noise):
        """(entering leg, leaving leg) per mouth: each a pair of end points, mouth -> junction and back."""
        c, half_col, half_beam = self._corner, self.column_width / 2.0, self.beam_width / 2
        centre_bottom = np.array([c['o'][0] + half_col, c['b'][1]])
        centre_left = np.array([c['g'][0], c['g'][1] + half_beam])
        centre_right = np.array([c['b'][0], c['b'][1] + half_beam])
        mouth = dict(bottom=np.array([c['o'][0] + half_col, c['o'][1]]),
                     left=np.array([c['f'][0], c['f'][1] + half_beam]),
                     right=np.array([c['c'][0], c['c'][1] + half_beam]))
        centre = dict(bottom=centre_bottom, left=centre_left, right=centre_right)
        entering = {k: (np.array([mouth[k][0] + noise[0], mouth[k][1] + noise[1]]), centre[k]) for k in _MOUTHS}
        leaving = {k: (centre[k], mouth[k]) for k in _MOUTHS}
        return entering, leaving

    def get_path(self, starting_position="bottom", ending_position="right"):
        if not isinstance(starting_position, str):
            raise TypeError("starting_position must be <type str>")
        if not isinstance(ending_position, str):
            raise TypeError("starting_position must be <type, str>")
        if starting_position == ending_position:
            raise ValueError("starting_position cannot equal the ending_position")
        if starting_position not in _MOUTHS:
            raise ValueError("starting_position can only be 'left', 'right', 'bottom'")
        if ending_position not in _MOUTHS:
            raise ValueError("ending_position can only be 'left', 'right', 'bottom'")
        noise = self.start_noise_scale * np.random.randn(2)          # the reference draws from the global RNG too
        entering, leaving = self._legs(noise)
        first = self._way_points(entering[starting_position], Bearing.starting[starting_position])
        second = self._way_points(leaving[ending_position], Bearing.ending[ending_position])
        return np.array(first + second)

    @staticmethod
    def _way_points(leg, bearing):
        start, end = leg
        t = np.linspace(0, 1, num=WAY_POINTS_PER_LEG)
        delta = end - start
        xs, ys = start[0] + t * delta[0], start[1] + t * delta[1]
        return [np.array([xs[k], ys[k], bearing]) for k in range(WAY_POINTS_PER_LEG)]
