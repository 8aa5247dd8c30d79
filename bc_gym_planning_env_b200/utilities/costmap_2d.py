"""CostMap2D: the reference's costmap data contract (utilities/costmap_2d.py:13-174) --
uint8 [H, W] cells + fp64 (x, y) origin of cell [0, 0] + resolution in metres per cell."""
import numpy as np


class CostMap2D(object):
    FREE_SPACE = 0
    LETHAL_OBSTACLE = 254
    NO_INFORMATION = 255

    def __init__(self, data, resolution, origin):
        origin = np.array(origin, copy=True)
        assert origin.dtype == np.float64 and origin.shape == (2,)
        origin.flags.writeable = False   # costmap_2d.py:33-37: the origin is frozen
        self._data = data
        self._resolution = resolution
        self._origin = origin

    @staticmethod
    def create_empty(world_size, resolution, world_origin=(0., 0.), dtype=np.uint8):
        size = np.round(np.asarray(world_size, dtype=np.float64) * (1. / resolution)).astype(int)
        return CostMap2D(np.zeros(size[::-1], dtype=dtype), resolution, np.asarray(world_origin, dtype=np.float64))

    def get_data(self):
        return self._data

    def get_resolution(self):
        return self._resolution

    def get_origin(self):
        return self._origin

    def in_bounds(self, map_x, map_y):
        return 0 <= map_x < self._data.shape[1] and 0 <= map_y < self._data.shape[0]

    def world_bounds(self):
        x, y = self._origin
        return (x, x + self._resolution * self._data.shape[1], y, y + self._resolution * self._data.shape[0])

    def world_size(self):
        xmin, xmax, ymin, ymax = self.world_bounds()
        return np.array([xmax - xmin, ymax - ymin])

    def world_center(self):
        xmin, xmax, ymin, ymax = self.world_bounds()
        return np.array([xmax + xmin, ymax + ymin]) * 0.5

    def world_to_pixel(self, world_coords):
        return np.round((np.asarray(world_coords, dtype=np.float64) - self._origin) * (1. / self._resolution)).astype(int)

    def pixel_to_world(self, pixel_coords):
        return np.asarray(pixel_coords) * self._resolution + self._origin

    def copy(self):
        return CostMap2D(self._data.copy(), self._resolution, self._origin.copy())

    def __eq__(self, other):
        return (isinstance(other, CostMap2D) and self._resolution == other.get_resolution()
                and bool((self._origin == other.get_origin()).all())
                and self._data.shape == other.get_data().shape and bool((self._data == other.get_data()).all()))

    def __ne__(self, other):
        return not self.__eq__(other)

    # wire format of costmap_2d.py:150-174
    def get_state(self):
        return dict(version=1, data=self._data, resolution=self._resolution, origin=self._origin)

    @classmethod
    def from_state(cls, state):
        assert state['version'] == 1
        return cls(data=state['data'], resolution=state['resolution'], origin=np.array(state['origin'], dtype=np.float64))
