"""Host-side map building at reset time: `Wall` obstacles rasterised onto a CostMap2D.

Follows the reference's contract (envs/base/maps.py:27-42 -> utilities/map_drawing_utils.py:140-156):
a wall is a cv2.line between the world_to_pixel'd end points, thickness max(1, int(width / res)),
painted with the wall's cost.  Map building is outside the per-step hot path.
"""
import attr
import numpy as np

try:
    import cv2
except ImportError:  # pragma: no cover
    cv2 = None

from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D


@attr.s
class Wall(object):
    from_pt = attr.ib(type=np.ndarray)
    to_pt = attr.ib(type=np.ndarray)
    width = attr.ib(type=float, default=0.05)
    cost = attr.ib(default=CostMap2D.LETHAL_OBSTACLE)

    def render(self, costmap):
        if cv2 is None:
            raise RuntimeError("rasterising walls needs OpenCV (cv2.line), like the reference")
        res = costmap.get_resolution()
        thickness = max(1, int(self.width / res))
        p0 = costmap.world_to_pixel(np.array(self.from_pt, dtype=np.float64))
        p1 = costmap.world_to_pixel(np.array(self.to_pt, dtype=np.float64))
        cv2.line(costmap.get_data(), (int(p0[0]), int(p0[1])), (int(p1[0]), int(p1[1])), color=self.cost,
                 thickness=thickness)
        return costmap
