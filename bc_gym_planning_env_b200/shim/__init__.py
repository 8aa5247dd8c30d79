"""`brain.shining_utils` shim: put this directory on sys.path (`bc_gym_planning_env_b200.shim.install()`) BEFORE
importing the reference and its optional native hooks resolve to libbcg_b200.so (SURVEY 8b seam ii):

    utilities/coordinate_transformations.py:17-20     normalize_angle  <- transform_utils.normalize_angle_impl
    utilities/coordinate_transformations.py:169-171   world_to_pixel   <- costmap_utils.world_to_pixel_impl
    utilities/path_tools.py:101-103                   get_pixel_footprint <- costmap_utils.get_pixel_footprint_impl
    utilities/costmap_utils.py:106-107                is_robot_colliding's mask test <- costmap_utils.is_footprint_colliding_impl
    utilities/coordinate_transformations.py:39-41     inverse_transform <- transform_utils.inverse_transform_2d_impl
    utilities/coordinate_transformations.py:289-290   project_poses     <- env_utils.native_project_poses

Each call is a host<->device round trip: the shim is for conformance (the reference's own tests exercise this
library), not for speed -- speed comes from the batch API.
"""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
