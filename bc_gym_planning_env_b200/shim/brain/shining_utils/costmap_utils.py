"""world_to_pixel_impl, get_pixel_footprint_impl and is_footprint_colliding_impl backed by libbcg_b200 / the exact footprint table."""
import ctypes as C

import numpy as np
import torch

from bc_gym_planning_env_b200 import _native as nat
from bc_gym_planning_env_b200.footprint_lut import FootprintLut

_LUTS = {}


def world_to_pixel_impl(world_coords, origin, resolution):
    """replaces utilities/coordinate_transformations.py:185-205 (contract: test_coordinate_transformations.py:1544-1649):
    round half to even of (xy - origin) * (1 / resolution), same shape, C int."""
    if not isinstance(world_coords, np.ndarray) or not isinstance(origin, np.ndarray):
        raise TypeError("world_to_pixel works with numpy arrays only")
    if world_coords.shape[world_coords.ndim - 1] != 2 or len(origin) != 2:
        raise ValueError("expected (..., 2) coordinates and a 2-element origin")
    nat.require_cuda()
    xy = torch.from_numpy(np.ascontiguousarray(world_coords, dtype=np.float64).reshape(-1, 2)).cuda()
    out = torch.empty(xy.shape, dtype=torch.int32, device=xy.device)
    if xy.numel():
        nat.check(nat.lib().bcg_world_to_pixel(nat.ptr(xy), xy.shape[0], float(origin[0]), float(origin[1]), float(resolution),
                                               nat.ptr(out), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy().astype(np.intc).reshape(world_coords.shape)


def get_pixel_footprint_impl(angle, robot_footprint, map_resolution, fill=True):
    """replaces utilities/path_tools.py:122-162: the mask of the angle's bin in the exact footprint table (one table
    per footprint x resolution, built once with the same cv2.fillPoly call the reference makes)."""
    footprint = np.ascontiguousarray(robot_footprint, dtype=np.float64)
    if not fill:
        # contour only (not on the env's path; the table holds filled masks): drawn directly, like the reference does
        import cv2
        c, s = np.cos(angle), np.sin(angle)
        rot = np.dot(footprint / map_resolution, np.array([[c, -s], [s, c]], dtype=np.float64).reshape(2, 2, 1))[:, :, 0]
        half = np.ceil(np.maximum(rot.max(axis=0), -rot.min(axis=0))).astype(np.int32)
        out = np.zeros((2 * half[1] + 1, 2 * half[0] + 1), dtype=np.uint8)
        cv2.polylines(out, [np.round(rot).astype(np.int32) + half], 1, (255, 255, 255))
        return out
    key = (footprint.tobytes(), float(map_resolution))
    if key not in _LUTS:
        _LUTS[key] = FootprintLut(footprint, float(map_resolution))
    return _LUTS[key].canvas(float(angle))


def is_footprint_colliding_impl(image_slice, blit_mask, lethal_value):
    """replaces the native hook of utilities/costmap_utils.py:106-136 (contract: test_costmap_utils.py:251-325): does any
    costmap cell under the blitted footprint equal `lethal_value`?  image_slice: the costmap view get_blit_mask cut out,
    blit_mask: the footprint's boolean mask over it."""
    image_slice, blit_mask = np.asarray(image_slice), np.asarray(blit_mask)
    if image_slice.shape[:2] != blit_mask.shape[:2] or image_slice.dtype != np.uint8:
        raise TypeError("is_footprint_colliding_impl takes a uint8 image slice and a mask of the same shape")
    nat.require_cuda()
    vals = torch.from_numpy(np.ascontiguousarray(image_slice).reshape(-1)).cuda()
    mask = torch.from_numpy(np.ascontiguousarray(blit_mask, dtype=np.uint8).reshape(-1)).cuda()
    flag = torch.zeros(1, dtype=torch.int32, device=vals.device)
    nat.check(nat.lib().bcg_masked_any_equal(nat.ptr(vals), nat.ptr(mask), vals.numel(), int(lethal_value), nat.ptr(flag),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return bool(flag.item())
