"""normalize_angle_impl backed by bcg_normalize_angle (replaces utilities/coordinate_transformations.py:28-36)."""
import ctypes as C

import numpy as np
import torch

from bc_gym_planning_env_b200 import _native as nat


def normalize_angle_impl(z):
    """Wrap angles into [-pi, pi).  Accepts a float or a contiguous 1-d float array; like the reference's native
    hook it raises TypeError on 2-d or non-contiguous input (utilities/test_coordinate_transformations.py:90-100)."""
    nat.require_cuda()
    scalar = np.isscalar(z) or (isinstance(z, np.ndarray) and z.ndim == 0)
    if not scalar:
        if not isinstance(z, np.ndarray):
            z = np.asarray(z, dtype=np.float64)
        if z.ndim != 1 or not z.flags['C_CONTIGUOUS']:
            raise TypeError("normalize_angle_impl takes a float or a contiguous 1-d array")
    arr = np.atleast_1d(np.asarray(z, dtype=np.float64))
    dev = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    out = torch.empty_like(dev)
    if dev.numel():
        nat.check(nat.lib().bcg_normalize_angle(nat.ptr(dev), dev.numel(), nat.ptr(out),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    res = out.cpu().numpy()
    return float(res[0]) if scalar else res
