"""normalize_angle_impl (replaces utilities/coordinate_transformations.py:28-36) and inverse_transform_2d_impl (:39-84)
backed by libbcg_b200."""
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


def inverse_transform_2d_impl(transform):
    """replaces the native hook of utilities/coordinate_transformations.py:39-84 (contract:
    test_coordinate_transformations.py:105-178): (x, y, angle) or array(N, 3) of them -> the inverse transform(s)."""
    t = np.asarray(transform, dtype=np.float64)
    if t.ndim not in (1, 2) or t.shape[-1] != 3:
        raise TypeError("inverse_transform takes an (x, y, angle) transform or an array(N, 3) of them")
    nat.require_cuda()
    dev = torch.from_numpy(np.ascontiguousarray(t).reshape(-1, 3)).cuda()
    out = torch.empty_like(dev)
    nat.check(nat.lib().bcg_inverse_transform(nat.ptr(dev), dev.shape[0], nat.ptr(out),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy().reshape(t.shape)
