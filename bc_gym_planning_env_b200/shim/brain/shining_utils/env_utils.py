"""native_project_poses backed by bcg_project_poses (replaces the native hook of
utilities/coordinate_transformations.py:289-328; contract: test_coordinate_transformations.py:1298-1325)."""
import ctypes as C

import numpy as np
import torch

from bc_gym_planning_env_b200 import _native as nat


def native_project_poses(transform, poses, out):
    """out[i] = poses[i] rotated by transform's angle and moved by its translation, angle wrapped into [-pi, pi).
    transform: contiguous float64 (3,), poses: float64 (N, 3), out: float64 (N, 3), written in place."""
    if not (isinstance(transform, np.ndarray) and isinstance(poses, np.ndarray) and isinstance(out, np.ndarray)):
        raise TypeError("native_project_poses works with numpy arrays only")
    if transform.shape != (3,) or poses.ndim != 2 or poses.shape[1] != 3 or out.shape != poses.shape:
        raise TypeError("native_project_poses(transform (3,), poses (N, 3), out (N, 3))")
    nat.require_cuda()
    t = (C.c_double * 3)(*[float(v) for v in transform])
    dev = torch.from_numpy(np.ascontiguousarray(poses, dtype=np.float64)).cuda()
    res = torch.empty_like(dev)
    nat.check(nat.lib().bcg_project_poses(t, nat.ptr(dev), dev.shape[0], nat.ptr(res),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    out[...] = res.cpu().numpy()
