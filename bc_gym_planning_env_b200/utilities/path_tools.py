"""Host-side path helpers that run at reset time (not per step)."""
import numpy as np


def refine_path(data, delta):
    """Densify a path to <= delta spacing the way make_initial_state needs it (reference
    utilities/path_tools.py:178-240, angle_delta=None): a segment longer than delta is replaced by
    int(d / delta) + 2 evenly spaced points whose angle is the segment start's; the last original
    point closes the path."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim != 2 or data.shape[1] not in (2, 3):
        raise Exception("This function takes n x (x, y) or n x (x, y, angle) arrays")
    lengths = np.linalg.norm(np.diff(data[:, :2], axis=0), axis=1)
    pieces = []
    for i, d in enumerate(lengths):
        if d > delta:
            n = int(d / delta) + 2
            cols = [np.linspace(data[i, j], data[i + 1, j], num=n) for j in range(2)]
            if data.shape[1] == 3:
                cols.append(np.full(n, data[i, 2]))
            pieces.append(np.stack(cols, axis=1)[:-1])
        else:
            pieces.append(data[i:i + 1])
    pieces.append(data[-1:])
    return np.ascontiguousarray(np.concatenate(pieces, axis=0))
