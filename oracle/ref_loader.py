"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* upstream reference.

Imports braincorp/bc-gym-planning-env read-only from /root/reference (present in the
build container, absent on the GPU box) so that

  * the oracle restatement in this directory can be validated against the real thing, and
  * golden fixtures under tests/golden/ can be (re)generated (oracle/gen_golden.py).

Nothing in the product package may import this module.  The shim does not edit the
reference: it only restores the numpy aliases the reference was written against
(`np.int`, `np.float`, `np.bool`; used at utilities/coordinate_transformations.py:205 and
utilities/path_tools.py:174,209) and coerces the `center` argument of
cv2.getRotationMatrix2D to a float tuple (utilities/costmap_utils.py:44 passes np.int64,
which cv2 4.13 rejects).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("BCG_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bc_gym_planning_env"))


_loaded = False


def load_reference():
    """Make `import bc_gym_planning_env` resolve to the upstream tree.  Idempotent."""
    global _loaded
    if _loaded:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import numpy as np
    import cv2

    for name, typ in (("int", int), ("float", float), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, typ)

    if not getattr(cv2.getRotationMatrix2D, "_bcg_shim", False):
        _orig = cv2.getRotationMatrix2D

        def _get_rotation_matrix_2d(center, angle, scale):
            return _orig((float(center[0]), float(center[1])), float(angle), float(scale))

        _get_rotation_matrix_2d._bcg_shim = True
        cv2.getRotationMatrix2D = _get_rotation_matrix_2d

    sys.dont_write_bytecode = True  # the reference mount is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _loaded = True
