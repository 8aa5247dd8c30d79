"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* upstream reference.

Imports braincorp/bc-gym-planning-env read-only from /root/reference (present in the
build container, absent on the GPU box) or from the install oracle/make_ref.py leaves in oracle/_ref/
(git-ignored; it travels to the GPU box) so that

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

HERE = os.path.dirname(os.path.abspath(__file__))
# where the reference may live, in order: an explicit override, the build container's read-only mount, the driver's
# `pip install --target baseline/_ref`, and this repo's own install (oracle/make_ref.py) -- the one that travels to the GPU box
CANDIDATE_ROOTS = [r for r in (os.environ.get("BCG_REFERENCE_ROOT"), "/root/reference",
                               os.path.join(os.path.dirname(HERE), "baseline", "_ref"), os.path.join(HERE, "_ref")) if r]


def reference_root():
    for root in CANDIDATE_ROOTS:
        if os.path.isfile(os.path.join(root, "bc_gym_planning_env", "envs", "base", "env.py")):
            return root
    return None


def reference_available():
    return reference_root() is not None


_loaded = False


def load_reference():
    """Make `import bc_gym_planning_env` resolve to the upstream tree.  Idempotent."""
    global _loaded
    if _loaded:
        return
    root = reference_root()
    if root is None:
        raise RuntimeError("reference package not found in any of %s" % CANDIDATE_ROOTS)
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
    if root not in sys.path:
        sys.path.insert(0, root)
    _loaded = True
