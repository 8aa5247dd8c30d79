"""Run tests OF THE REFERENCE (read-only tree under /root/reference) with the `brain.shining_utils` shim on
sys.path, so that the reference's own call sites bind to this repo's hooks.  Usage:
    python tests/helpers/run_reference_tests_with_shim.py <reference test file> <-k expression>
Exit status = pytest's.  Only hooks that need no GPU are exercised this way (the build container has none)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference, reference_root  # noqa: E402

load_reference()
from bc_gym_planning_env_b200 import shim  # noqa: E402

shim.install()
import bc_gym_planning_env.utilities.path_tools as path_tools  # noqa: E402
from brain.shining_utils.costmap_utils import get_pixel_footprint_impl  # noqa: E402

assert path_tools.get_pixel_footprint is get_pixel_footprint_impl, "the reference did not pick the hook up"
import pytest  # noqa: E402

# the reference's pytest.ini turns numpy's deprecation warnings into errors; its tests are run as they are otherwise
sys.exit(pytest.main(['-q', '-p', 'no:cacheprovider', '-c', os.devnull, '--rootdir', '/tmp', '-W', 'ignore::DeprecationWarning',
                      os.path.join(reference_root(), sys.argv[1]), '-k', sys.argv[2]]))
