"""The reference's OWN tests, run against this repo's `brain.shining_utils` shim (SURVEY 8b seam ii).  Needs
/root/reference (build container); the hooks that need a GPU are covered by tests/test_gpu_shim.py instead."""
import os
import subprocess
import sys

import pytest

from oracle.ref_loader import reference_available

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not reference_available(), reason="the reference tree is only present in the build container")
def test_reference_footprint_tests_pass_on_the_hook():
    """utilities/test_path_tools.py::test_get_pixel_footprint_consistency (8 000 angles, filled and contour, two
    footprints, two resolutions: hook array == the reference's Python array) and ::test_compute_robot_area."""
    res = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "run_reference_tests_with_shim.py"),
                          "bc_gym_planning_env/utilities/test_path_tools.py", "footprint_consistency or compute_robot_area"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "2 passed" in res.stdout


def test_shim_modules_resolve_without_touching_the_gpu():
    from bc_gym_planning_env_b200 import shim
    path = shim.install()
    assert os.path.isdir(os.path.join(path, "brain", "shining_utils"))
    import importlib
    cu = importlib.import_module("brain.shining_utils.costmap_utils")
    tu = importlib.import_module("brain.shining_utils.transform_utils")
    assert callable(cu.world_to_pixel_impl) and callable(cu.get_pixel_footprint_impl) and callable(tu.normalize_angle_impl)
    eu = importlib.import_module("brain.shining_utils.env_utils")
    assert callable(cu.is_footprint_colliding_impl) and callable(tu.inverse_transform_2d_impl) and callable(eu.native_project_poses)
