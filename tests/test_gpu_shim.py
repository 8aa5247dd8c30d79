"""-m gpu: the `brain.shining_utils` shim (SURVEY 8b seam ii) against the reference's own known-answer vectors
for the hooks it provides: utilities/test_coordinate_transformations.py:32-100 (normalize_angle, incl. the
TypeError contract of the native hook) and :1544-1649 (world_to_pixel), test_path_tools.py:453-468 (footprint)."""
import numpy as np
import pytest

from oracle import plan_env_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hooks():
    from bc_gym_planning_env_b200 import shim
    shim.install()
    from brain.shining_utils import costmap_utils, transform_utils      # resolves to the shim
    return costmap_utils, transform_utils


def test_normalize_angle_hook(hooks):
    _, tu = hooks
    f = tu.normalize_angle_impl
    for z, want in ((0., 0.), (np.pi / 2, np.pi / 2), (-np.pi / 2, -np.pi / 2), (np.pi + 0.1, -np.pi + 0.1), (-np.pi - 0.1, np.pi - 0.1)):
        np.testing.assert_array_almost_equal(f(z), want)
    np.testing.assert_array_almost_equal(f(np.array([0., 1, -2])), [0., 1, -2])
    np.testing.assert_array_almost_equal(
        f(np.array([-np.pi - 0.1, 2 * np.pi + 0.1, 99 * np.pi + 0.1, 100 * np.pi + 0.1, -1001 * np.pi - 0.3])),
        [np.pi - 0.1, 0.1, -np.pi + 0.1, 0.1, np.pi - 0.3])
    rng = np.random.RandomState(0)
    for a in rng.randn(40, 100) * 10. - 5:
        assert np.array_equal(f(a), O.normalize_angle(a))                 # bit-equal to the Python version
    assert f(np.pi) == O.normalize_angle(np.pi) and f(-np.pi) == O.normalize_angle(-np.pi)
    data = np.array([[0., np.pi + 0.1], [0., 0.2], [0., 2 * np.pi + 0.3]])
    with pytest.raises(TypeError):
        f(data)
    with pytest.raises(TypeError):
        f(data[:, 1])


def test_world_to_pixel_hook(hooks):
    cu, _ = hooks
    f = cu.world_to_pixel_impl
    origin = np.array([-2.75, 1.5])
    rng = np.random.RandomState(1)
    for shape in ((2,), (7, 2), (3, 5, 2), (0, 2)):
        xy = rng.uniform(-30, 30, size=shape)
        got = f(xy, origin, 0.03)
        assert got.shape == xy.shape and got.dtype == np.intc
        assert np.array_equal(got, O.world_to_pixel(xy, origin, 0.03))
    # half-way cases round to even, like np.round
    xy = np.array([[0.5, 1.5], [2.5, -0.5], [-1.5, 3.5]]) * 0.05
    assert np.array_equal(f(xy, np.zeros(2), 0.05), O.world_to_pixel(xy, np.zeros(2), 0.05))
    with pytest.raises(TypeError):
        f([[1., 2.]], origin, 0.03)
    with pytest.raises(ValueError):
        f(np.zeros((3, 3)), origin, 0.03)


def test_pixel_footprint_hook(hooks):
    cu, _ = hooks
    rect = np.array([[-0.77, -0.385], [-0.77, 0.385], [0.67, 0.385], [0.67, -0.385]])
    assert int(np.count_nonzero(cu.get_pixel_footprint_impl(0., rect, 0.05))) == 493   # test_compute_robot_area
    rng = np.random.RandomState(2)
    for a in rng.uniform(-np.pi, np.pi, 60):
        got = cu.get_pixel_footprint_impl(a, O.TRICYCLE_FOOTPRINT, 0.03)
        assert np.array_equal(got, O.pixel_footprint(a, O.TRICYCLE_FOOTPRINT, 0.03))
