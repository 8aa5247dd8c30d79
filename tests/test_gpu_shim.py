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


def test_remaining_hooks_match_the_python_versions(hooks):
    """is_footprint_colliding_impl, inverse_transform_2d_impl and native_project_poses against the oracle / NumPy"""
    cu, tu = hooks
    from brain.shining_utils import env_utils
    rng = np.random.RandomState(3)
    for _ in range(30):
        h, w = rng.randint(1, 70, 2)
        img = rng.choice([0, 0, 0, 253, 254, 255], size=(h, w)).astype(np.uint8)
        mask = rng.rand(h, w) < 0.2
        assert cu.is_footprint_colliding_impl(img, mask, 254) == bool(np.any(img[mask] == 254))
        big = np.zeros((200, 300), dtype=np.uint8)
        big[50:50 + h, 60:60 + w] = img
        assert cu.is_footprint_colliding_impl(big[50:50 + h, 60:60 + w], mask, 254) == bool(np.any(img[mask] == 254))   # a view
    t = rng.rand(200, 3) * 10. - 5.
    want = np.stack([-t[:, 0] * np.cos(t[:, 2]) - t[:, 1] * np.sin(t[:, 2]), t[:, 0] * np.sin(t[:, 2]) - t[:, 1] * np.cos(t[:, 2]),
                     O.normalize_angle(-t[:, 2])], axis=1)
    np.testing.assert_allclose(tu.inverse_transform_2d_impl(t), want, rtol=0, atol=1e-12)
    np.testing.assert_allclose(tu.inverse_transform_2d_impl(t[7]), want[7], rtol=0, atol=1e-12)
    assert tu.inverse_transform_2d_impl(t[7]).shape == (3,)
    poses = rng.rand(500, 3) * 1000. - 500.
    out = np.empty_like(poses)
    env_utils.native_project_poses(np.ascontiguousarray(t[3]), poses, out)
    c, s_ = np.cos(t[3, 2]), np.sin(t[3, 2])
    np.testing.assert_allclose(out[:, 0], c * poses[:, 0] - s_ * poses[:, 1] + t[3, 0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out[:, 1], s_ * poses[:, 0] + c * poses[:, 1] + t[3, 1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out[:, 2], O.normalize_angle(poses[:, 2] + t[3, 2]), rtol=0, atol=1e-9)


def test_reference_tests_pass_on_the_gpu_hooks():
    """The reference's OWN tests of the three hooks, run with the shim on sys.path (needs the reference package:
    oracle/_ref on the GPU box): test_costmap_utils.py::test_is_robot_colliding (:251-325) and the collision consistency
    tests beside it, test_coordinate_transformations.py::test_inverse_transform* (:105-178) and ::test_fast_project_poses*
    (:1298-1325)."""
    import os
    import subprocess
    import sys
    from oracle.ref_loader import reference_available
    if not reference_available():
        pytest.skip("the reference package is not importable here")
    here = os.path.dirname(os.path.abspath(__file__))
    for path, expr, n_pass in (("bc_gym_planning_env/utilities/test_costmap_utils.py", "is_robot_colliding", 1),
                               ("bc_gym_planning_env/utilities/test_coordinate_transformations.py",
                                "inverse_transform or fast_project_poses or project_poses_with_time", 6)):
        res = subprocess.run([sys.executable, os.path.join(here, "helpers", "run_reference_tests_with_shim.py"), path, expr],
                             capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
        assert "%d passed" % n_pass in res.stdout, res.stdout[-600:]


def test_ego_path_tensor_matches_the_reference_transform():
    """bcg_observe_ego_path: Observation.path in the robot frame for every env, against the oracle's restatement of
    from_global_to_egocentric on the same remaining paths (fixture aisle_delays_211: the reference's own worlds)."""
    import torch
    from tests import common
    d = common.load("aisle_delays_211")
    env = common.make_vec_env(d)
    oracles = common.make_oracles(d)
    actions = torch.from_numpy(d["actions"]).cuda()
    for t in range(40):
        env.step(actions[:, t].contiguous())
        for e, o in enumerate(oracles):
            obs = o.step(d["actions"][e, t])[0]
            if t % 13 == 12:
                got, left = env.observe_ego_path(max_points=48)
                want = O.ego_path(np.asarray(obs["path"]), obs["pose"])[:48]
                assert int(left[e]) == len(obs["path"]) == d["ref_path_len"][e, t]
                np.testing.assert_allclose(got[e, :len(want)].cpu().numpy(), want, rtol=0, atol=1e-9)
                assert not got[e, len(want):].any()
