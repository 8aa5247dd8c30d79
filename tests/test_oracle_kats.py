"""CPU: pins the oracle (oracle/plan_env_oracle.py) against the reference's own known-answer tests.
Vectors restated from /root/reference/bc_gym_planning_env/utilities/test_coordinate_transformations.py
(:32-83 normalize_angle, :1544-1649 world_to_pixel), test_path_tools.py (:215-272 path_velocity,
:275-293 pose_distances, :465-468 robot area) and test_costmap_utils.py (:251-306 is_robot_colliding)."""
import numpy as np
import pytest

from oracle import plan_env_oracle as O
from tests import common


def test_world_to_pixel_kats():
    w2p = O.world_to_pixel
    assert list(w2p([0., 0.], [0., 0.], 1.)) == [0, 0]
    assert list(w2p([0., 0.], [1., 1.], 1.)) == [-1, -1]
    assert list(w2p([-1., 0.], [0., 1.], 0.05)) == [-20, -20]
    assert w2p([[-1., 0.], [-1., 10.], [3., -7.]], [0., 1.], 0.05).tolist() == [[-20, -20], [-20, 180], [60, -160]]
    assert w2p([[0, 4.188]], [0., 0.], 0.03).tolist() == [[0, 140]]
    # half-to-even, like np.round
    assert list(w2p([0.5, -0.5], [0., 0.], 1.)) == [0, 0]
    assert list(w2p([1.5, -1.5], [0., 0.], 1.)) == [2, -2]
    # multiply by the reciprocal, do not divide
    assert list(w2p([0., 1.075], [0., 0.], 0.05)) == [0, 22]
    assert list(w2p([0., 4.275], [0., 0.], 0.03)) == [0, 143]
    assert list(w2p([0, 2.775], [0, 0.], 0.05)) == [0, 56]


def test_normalize_angle_kats():
    na = O.normalize_angle
    np.testing.assert_array_almost_equal(na(0.), 0.)
    np.testing.assert_array_almost_equal(na(np.pi / 2), np.pi / 2)
    np.testing.assert_array_almost_equal(na(np.pi + 0.1), -np.pi + 0.1)
    np.testing.assert_array_almost_equal(na(-np.pi - 0.1), np.pi - 0.1)
    np.testing.assert_array_almost_equal(
        na(np.array([-np.pi - 0.1, 2 * np.pi + 0.1, 99 * np.pi + 0.1, 100 * np.pi + 0.1, -1001 * np.pi - 0.3])),
        [np.pi - 0.1, 0.1, -np.pi + 0.1, 0.1, np.pi - 0.3])
    assert float(na(np.pi)) == -np.pi and float(na(-np.pi)) == -np.pi     # [-pi, pi)


@pytest.mark.parametrize("p0,p1,v,w", [
    ((0., 0, 0, 0), (1., 0, 0, 0), 0, 0),
    ((0., 0, 0, 0), (1., 1, 0, 0), 1, 0),
    ((0., 0, 0, 1), (1., 0, 1, 1), 1, 0),
    ((0., 0, 0, 0), (1., 0, 0, 1), 0, 1),
    ((0., 0, 0, 0), (1., 3, 4, 1), 5, 1),
    ((0, 0, 0, 0), (0.1, 3, 4, 1), 50, 10),
    ((0.1, 3, 4, 1), (0.2, 3, 5, -1), 10, -20),
    ((0, 0, 0, 0), (0.1, -3, -4, 1), -50, 10),
    ((0.1, -3, -4, 1), (0.2, 0, 0, 0), 50, -10),
    ((0, 0, 0, np.pi / 2. - 0.01), (1, 0, -0.1, np.pi / 2. - 0.01), -0.1, 0),
])
def test_measured_velocity_kats(p0, p1, v, w):
    got = O.measured_velocity(p0[1], p0[2], p0[3], p1[1], p1[2], p1[3], p1[0] - p0[0])
    np.testing.assert_array_almost_equal(got, (v, w))


def test_footprint_area_kat():
    footprint = np.array([[-0.77, -0.385], [-0.77, 0.385], [0.67, 0.385], [0.67, -0.385]])
    assert int((O.pixel_footprint(0., footprint, 0.05) != 0).sum()) == 493


def test_delay_line_kat():
    """envs/base/env.py:27-49 docstring: delay 3 over 0..8 prints 0 0 0 1 2 3 4 5 ... (off by the doc's typo)."""
    q = []
    assert [O.delay_line(q, i, 3) for i in range(9)] == [0, 0, 0, 0, 1, 2, 3, 4, 5]
    q = []
    assert [O.delay_line(q, i, 0) for i in range(4)] == [0, 1, 2, 3]
    q = []
    assert [O.delay_line(q, i, 2) for i in range(6)] == [0, 0, 0, 1, 2, 3]


def test_is_robot_colliding_golden_poses():
    """The 20 golden verdicts of test_costmap_utils.py:260-306 (the costmap itself comes from the fixture,
    rasterised by the reference's add_wall_to_static_map)."""
    d = common.load("kat_is_robot_colliding")
    expected = [False, True, True, False, True, True, True, False, False, False, False, True, True, True,
                False, True, True, False, False, False]
    assert d["ref_is_robot_colliding"][:20].tolist() == expected
    cm, origin, fp = d["costmap"], d["origin"], d["footprint"]
    for pose, want_irc, want_pc in zip(d["poses"], d["ref_is_robot_colliding"], d["ref_pose_collides"]):
        got = O.pose_collides(pose[0], pose[1], pose[2], fp, cm, origin, 0.05)
        assert got == want_pc
        px, py = O.world_to_pixel(pose[:2], origin, 0.05)
        in_bounds = 0 <= px < cm.shape[1] and 0 <= py < cm.shape[0]     # costmap_utils.py:150-152
        assert (got and in_bounds) == want_irc


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    assert O.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert O.philox4x32_10((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert O.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
