"""Robot constant tables: footprint polygons and kinematic limits of the two example robots.

Values are the reference's data (robot_models/robot_dimensions_examples.py:51-85 diff-drive,
:109-190 tricycle); the layout here is a flat record that the host flattens into BcgParams.
"""
import attr
import numpy as np

INDUSTRIAL_TRICYCLE_V1 = 'industrial_tricycle_v1'      # standard_robot_names_examples.py
INDUSTRIAL_DIFFDRIVE_V1 = 'industrial_diffdrive_v1'
TRICYCLE, DIFF = 'tricycle', 'diff'                    # robot_drive_types.py


@attr.s(frozen=True)
class RobotDimensions(object):
    name = attr.ib(type=str)
    drive_type = attr.ib(type=str)
    footprint_mm = attr.ib(type=tuple, repr=False)
    front_wheel_from_axis = attr.ib(type=float, default=0.0)
    max_front_wheel_angle = attr.ib(type=float, default=0.0)
    max_front_wheel_speed = attr.ib(type=float, default=0.0)
    max_linear_acceleration = attr.ib(type=float, default=0.0)
    max_angular_acceleration = attr.ib(type=float, default=0.0)
    front_column_model_p_gain = attr.ib(type=float, default=0.0)

    def footprint(self):
        """(n, 2) fp64 polygon in metres, bumper front-centre first."""
        return np.array(self.footprint_mm, dtype=np.float64) / 1000.

    def get_name(self):
        return self.name


_TRICYCLE_V1 = RobotDimensions(
    name=INDUSTRIAL_TRICYCLE_V1, drive_type=TRICYCLE,
    footprint_mm=(
        (1348.35, 0.), (1338.56, 139.75), (1306.71, 280.12), (1224.36, 338.62), (1093.81, 374.64),
        (-214.37, 374.64), (-313.62, 308.56), (-366.36, 117.44), (-374.01, -135.75), (-227.96, -459.13),
        (-156.72, -458.78), (759.8, -442.96), (849.69, -426.4), (1171.05, -353.74), (1303.15, -286.54),
        (1341.34, -118.37)),
    front_wheel_from_axis=0.964,
    max_front_wheel_angle=0.5 * 170 * np.pi / 180.,
    max_front_wheel_speed=60. * np.pi / 180.,
    max_linear_acceleration=1. / 2.5,
    max_angular_acceleration=1. / 2.,
    front_column_model_p_gain=0.16,
)

_DIFFDRIVE_V1 = RobotDimensions(
    name=INDUSTRIAL_DIFFDRIVE_V1, drive_type=DIFF,
    footprint_mm=(
        (644.5, 0), (634.86, 61), (571.935, 130.54), (553.38, 161), (360.36, 186), (250, 186), (250, 186),
        (100, 186), (100, 186), (0, 196), (-119.21, 190.5), (-173.4, 146), (-193, 0), (-173.4, -143),
        (-111.65, -246), (-71.57, -246), (100, -246), (100, -246), (250, -246), (250, -246),
        (413.085, -223), (491.5, -204.5), (553, -161), (634.86, -62)),
)

_BY_NAME = {INDUSTRIAL_TRICYCLE_V1: _TRICYCLE_V1, INDUSTRIAL_DIFFDRIVE_V1: _DIFFDRIVE_V1}


def get_dimensions_example(footprint_name):
    """Same lookup (and same AssertionError on unknown names) as
    robot_dimensions_examples.py:14-30."""
    if footprint_name not in _BY_NAME:
        raise AssertionError("Unknown footprint {}. Should be one of {}".format(footprint_name, list(_BY_NAME)))
    return _BY_NAME[footprint_name]
