"""Observation returned by step()/reset() (reference envs/base/obs.py:14-23) and the robot state
record it carries (robot_models/tricycle_model.py:234-290, differential_drive.py:77-125)."""
import attr
import numpy as np

from bc_gym_planning_env_b200.robot_models.robot_dimensions import INDUSTRIAL_DIFFDRIVE_V1, INDUSTRIAL_TRICYCLE_V1
from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D


@attr.s
class TricycleRobotState(object):
    x = attr.ib(default=0.0, type=float)
    y = attr.ib(default=0.0, type=float)
    angle = attr.ib(default=0.0, type=float)
    v = attr.ib(default=0.0, type=float)
    w = attr.ib(default=0.0, type=float)
    steering_motor_command = attr.ib(default=0.0, type=float)
    wheel_angle = attr.ib(default=0.0, type=float)

    VERSION = 1
    robot_type_name = INDUSTRIAL_TRICYCLE_V1

    def copy(self):
        return attr.evolve(self)

    def get_pose(self):
        return self.x, self.y, self.angle

    def set_pose(self, pose):
        self.x, self.y, self.angle = pose

    def to_numpy_array(self):
        return np.array([self.x, self.y, self.angle, self.v, self.w, self.wheel_angle], dtype=np.float64)

    def egocentric_state_numpy_array(self):
        return np.array([self.v, self.w, self.wheel_angle], dtype=np.float64)

    def old_style(self):
        return [self.wheel_angle, self.v, self.w, self.steering_motor_command]

    def get_robot_type_name(self):
        return self.robot_type_name

    def as_row(self):
        """The 7 fp64 state rows of include/bcg_b200.h (x, y, th, v, w, steer_cmd, wheel)."""
        return [self.x, self.y, self.angle, self.v, self.w, self.steering_motor_command, self.wheel_angle]

    @classmethod
    def from_row(cls, row):
        return cls(*[float(v) for v in row])

    def serialize(self):
        state = attr.asdict(self)
        state['version'] = self.VERSION
        return state

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        return cls(**state)


@attr.s
class DiffdriveRobotState(object):
    x = attr.ib(default=0.0, type=float)
    y = attr.ib(default=0.0, type=float)
    angle = attr.ib(default=0.0, type=float)
    v = attr.ib(default=0.0, type=float)
    w = attr.ib(default=0.0, type=float)

    VERSION = 1
    robot_type_name = INDUSTRIAL_DIFFDRIVE_V1

    def copy(self):
        return attr.evolve(self)

    def get_pose(self):
        return np.array([self.x, self.y, self.angle])

    def set_pose(self, pose):
        self.x, self.y, self.angle = pose

    def to_numpy_array(self):
        return np.array([self.x, self.y, self.angle, self.v, self.w], dtype=np.float64)

    def get_robot_type_name(self):
        return self.robot_type_name

    def as_row(self):
        return [self.x, self.y, self.angle, self.v, self.w, 0.0, 0.0]

    @classmethod
    def from_row(cls, row):
        return cls(*[float(v) for v in row[:5]])

    def serialize(self):
        state = attr.asdict(self)
        state['version'] = self.VERSION
        return state

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        return cls(**state)


def robot_state_type(robot_name):
    return DiffdriveRobotState if robot_name == INDUSTRIAL_DIFFDRIVE_V1 else TricycleRobotState


@attr.s(frozen=True, eq=False)
class Observation(object):
    pose = attr.ib(type=np.ndarray)                 # (delayed) oriented 2d pose of the robot
    path = attr.ib(type=np.ndarray, repr=False)     # path left to follow
    costmap = attr.ib(type=CostMap2D)
    time = attr.ib(type=float)
    dt = attr.ib(type=float)
    robot_state = attr.ib(type=object)              # (delayed) robot state record
    VERSION = 1

    def __eq__(self, other):
        """Field-wise equality like the reference (envs/base/obs.py:99-122)."""
        return (isinstance(other, Observation) and np.shape(self.pose) == np.shape(other.pose)
                and bool((np.asarray(self.pose) == np.asarray(other.pose)).all())
                and np.shape(self.path) == np.shape(other.path)
                and bool((np.asarray(self.path) == np.asarray(other.path)).all())
                and self.costmap == other.costmap and self.time == other.time and self.dt == other.dt
                and self.robot_state == other.robot_state)

    def __ne__(self, other):
        return not self.__eq__(other)

    def serialize(self):
        return dict(pose=self.pose, path=self.path, costmap=self.costmap.get_state(), time=self.time, dt=self.dt,
                    robot_state=self.robot_state.serialize(), robot_type_name=self.robot_state.get_robot_type_name(),
                    version=self.VERSION)

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        state['costmap'] = CostMap2D.from_state(state['costmap'])
        state['robot_state'] = robot_state_type(state.pop('robot_type_name')).deserialize(state['robot_state'])
        return cls(**state)
