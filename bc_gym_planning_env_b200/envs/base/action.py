"""Action: one motion primitive (reference envs/base/action.py:12-30).  For the tricycle
`command` = (front wheel linear velocity, desired front wheel angle); for diff-drive (v, w)."""
import attr
import numpy as np


@attr.s(eq=False)
class Action(object):
    VERSION = 1
    command = attr.ib(type=np.ndarray)

    @classmethod
    def from_cmds(cls, wanted_linear_velocity_of_baselink, wanted_front_wheel_angle):
        return cls(command=np.array([wanted_linear_velocity_of_baselink, wanted_front_wheel_angle]))

    def __eq__(self, other):
        return isinstance(other, Action) and not (np.asarray(self.command) != np.asarray(other.command)).any()

    def __ne__(self, other):
        return not self.__eq__(other)

    def serialize(self):
        return dict(command=self.command, version=self.VERSION)

    @classmethod
    def deserialize(cls, state):
        state = dict(state)
        assert state.pop('version') == cls.VERSION
        return cls(**state)
