"""Pose / intrinsics value types with the reference's interface (libs/geometry/camera_modules.py:14-133).

They stay NumPy objects on purpose: the driver does matrix algebra on them directly
(dfvo.py:109-119,176-254), reading and *assigning* ``pose``, ``inv_pose``, ``R`` and ``t``.
"""
import numpy as np


class SE3:
    """4x4 rigid transform.  ``R`` / ``t`` are writable views ([3,3] and [3,1]); ``inv_pose`` is the
    matrix inverse, and assigning to it stores the inverse of the assigned value."""

    def __init__(self, np_arr=None):
        self._pose = np.eye(4) if np_arr is None else np_arr

    pose = property(lambda self: self._pose, lambda self, v: setattr(self, "_pose", v))

    @property
    def inv_pose(self):
        return np.linalg.inv(self._pose)

    @inv_pose.setter
    def inv_pose(self, value):
        self._pose = np.linalg.inv(value)

    @property
    def R(self):
        return self._pose[:3, :3]

    @R.setter
    def R(self, value):
        self._pose[:3, :3] = value

    @property
    def t(self):
        return self._pose[:3, 3:]

    @t.setter
    def t(self, value):
        self._pose[:3, 3:] = value


def _entry(i, j):
    def get(self):
        return self._mat[i, j]

    def set_(self, v):
        self._mat[i, j] = v
    return property(get, set_)


class Intrinsics:
    """3x3 pinhole intrinsics built from ``[cx, cy, fx, fy]`` (camera_modules.py:64-133)."""

    def __init__(self, param=None):
        if param is None:
            self._mat = np.zeros((3, 3))
        else:
            cx, cy, fx, fy = param
            self._mat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])

    mat = property(lambda self: self._mat, lambda self, m: setattr(self, "_mat", m))

    @property
    def inv_mat(self):
        return np.linalg.inv(self._mat)

    @inv_mat.setter
    def inv_mat(self, m):
        self._mat = np.linalg.inv(m)

    fx, fy, cx, cy = _entry(0, 0), _entry(1, 1), _entry(0, 2), _entry(1, 2)

    def as_list(self):
        """[cx, cy, fx, fy] -- the order the dfvo_b200 entry points take."""
        return [float(self.cx), float(self.cy), float(self.fx), float(self.fy)]


class PinholeCamera:
    """camera_modules.py:136-189."""

    def __init__(self, pose=None, K=None):
        self.height, self.width = 0, 0
        self.SE3 = SE3(pose)
        self.K = Intrinsics(K)
