"""Rigid transforms for the host-side robot model.

The reference leans on the external `liegroups` package for SE(3)
(robot_revolute.py:100, graph_revolute.py:243-318, geometry.py:26-43).  The
product needs only homogeneous 4x4 matrices on the host, so `SE3` here is a thin
wrapper over one ndarray with the handful of methods reference user code calls
(`as_matrix`, `trans`, `rot.as_matrix()`, `dot`, `inv`, `exp`, `from_matrix`).
Anything exposing `as_matrix()` (e.g. a liegroups SE3Matrix) or a raw 4x4 array
is accepted wherever a pose is expected -- see `as_matrix4`.
"""
import numpy as np


def hat(v):
    v = np.asarray(v, dtype=float).ravel()
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def as_matrix4(T):
    """4x4 ndarray from an SE3-like object or array."""
    if hasattr(T, "as_matrix"):
        T = T.as_matrix()
    T = np.asarray(T, dtype=float)
    if T.shape != (4, 4):
        raise ValueError("expected a 4x4 homogeneous transform, got shape %r" % (T.shape,))
    return T


class SO3:
    __slots__ = ("mat",)

    def __init__(self, mat):
        self.mat = np.array(mat, dtype=float)

    def as_matrix(self):
        return self.mat

    def inv(self):
        return SO3(self.mat.T)

    def dot(self, other):
        if isinstance(other, SO3):
            return SO3(self.mat @ other.mat)
        return self.mat @ np.asarray(other, dtype=float)

    @staticmethod
    def identity():
        return SO3(np.eye(3))

    @staticmethod
    def rot(axis, angle):
        c, s = np.cos(angle), np.sin(angle)
        if axis == "x":
            return SO3([[1, 0, 0], [0, c, -s], [0, s, c]])
        if axis == "y":
            return SO3([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        if axis == "z":
            return SO3([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        raise ValueError("Invalid Axis")

    rotx = staticmethod(lambda a: SO3.rot("x", a))
    roty = staticmethod(lambda a: SO3.rot("y", a))
    rotz = staticmethod(lambda a: SO3.rot("z", a))


class SE3:
    """Homogeneous transform; twist ordering [rho; phi] as in liegroups."""

    __slots__ = ("M",)

    def __init__(self, rot=None, trans=None):
        self.M = np.eye(4)
        if rot is not None:
            self.M[:3, :3] = rot.as_matrix() if hasattr(rot, "as_matrix") else np.asarray(rot, float)
        if trans is not None:
            self.M[:3, 3] = np.asarray(trans, dtype=float).ravel()

    @staticmethod
    def from_matrix(M):
        T = SE3()
        T.M = np.array(as_matrix4(M), dtype=float)
        return T

    @staticmethod
    def identity():
        return SE3()

    def as_matrix(self):
        return self.M

    @property
    def trans(self):
        return self.M[:3, 3]

    @trans.setter
    def trans(self, t):
        self.M[:3, 3] = np.asarray(t, dtype=float).ravel()

    @property
    def rot(self):
        return SO3(self.M[:3, :3])

    def inv(self):
        T = SE3()
        R = self.M[:3, :3]
        T.M[:3, :3] = R.T
        T.M[:3, 3] = -R.T @ self.M[:3, 3]
        return T

    def dot(self, other):
        if hasattr(other, "as_matrix"):
            return SE3.from_matrix(self.M @ as_matrix4(other))
        p = np.asarray(other, dtype=float)
        if p.shape[-1] == 3:
            return p @ self.M[:3, :3].T + self.M[:3, 3]
        return p @ self.M.T

    @staticmethod
    def exp(xi):
        xi = np.asarray(xi, dtype=float).ravel()
        return SE3.from_matrix(twist_exp(xi[3:6], xi[0:3], 1.0))

    def __repr__(self):
        return "SE3(\n%r)" % (self.M,)


def twist_exp(omega, v, theta):
    """exp of the screw (v, omega)*theta as a 4x4 matrix; omega need not be unit."""
    omega = np.asarray(omega, dtype=float) * theta
    v = np.asarray(v, dtype=float) * theta
    angle = np.linalg.norm(omega)
    M = np.eye(4)
    if angle < 1e-12:
        M[:3, :3] += hat(omega)
        M[:3, 3] = v + 0.5 * np.cross(omega, v)
        return M
    a = omega / angle
    s, c = np.sin(angle), np.cos(angle)
    A = hat(a)
    aaT = np.outer(a, a)
    M[:3, :3] = c * np.eye(3) + (1 - c) * aaT + s * A
    J = (s / angle) * np.eye(3) + (1 - s / angle) * aaT + ((1 - c) / angle) * A
    M[:3, 3] = J @ v
    return M


def trans_axis(t, axis="z"):
    k = {"x": 0, "y": 1, "z": 2}.get(axis)
    if k is None:
        raise Exception("Invalid Axis")
    e = np.zeros(3)
    e[k] = t
    return SE3(None, e)


def rot_axis(theta, axis="z"):
    if axis not in ("x", "y", "z"):
        raise Exception("Invalid Axis")
    return SE3(SO3.rot(axis, theta), None)
