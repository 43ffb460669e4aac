"""Minimal stand-in for the utiasSTARS `liegroups` package (numpy backend).

TEST INFRASTRUCTURE ONLY.  The reference (GraphIK, setup.py:24) depends on
`liegroups @ utiasSTARS/liegroups@generative_ik`, which is not vendored under
/root/reference and is not installable offline.  This file restates the small,
standard slice of SO(2)/SO(3)/SE(2)/SE(3) algebra the reference's revolute
path calls (robot_revolute.py:100, graph_revolute.py:120-150,243-318,
geometry.py:26-43, roboturdf.py:137-295) so that the unmodified reference can
be imported in this container to generate golden vectors.  Conventions follow
the public liegroups API: twists are ordered [rho (translation); phi
(rotation)], `dot` composes groups or transforms points.
"""
import numpy as np


def _wedge3(phi):
    phi = np.asarray(phi, dtype=float).ravel()
    return np.array([[0.0, -phi[2], phi[1]],
                     [phi[2], 0.0, -phi[0]],
                     [-phi[1], phi[0], 0.0]])


class SO3Matrix:
    dim, dof = 3, 3

    def __init__(self, mat):
        self.mat = np.array(mat, dtype=float)

    @classmethod
    def identity(cls):
        return cls(np.eye(3))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        return cls(mat)

    def as_matrix(self):
        return self.mat

    @classmethod
    def wedge(cls, phi):
        return _wedge3(phi)

    @classmethod
    def vee(cls, Phi):
        return np.array([Phi[2, 1], Phi[0, 2], Phi[1, 0]])

    @classmethod
    def exp(cls, phi):
        phi = np.asarray(phi, dtype=float).ravel()
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.0):
            return cls(np.eye(3) + _wedge3(phi))
        axis = phi / angle
        s, c = np.sin(angle), np.cos(angle)
        return cls(c * np.eye(3) + (1 - c) * np.outer(axis, axis) + s * _wedge3(axis))

    def log(self):
        cos_angle = np.clip(0.5 * np.trace(self.mat) - 0.5, -1.0, 1.0)
        angle = np.arccos(cos_angle)
        if np.isclose(angle, 0.0):
            return self.vee(self.mat - np.eye(3))
        return self.vee((0.5 * angle / np.sin(angle)) * (self.mat - self.mat.T))

    @classmethod
    def left_jacobian(cls, phi):
        phi = np.asarray(phi, dtype=float).ravel()
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.0):
            return np.eye(3) + 0.5 * _wedge3(phi)
        axis = phi / angle
        s, c = np.sin(angle), np.cos(angle)
        return ((s / angle) * np.eye(3) + (1 - s / angle) * np.outer(axis, axis)
                + ((1 - c) / angle) * _wedge3(axis))

    @classmethod
    def inv_left_jacobian(cls, phi):
        return np.linalg.inv(cls.left_jacobian(phi))

    @classmethod
    def rotx(cls, a):
        c, s = np.cos(a), np.sin(a)
        return cls(np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]]))

    @classmethod
    def roty(cls, a):
        c, s = np.cos(a), np.sin(a)
        return cls(np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]]))

    @classmethod
    def rotz(cls, a):
        c, s = np.cos(a), np.sin(a)
        return cls(np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]))

    def inv(self):
        return self.__class__(self.mat.T)

    def adjoint(self):
        return self.mat

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.mat.dot(other.mat))
        other = np.asarray(other, dtype=float)
        if other.ndim == 1:
            return self.mat.dot(other)
        return self.mat.dot(other.T).T

    def __repr__(self):
        return "SO3Matrix(\n%r)" % (self.mat,)


class SO2Matrix:
    dim, dof = 2, 1

    def __init__(self, mat):
        self.mat = np.array(mat, dtype=float)

    @classmethod
    def identity(cls):
        return cls(np.eye(2))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        return cls(mat)

    @classmethod
    def from_angle(cls, a):
        c, s = np.cos(a), np.sin(a)
        return cls(np.array([[c, -s], [s, c]]))

    exp = from_angle

    def to_angle(self):
        return np.arctan2(self.mat[1, 0], self.mat[0, 0])

    log = to_angle

    @classmethod
    def wedge(cls, phi):
        return np.array([[0.0, -float(phi)], [float(phi), 0.0]])

    @classmethod
    def left_jacobian(cls, phi):
        phi = float(phi)
        if np.isclose(phi, 0.0):
            return np.eye(2) + 0.5 * cls.wedge(phi)
        s, c = np.sin(phi), np.cos(phi)
        return (s / phi) * np.eye(2) + ((1 - c) / phi) * cls.wedge(1.0)

    def as_matrix(self):
        return self.mat

    def inv(self):
        return self.__class__(self.mat.T)

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.mat.dot(other.mat))
        other = np.asarray(other, dtype=float)
        if other.ndim == 1:
            return self.mat.dot(other)
        return self.mat.dot(other.T).T


class _SEBase:
    RotationType = None
    dim = None

    def __init__(self, rot, trans):
        self.rot = rot
        self.trans = np.array(trans, dtype=float)

    @classmethod
    def identity(cls):
        return cls(cls.RotationType.identity(), np.zeros(cls.dim - 1))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        mat = np.asarray(mat, dtype=float)
        d = cls.dim - 1
        return cls(cls.RotationType(mat[:d, :d]), mat[:d, d])

    def as_matrix(self):
        d = self.dim - 1
        M = np.eye(self.dim)
        M[:d, :d] = self.rot.as_matrix()
        M[:d, d] = np.asarray(self.trans, dtype=float).ravel()
        return M

    def inv(self):
        inv_rot = self.rot.inv()
        return self.__class__(inv_rot, -inv_rot.dot(self.trans))

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.rot.dot(other.rot),
                                  self.rot.dot(other.trans) + self.trans)
        other = np.asarray(other, dtype=float)
        d = self.dim - 1
        if other.ndim == 1 and other.shape[0] == d:
            return self.rot.dot(other) + self.trans
        if other.ndim == 1 and other.shape[0] == self.dim:
            return self.as_matrix().dot(other)
        if other.ndim == 2 and other.shape[1] == d:
            return self.rot.dot(other) + self.trans
        if other.ndim == 2 and other.shape[1] == self.dim:
            return self.as_matrix().dot(other.T).T
        raise ValueError("cannot apply transform to array of shape %r" % (other.shape,))

    def __repr__(self):
        return "%s(\n%r)" % (self.__class__.__name__, self.as_matrix())


class SE3Matrix(_SEBase):
    RotationType = SO3Matrix
    dim, dof = 4, 6

    @classmethod
    def wedge(cls, xi):
        xi = np.asarray(xi, dtype=float).ravel()
        Xi = np.zeros((4, 4))
        Xi[:3, :3] = _wedge3(xi[3:6])
        Xi[:3, 3] = xi[0:3]
        return Xi

    @classmethod
    def exp(cls, xi):
        xi = np.asarray(xi, dtype=float).ravel()
        rho, phi = xi[0:3], xi[3:6]
        return cls(SO3Matrix.exp(phi), SO3Matrix.left_jacobian(phi).dot(rho))

    def log(self):
        phi = self.rot.log()
        rho = SO3Matrix.inv_left_jacobian(phi).dot(self.trans)
        return np.hstack([rho, phi])

    def adjoint(self):
        R = self.rot.as_matrix()
        A = np.zeros((6, 6))
        A[:3, :3] = R
        A[:3, 3:] = _wedge3(self.trans).dot(R)
        A[3:, 3:] = R
        return A


class SE2Matrix(_SEBase):
    RotationType = SO2Matrix
    dim, dof = 3, 3

    @classmethod
    def exp(cls, xi):
        xi = np.asarray(xi, dtype=float).ravel()
        rho, phi = xi[0:2], xi[2]
        return cls(SO2Matrix.exp(phi), SO2Matrix.left_jacobian(phi).dot(rho))

    def adjoint(self):
        R = self.rot.as_matrix()
        A = np.eye(3)
        A[:2, :2] = R
        A[0, 2] = self.trans[1]
        A[1, 2] = -self.trans[0]
        return A
