"""Host-side model of a serial revolute manipulator.

Mirrors the slice of the reference's `Robot` / `RobotRevolute`
(robots/robot_base.py:18-50,76-98, robots/robot_revolute.py:15-103) that the
Riemannian IK path touches: zero-configuration joint frames T0[p_i], screw axes
S_i = [-w x q ; w] taken from the z axis of T0 (robot_revolute.py:41-44) and the
product-of-exponentials forward kinematics (robot_revolute.py:85-103).

Unlike the reference this is NOT a networkx graph: the kinematic chain is stored
as dense arrays so that forward kinematics is batched over thousands of
configurations (the benchmark's goal generator), and so that the frames can be
packed into the device-side plan.  Only serial chains are supported (all
BASELINE configs are chains); tree-structured robots raise.
"""
from math import pi
from typing import Any, Dict, List

import numpy as np

from graphik_b200.utils.se3 import SE3, as_matrix4, rot_axis, trans_axis

ROOT = "p0"


def _as_label_dict(values, n, first=1):
    if isinstance(values, dict):
        return dict(values)
    vals = np.asarray(values, dtype=float).ravel()
    return {"p%d" % (first + i): float(vals[i]) for i in range(len(vals))}


def _dh_frame(a, alpha, d, theta, modified):
    """One DH link (kinematics.py:39-85)."""
    tx, rx, tz, rz = trans_axis(a, "x"), rot_axis(alpha, "x"), trans_axis(d, "z"), rot_axis(theta, "z")
    if modified:
        return tx.dot(rx.dot(tz.dot(rz)))
    return tz.dot(rz.dot(tx.dot(rx)))


class RobotRevolute:
    dim = 3

    def __init__(self, params: Dict[str, Any]):
        self.params = params
        self.n = int(params["num_joints"])
        if "parents" in params:
            par = params["parents"]
            if any(len(ch) > 1 for ch in par.values()):
                raise NotImplementedError("graphik_b200 supports serial chains only")
        n = self.n
        self.joint_ids: List[str] = ["p%d" % i for i in range(n + 1)]
        self.end_effectors: List[str] = ["p%d" % n]
        self.lb = _as_label_dict(params.get("joint_limits_lower", n * [-pi]), n)
        self.ub = _as_label_dict(params.get("joint_limits_upper", n * [pi]), n)

        if "T_zero" in params:
            tz = params["T_zero"]
            if isinstance(tz, dict):
                frames = [as_matrix4(tz["p%d" % i]) for i in range(n + 1)]
            else:
                frames = [as_matrix4(T) for T in tz]
        elif all(k in params for k in ("a", "d", "alpha", "theta", "modified_dh")):
            frames = self._frames_from_dh(params)
        else:
            raise Exception("Robot description not provided.")
        self.T0 = np.stack(frames).astype(float)  # [n+1,4,4]

        w = self.T0[:, :3, 2]
        q = self.T0[:, :3, 3]
        self.S = np.hstack([np.cross(-w, q), w])  # [n+1,6] rows: [v ; omega]
        # relative zero-config transforms pred -> cur (robot_revolute.py:47-51)
        self.T_rel = np.stack([np.linalg.inv(self.T0[i - 1]) @ self.T0[i] for i in range(1, n + 1)])

    # -- construction ------------------------------------------------------
    def _frames_from_dh(self, params):
        n = self.n
        cols = {k: _as_label_dict(params[k], n) for k in ("a", "d", "alpha", "theta")}
        mod = bool(params["modified_dh"])
        frames = [np.eye(4)]
        T = SE3.identity()
        for i in range(1, n + 1):
            lab = "p%d" % i
            T = T.dot(_dh_frame(cols["a"][lab], cols["alpha"][lab], cols["d"][lab], cols["theta"][lab], mod))
            frames.append(T.as_matrix().copy())
        return frames

    # -- reference-compatible accessors --------------------------------------
    @property
    def T_base(self) -> SE3:
        return SE3.from_matrix(self.T0[0])

    @property
    def spherical(self) -> bool:
        return False

    def T_zero(self, node: str) -> SE3:
        return SE3.from_matrix(self.T0[int(node[1:])])

    def random_configuration(self) -> Dict[str, float]:
        """robot_base.py:76-85 -- same draw order from numpy's global RNG."""
        q = {}
        for key in self.joint_ids:
            if key != ROOT:
                q[key] = self.lb[key] + (self.ub[key] - self.lb[key]) * np.random.rand()
        return q

    def zero_configuration(self) -> Dict[str, float]:
        return {key: 0 for key in self.joint_ids if key != ROOT}

    def q_array(self, joint_angles) -> np.ndarray:
        if isinstance(joint_angles, dict):
            return np.array([joint_angles["p%d" % i] for i in range(1, self.n + 1)], dtype=float)
        return np.asarray(joint_angles, dtype=float)

    def q_dict(self, q) -> Dict[str, float]:
        return {"p%d" % (i + 1): float(q[i]) for i in range(self.n)}

    # -- forward kinematics ---------------------------------------------------
    def fk_all(self, Q) -> np.ndarray:
        """Frames of p0..pn for a batch of configurations.  Q[B,n] -> T[B,n+1,4,4].

        pose(q, p_k) = T0[p0] * prod_{i<k} exp(S_i q_{i+1}) * T0[p_k]
        (robot_revolute.py:96-103)."""
        Q = np.atleast_2d(np.asarray(Q, dtype=float))
        B, n = Q.shape[0], self.n
        out = np.empty((B, n + 1, 4, 4))
        acc = np.broadcast_to(self.T0[0], (B, 4, 4)).copy()
        out[:, 0] = acc @ self.T0[0]
        for i in range(n):
            acc = acc @ _screw_exp(self.S[i], Q[:, i])
            out[:, i + 1] = acc @ self.T0[i + 1]
        return out

    def pose(self, joint_angles, query_node: str) -> SE3:
        k = int(query_node[1:])
        T = self.fk_all(self.q_array(joint_angles)[None, :])[0, k]
        return SE3.from_matrix(T)

    def get_all_poses(self, joint_angles) -> Dict[str, SE3]:
        T = self.fk_all(self.q_array(joint_angles)[None, :])[0]
        return {"p%d" % i: SE3.from_matrix(T[i]) for i in range(self.n + 1)}


def _screw_exp(S, theta):
    """Batched exp(S*theta) for a unit-rotation screw S=[v;w].  theta[B] -> [B,4,4]."""
    v, w = S[:3], S[3:]
    theta = np.asarray(theta, dtype=float)
    B = theta.shape[0]
    W = np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])
    wwT = np.outer(w, w)
    s, c = np.sin(theta)[:, None, None], np.cos(theta)[:, None, None]
    th = theta[:, None, None]
    M = np.zeros((B, 4, 4))
    M[:, 3, 3] = 1.0
    M[:, :3, :3] = c * np.eye(3) + (1 - c) * wwT + s * W
    # translation = (sin I + (theta - sin) w w^T + (1-cos) W) v   (|w| = 1)
    Jv = s * np.eye(3) + (th - s) * wwT + (1 - c) * W
    M[:, :3, 3] = Jv @ v
    return M
