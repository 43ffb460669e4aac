"""Stand-in for the slice of `urdfpy` (utiasSTARS fork, reference setup.py:21)
that graphik/utils/roboturdf.py calls (:14,32-36,132-151): a plain ElementTree
reader for serial-chain URDFs.  TEST INFRASTRUCTURE ONLY.
"""
import xml.etree.ElementTree as ET

import numpy as np


def _rpy_to_matrix(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=float)
    n = np.linalg.norm(axis)
    if n == 0:
        return np.eye(3)
    a = axis / n
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


class Link:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "Link(%s)" % self.name


class JointLimit:
    def __init__(self, lower, upper):
        self.lower, self.upper = lower, upper


class Joint:
    def __init__(self, name, joint_type, parent, child, origin, axis, limit, mimic):
        self.name, self.joint_type = name, joint_type
        self.parent, self.child = parent, child
        self.origin, self.axis, self.limit, self.mimic = origin, axis, limit, mimic

    def __repr__(self):
        return "Joint(%s)" % self.name


class URDF:
    def __init__(self, links, joints):
        self.links, self.joints = links, joints
        self._link_map = {l.name: l for l in links}
        child_links = {j.child for j in joints}
        roots = [l for l in links if l.name not in child_links]
        self.base_link = roots[0]
        # actuated joints in base -> tip (BFS) order
        self.actuated_joints = []
        frontier = [self.base_link.name]
        while frontier:
            nxt = []
            for ln in frontier:
                for j in joints:
                    if j.parent == ln:
                        if j.joint_type in ("revolute", "continuous", "prismatic") and j.mimic is None:
                            self.actuated_joints.append(j)
                        nxt.append(j.child)
            frontier = nxt

    @staticmethod
    def load(path):
        root = ET.parse(path).getroot()
        links = [Link(e.get("name")) for e in root.findall("link")]
        joints = []
        for e in root.findall("joint"):
            org = e.find("origin")
            xyz = np.zeros(3)
            rpy = np.zeros(3)
            if org is not None:
                if org.get("xyz"):
                    xyz = np.array([float(v) for v in org.get("xyz").split()])
                if org.get("rpy"):
                    rpy = np.array([float(v) for v in org.get("rpy").split()])
            T = np.eye(4)
            T[:3, :3] = _rpy_to_matrix(rpy)
            T[:3, 3] = xyz
            ax = e.find("axis")
            axis = (np.array([float(v) for v in ax.get("xyz").split()])
                    if ax is not None else np.array([1.0, 0.0, 0.0]))
            lim = e.find("limit")
            limit = None
            if lim is not None:
                limit = JointLimit(float(lim.get("lower", 0.0)), float(lim.get("upper", 0.0)))
            joints.append(Joint(e.get("name"), e.get("type"), e.find("parent").get("link"),
                                e.find("child").get("link"), T, axis, limit, e.find("mimic")))
        return URDF(links, joints)

    def link_fk(self, cfg=None):
        cfg = cfg or {}
        fk = {self.base_link: np.eye(4)}
        frontier = [self.base_link.name]
        while frontier:
            nxt = []
            for ln in frontier:
                for j in self.joints:
                    if j.parent != ln:
                        continue
                    T = fk[self._link_map[ln]] @ j.origin
                    q = cfg.get(j.name, 0.0)
                    if j.joint_type in ("revolute", "continuous"):
                        M = np.eye(4)
                        M[:3, :3] = _axis_angle(j.axis, q)
                        T = T @ M
                    elif j.joint_type == "prismatic":
                        M = np.eye(4)
                        M[:3, 3] = np.asarray(j.axis) * q
                        T = T @ M
                    fk[self._link_map[j.child]] = T
                    nxt.append(j.child)
            frontier = nxt
        return fk
