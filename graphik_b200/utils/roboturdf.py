"""Robot loaders with the reference's names (utils/roboturdf.py:299-402).

The reference parses URDF files through `urdfpy` at call time.  The product
ships the resulting zero-configuration joint frames as small JSON models
(graphik_b200/robots/models/*.json, produced once by oracle/gen_golden.py from
the reference's own loader), so no URDF parser or reference tree is needed at
run time.  As in the reference, URDF joint limits are ignored: `limits=None`
means +-pi on every joint (roboturdf.py:362-364).
"""
import json
import os

import numpy as np

from graphik_b200.graphs.graph_revolute import ProblemGraphRevolute
from graphik_b200.robots.robot_revolute import RobotRevolute

_MODELS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "robots", "models")


def available_models():
    return sorted(f[:-5] for f in os.listdir(_MODELS) if f.endswith(".json"))


def load_model(name, limits=None, graph_params=None):
    with open(os.path.join(_MODELS, name + ".json")) as f:
        model = json.load(f)
    n = int(model["num_joints"])
    if limits is None:
        ub = np.ones(n) * np.pi
        lb = -ub
    else:
        lb, ub = limits[0], limits[1]
    params = {
        "T_zero": {"p%d" % i: np.array(T) for i, T in enumerate(model["T_zero"])},
        "num_joints": n,
        "joint_limits_upper": ub,
        "joint_limits_lower": lb,
    }
    robot = RobotRevolute(params)
    graph = ProblemGraphRevolute(robot, graph_params or {})
    return robot, graph


def load_ur10(limits=None, **kw):
    return load_model("ur10", limits, **kw)


def load_kuka(limits=None, **kw):
    return load_model("kuka", limits, **kw)


def load_schunk_lwa4d(limits=None, **kw):
    return load_model("lwa4d", limits, **kw)


def load_schunk_lwa4p(limits=None, **kw):
    return load_model("lwa4p", limits, **kw)


def load_panda(limits=None, **kw):
    return load_model("panda", limits, **kw)


def load_truncated_ur10(n: int):
    """First n links of a UR10 from its DH table (roboturdf.py:374-402)."""
    a = [0, -0.612, -0.5723, 0, 0, 0][:n]
    d = [0.1273, 0, 0, 0.1639, 0.1157, 0.0922][:n]
    al = [np.pi / 2, 0, 0, np.pi / 2, -np.pi / 2, 0][:n]
    params = {"a": a, "alpha": al, "d": d, "theta": [0] * n, "modified_dh": False, "num_joints": n}
    robot = RobotRevolute(params)
    return robot, ProblemGraphRevolute(robot)
