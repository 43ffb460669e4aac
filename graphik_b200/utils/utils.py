"""Small host helpers with the reference's names (utils/utils.py)."""
import numpy as np


def wraptopi(e):
    return np.mod(e + np.pi, 2 * np.pi) - np.pi


def list_to_variable_dict(l, label="p", index_start=1):
    if isinstance(l, dict):
        return l
    return {label + str(index_start + i): v for i, v in enumerate(l)}


def table_environment(height=0.9, width=0.8, n_height=9, n_width=8, obs_inflation=2.0):
    """Spherical-obstacle model of a table (utils/utils.py:179-191): an n_width^2 grid of
    spheres for the top and n_height spheres per leg; returns [(centre, radius)]."""
    r = 0.5 * height / n_height
    half = n_width // 2
    grid = np.arange(-half, half)
    obs = [(np.asarray([2 * (i + 0.5) * r, 2 * (j + 0.5) * r, height + r]), obs_inflation * r)
           for i in grid for j in grid]
    for sx in (-1, 1):
        for sy in (-1, 1):
            cx, cy = sx * (width / 2 - r), sy * (width / 2 - r)
            obs += [(np.asarray([cx, cy, (2 * k + 1) * r]), obs_inflation * r) for k in range(n_height)]
    return obs
