"""Error metrics over the per-episode CSV stream (second half).

The reference's ``src/utils/error.py`` (``calculate_error_metrics(model_name)``, :148-190) reads the files an episode with
``save_data=True`` leaves in ``simulation_data/`` and ``reward_data/`` (``<quantity>_<model_name>.csv``, written by
``Ultrasound._save_data``, ultrasound.py:890-910 -- here by :meth:`rui_b200.env.Ultrasound._save_data`) and writes one
one-number CSV per metric into ``error_data/<model_name>/``.  This module produces the same files with the same numbers, from one
table instead of one function per quantity:

    mean-squared errors   x_pos_mse, y_pos_mse (eef x / y against the trajectory point, error.py:33-52), force_mse, mean_force_mse
                          (:55-70), der_force_mse (:73-86), velocity_mse (norm of the eef velocity), mean_velocity_mse (:89-105)
    means                 quat_diff_mean (:135-145), {pos, ori, force, der_force, vel}_reward_mean (:108-132)
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import pandas as pd

SIM, REW = "simulation_data", "reward_data"

# metric -> (folder, measured quantity, goal quantity or None, column or None, reduce the measured rows by their norm first)
_MSE = {
    "x_pos_mse": ("ee_pos", "ee_goal_pos", 0, False),
    "y_pos_mse": ("ee_pos", "ee_goal_pos", 1, False),
    "force_mse": ("ee_z_contact_force", "ee_z_goal_contact_force", None, False),
    "mean_force_mse": ("ee_z_running_mean_contact_force", "ee_z_goal_contact_force", None, False),
    "der_force_mse": ("ee_z_derivative_contact_force", "ee_z_goal_derivative_contact_force", None, False),
    "velocity_mse": ("ee_vel", "ee_goal_vel", None, True),
    "mean_velocity_mse": ("ee_running_mean_vel", "ee_goal_vel", None, False),
}
_MEAN = {
    "quat_diff_mean": (SIM, "ee_diff_quat"),
    "pos_reward_mean": (REW, "pos"),
    "ori_reward_mean": (REW, "ori"),
    "force_reward_mean": (REW, "force"),
    "der_force_reward_mean": (REW, "derivative_force"),
    "vel_reward_mean": (REW, "vel"),
}


def _read(folder: str, quantity: str, model_name: str, root: str) -> np.ndarray:
    a = pd.read_csv(os.path.join(root, folder, f"{quantity}_{model_name}.csv"), header=None).to_numpy(dtype=np.float64)
    return a


def compute_error_metrics(model_name: str, root: str = ".") -> Dict[str, float]:
    """All metrics of one recorded episode as a dict (nothing is written)."""
    out: Dict[str, float] = {}
    for name, (meas, goal, col, use_norm) in _MSE.items():
        m, g = _read(SIM, meas, model_name, root), _read(SIM, goal, model_name, root)
        if col is not None:
            m, g = m[:, col], g[:, col]
        elif use_norm:
            m, g = np.linalg.norm(m, axis=1), g[:, 0]
        else:
            m, g = m[:, 0], g[:, 0]
        out[name] = float(np.mean(np.square(m - g)))
    for name, (folder, quantity) in _MEAN.items():
        out[name] = float(_read(folder, quantity, model_name, root)[:, 0].mean())
    return out


def calculate_error_metrics(model_name: str, root: str = ".") -> Dict[str, float]:
    """``utils/error.py:calculate_error_metrics``: compute the metrics and write ``error_data/<model_name>/<metric>.csv``
    (one number per file, no header, no index -- the reference's ``save_data`` layout, error.py:5-16)."""
    metrics = compute_error_metrics(model_name, root)
    folder = os.path.join(root, "error_data", str(model_name))
    os.makedirs(folder, exist_ok=True)
    for name, value in metrics.items():
        pd.DataFrame(np.array([value])).to_csv(os.path.join(folder, name + ".csv"), header=None, index=None)
    return metrics
