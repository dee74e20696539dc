"""ctypes mirror of ``include/usim.h`` (structs + constants).

Kept free of any library loading so that the CPU oracle wrapper (tests only)
can reuse the struct definitions without touching the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List

import numpy as np

from .model import UltrasoundModel

USIM_ABI_VERSION = 2
OBS_DIM = 19
TASK_DIM = 48
MAX_CONTACTS = 128
DIAG_DIM = 28
ARM_RECORD_DIM = 180  # usim_get_arm_record row pitch (usim.h)

GOAL_QUAT_XYZW = (-0.69192486, 0.72186726, -0.00514253, -0.01100909)  # ultrasound.py:174

MODE_FIXED, MODE_TRACKING, MODE_VARIABLE_Z, MODE_WRENCH = 0, 1, 2, 3
MODE_BY_NAME = {"fixed": MODE_FIXED, "tracking": MODE_TRACKING, "variable_z": MODE_VARIABLE_Z, "wrench": MODE_WRENCH}

# task-state record indices (enum in usim.h)
TS_TRAJ_START, TS_TRAJ_END, TS_U0, TS_STIFFNESS, TS_DAMPING = 0, 3, 6, 7, 8
TS_VEL_MEAN, TS_FZ_MEAN, TS_FZ_PREV, TS_DFZ, TS_TOUCHED, TS_TIMESTEP = 9, 10, 11, 12, 13, 14
TS_TRAJ_PT, TS_INIT_JOINT, TS_EPISODE, TS_DONE, TS_POS_ERR, TS_ORI_ERR, TS_IN_CONTACT = 15, 18, 25, 26, 27, 29, 30
TS_GOAL_ORI, TS_GOAL_POS = 32, 41

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class UsimModel(C.Structure):
    _fields_ = [
        ("nbody", C.c_int32), ("nq", C.c_int32), ("nv", C.c_int32), ("npart", C.c_int32), ("npair", C.c_int32), ("soft", C.c_int32),
        ("narm", C.c_int32), ("reserved0", C.c_int32),
        ("table_body", C.c_int32), ("link1_body", C.c_int32), ("hand_body", C.c_int32), ("probe_body", C.c_int32),
        ("torso_body", C.c_int32), ("part_body0", C.c_int32),
        ("body_parent", _pi), ("body_jnt_type", _pi), ("body_qposadr", _pi), ("body_dofadr", _pi),
        ("body_pos", _pd), ("body_quat", _pd), ("body_mass", _pd), ("body_ipos", _pd), ("body_inertia", _pd), ("body_jnt_axis", _pd),
        ("dof_damping", _pd), ("jnt_range", _pd), ("ctrl_range", _pd), ("qpos0", _pd),
        ("probe_seg", _pd), ("part_pos", _pd), ("part_axis", _pd), ("part_seg_outer", _pd), ("part_seg_inner", _pd),
        ("eq_pairs", _pi), ("part_nbr", _pi),
        ("dof_invweight0", _pd), ("body_invweight0", _pd),
        ("arm_link", _pd), ("arm_tool", _pd),
        ("probe_radius", C.c_double), ("cap_radius", C.c_double), ("tendon_invweight0", C.c_double),
        ("timestep", C.c_double), ("gravity", C.c_double * 3), ("impratio", C.c_double), ("solref", C.c_double * 2),
        ("solimp", C.c_double * 5), ("solref_smooth", C.c_double * 2),
        ("table_top_z", C.c_double), ("table_half_xy", C.c_double), ("table_friction", C.c_double),
        ("probe_friction", C.c_double), ("particle_friction", C.c_double),
        ("init_qpos", C.c_double * 7),
        ("top_torso_offset", C.c_double), ("traj_x_range", C.c_double), ("traj_y_range", C.c_double),
    ]


class UsimConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("num_envs", C.c_int32), ("env_id_offset", C.c_int32), ("impedance_mode", C.c_int32),
        ("horizon", C.c_int32), ("early_termination", C.c_int32), ("solref_randomization", C.c_int32),
        ("probe_pos_randomization", C.c_int32), ("deterministic_trajectory", C.c_int32), ("uncouple_pos_ori", C.c_int32),
        ("solver_iterations", C.c_int32), ("precond_rebuilds", C.c_int32), ("ignore_done", C.c_int32), ("reserved0", C.c_int32),
        ("seed", C.c_uint64),
        ("control_freq", C.c_double),
        ("kp", C.c_double * 6), ("damping_ratio", C.c_double * 6),
        ("input_max", C.c_double), ("input_min", C.c_double),
        ("output_max", C.c_double * 6), ("output_min", C.c_double * 6),
        ("kp_limits", C.c_double * 2), ("kp_input_max", C.c_double), ("kp_input_min", C.c_double),
        ("solver_tolerance", C.c_double),
        ("reset_eef_bias", C.c_double * 3),
    ]


class PackedModel:
    """Keeps the numpy buffers alive next to the ctypes struct that points into them."""

    def __init__(self, model: UltrasoundModel):
        self.model = model
        self._keep: List[np.ndarray] = []
        s = UsimModel()
        p = model.params
        ids = model.ids
        s.nbody, s.nq, s.nv = model.nbody, model.nq, model.nv
        s.npart, s.npair, s.soft = int(ids[7]), len(model.eq_pairs), int(p.soft_torso)
        s.narm = len(p.link_pos)
        s.table_body, s.link1_body, s.hand_body, s.probe_body = int(ids[1]), int(ids[2]), int(ids[3]), int(ids[4])
        s.torso_body, s.part_body0 = int(ids[5]), int(ids[6])

        def dptr(name):
            a = np.ascontiguousarray(model.arrays[name], dtype=np.float64)
            self._keep.append(a)
            return a.ctypes.data_as(_pd)

        def iptr(name):
            a = np.ascontiguousarray(model.arrays[name], dtype=np.int32)
            self._keep.append(a)
            return a.ctypes.data_as(_pi)

        for f in ("body_parent", "body_jnt_type", "body_qposadr", "body_dofadr"):
            setattr(s, f, iptr("g_" + f))
        for f in ("body_pos", "body_quat", "body_mass", "body_ipos", "body_inertia", "body_jnt_axis"):
            setattr(s, f, dptr("g_" + f))
        s.dof_damping, s.jnt_range, s.ctrl_range, s.qpos0 = dptr("g_dof_damping"), dptr("g_jnt_range"), dptr("g_ctrl_range"), dptr("g_qpos0")
        for f in ("probe_seg", "part_pos", "part_axis", "part_seg_outer", "part_seg_inner", "dof_invweight0", "body_invweight0", "arm_link", "arm_tool"):
            setattr(s, f, dptr(f))
        s.eq_pairs, s.part_nbr = iptr("eq_pairs"), iptr("part_nbr")
        s.probe_radius, s.cap_radius = p.probe_radius, p.cap_radius
        s.tendon_invweight0 = float(model.tendon_invweight0[0])
        s.timestep, s.impratio = p.timestep, p.impratio
        s.gravity[:] = p.gravity
        s.solref[:] = p.solref
        s.solimp[:] = p.solimp
        s.solref_smooth[:] = p.solref_smooth
        s.table_top_z, s.table_half_xy, s.table_friction = p.table_top_z, p.table_half_xy, p.table_friction
        s.probe_friction, s.particle_friction = p.probe_friction, p.particle_friction
        s.init_qpos[:] = tuple(p.init_qpos) + (0.0,) * (7 - len(p.init_qpos))
        s.top_torso_offset, s.traj_x_range, s.traj_y_range = p.top_torso_offset, p.traj_x_range, p.traj_y_range
        self.struct = s


# systematic reset offset of the reference's toolbox IK, decoded from the shipped artifacts
ART_RESET_EEF_BIAS = (0.0028, 0.0008, 0.0066)


def _six(v):
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    return np.full(6, v[0]) if v.size == 1 else v[:6]


def make_config(
    num_envs: int,
    controller_configs: Dict[str, Any] | None = None,
    *,
    control_freq: float = 20,
    horizon: int = 1000,
    early_termination: bool = False,
    torso_solref_randomization: bool = False,
    initial_probe_pos_randomization: bool = False,
    deterministic_trajectory: bool = False,
    seed: int = 0,
    env_id_offset: int = 0,
    solver_iterations: int = 40,
    solver_tolerance: float = 3e-5,
    precond_rebuilds: int = 0,
    ignore_done: bool = False,
    reset_eef_bias=ART_RESET_EEF_BIAS,
) -> UsimConfig:
    """Translate the ``suite.make("Ultrasound", ...)`` kwargs (rl_config.yaml:18-57,
    defaults ultrasound.py:99-133) into the C config."""
    cc = dict(controller_configs or {})
    assert "OSC" in cc.get("type", "OSC_POSE"), "The robot controller must be of type OSC"
    c = UsimConfig()
    c.abi_version = USIM_ABI_VERSION
    c.num_envs, c.env_id_offset = int(num_envs), int(env_id_offset)
    c.impedance_mode = MODE_BY_NAME[cc.get("impedance_mode", "fixed")]
    c.horizon = int(horizon)
    c.early_termination = int(bool(early_termination))
    c.solref_randomization = int(bool(torso_solref_randomization))
    c.probe_pos_randomization = int(bool(initial_probe_pos_randomization))
    c.deterministic_trajectory = int(bool(deterministic_trajectory))
    c.uncouple_pos_ori = int(bool(cc.get("uncouple_pos_ori", True)))
    c.solver_iterations = int(solver_iterations)
    c.precond_rebuilds = int(precond_rebuilds)
    c.ignore_done = int(bool(ignore_done))
    c.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    c.control_freq = float(control_freq)
    c.kp[:] = _six(cc.get("kp", 150))
    c.damping_ratio[:] = _six(cc.get("damping_ratio", 1))
    c.input_max, c.input_min = float(cc.get("input_max", 1)), float(cc.get("input_min", -1))
    c.output_max[:] = _six(cc.get("output_max", [0.05, 0.05, 0.05, 0.5, 0.5, 0.5]))
    c.output_min[:] = _six(cc.get("output_min", [-0.05, -0.05, -0.05, -0.5, -0.5, -0.5]))
    c.kp_limits[:] = cc.get("kp_limits", [0, 300])
    c.kp_input_max, c.kp_input_min = float(cc.get("kp_input_max", 1)), float(cc.get("kp_input_min", 0))
    c.solver_tolerance = float(solver_tolerance)
    c.reset_eef_bias[:] = reset_eef_bias
    return c


def action_dim(cfg: UsimConfig) -> int:
    return 7 if cfg.impedance_mode == MODE_VARIABLE_Z else 6


def action_bounds(cfg: UsimConfig):
    """``env.action_spec``: (low, high).  Bounds per SURVEY §8(b) / [ART]."""
    m = cfg.impedance_mode
    if m == MODE_FIXED:
        return np.full(6, cfg.input_min), np.full(6, cfg.input_max)
    if m == MODE_TRACKING:
        return np.full(6, cfg.kp_input_min), np.full(6, cfg.kp_input_max)
    if m == MODE_VARIABLE_Z:
        return np.r_[np.full(6, cfg.kp_input_min), -1.0], np.r_[np.full(6, cfg.kp_input_max), 1.0]
    return np.full(6, -10.0), np.full(6, 10.0)
