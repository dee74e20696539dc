"""Model description of the Ultrasound scene (host side, float64 numpy).

This is the "MJCF compiler" of the B200 path: it turns the numbers of the
reference's scene into flat arrays that both the CUDA library (through the
C ABI, ``include/usim.h``) and the CPU oracle consume.  Nothing here runs per
step.

Provenance of every constant (paths relative to the reference checkout):

* table / arena ............ ``src/my_models/assets/arenas/ultrasound_arena.xml:22,31-32``
  and ``src/my_models/arenas/ultrasound_arena.py:21-23,55-58`` (half size
  (0.4,0.4,0.025), top at z=0.8, friction (1,0.005,0.0001)).
* probe gripper ............. ``src/my_models/assets/grippers/ultrasound_probe_gripper.xml:6-17``
  (body ``gripper_base`` at (-0.004,-0.063,0.128) in ``right_hand``, mass 1,
  friction (1e-4,0.005,1e-4), sites ``ft_frame``/``grip_site`` at the body
  origin).  The collision mesh is missing from the reference
  (``.MISSING_LARGE_BLOBS:1``); a capsule is substituted (assumption
  A-PROBE-1).
* soft torso ................ ``src/my_models/assets/objects/soft_box.xml:8-14``
  (composite box 9x4x11, spacing 0.035, capsule 0.0075/0.025, mass 0.01,
  friction (0.01,0.005,0.0001), contype 0, solrefsmooth (-1324.17,-17.59)).
* Panda arm, mount, world options: NOT in the reference tree (robosuite
  fork, unpinned).  Values follow SURVEY.md App. C.4/C.5 (recalled) and are
  all collected in :class:`SceneParams` so they can be flipped in one place.

The composite expansion follows the MuJoCo 2.0 "composite box" description
(SURVEY.md App. B.3): one radial slider per shell element, a soft "fix"
equality per slider, "smooth" equalities between grid-adjacent shell
elements carrying ``solrefsmooth``, and one fixed-tendon equality over all
sliders.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# small float64 helpers (wxyz quaternions, MuJoCo convention)
# ----------------------------------------------------------------------------


def quat2mat(q):
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def z2quat(axis):
    """Quaternion rotating the z axis onto ``axis`` (shortest arc)."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    z = np.array([0.0, 0.0, 1.0])
    c = np.cross(z, a)
    s = np.linalg.norm(c)
    if s < 1e-12:
        return np.array([1.0, 0, 0, 0]) if a[2] > 0 else np.array([0.0, 1.0, 0, 0])
    ang = np.arctan2(s, a[2])
    c = c / s
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * c])


def capsule_inertia(mass, radius, half_len):
    """Mass-normalised solid capsule inertia about its centre, long axis = z.

    Same split (cylinder + two hemispheres, uniform density) as MuJoCo's geom
    inertia [EXT-recall]."""
    h = 2.0 * half_len
    vc = np.pi * radius * radius * h
    vs = 4.0 / 3.0 * np.pi * radius**3
    mc = mass * vc / (vc + vs)
    ms = mass * vs / (vc + vs)
    izz = mc * radius * radius / 2.0 + ms * 2.0 * radius * radius / 5.0
    ixx = (
        mc * (3 * radius * radius + h * h) / 12.0
        + ms * 2.0 * radius * radius / 5.0
        + ms * h * (3 * radius + 2 * h) / 8.0
    )
    return np.diag([ixx, ixx, izz])


# ----------------------------------------------------------------------------
# scene parameters
# ----------------------------------------------------------------------------

JNT_NONE, JNT_HINGE, JNT_SLIDE, JNT_FREE = 0, 1, 2, 3
GEOM_TABLE, GEOM_CAPSULE = 0, 1
EQ_FIX, EQ_SMOOTH, EQ_TENDON = 0, 1, 2

# compact geom ids used for contact reporting (usim_get_contacts)
GEOM_ID_FLOOR = 0
GEOM_ID_TABLE = 1
GEOM_ID_PROBE = 2
GEOM_ID_TORSO_CENTER = 3
GEOM_ID_PARTICLE0 = 4


@dataclass
class SceneParams:
    """Every tunable of the scene.  [EXT-recall] items are assumptions."""

    # ---- global options (robosuite base.xml / MuJoCo 2.0 defaults) [EXT-recall A-OPT-1]
    timestep: float = 0.002
    gravity: Tuple[float, float, float] = (0.0, 0.0, -9.81)
    impratio: float = 20.0
    solref: Tuple[float, float] = (0.02, 1.0)
    solimp: Tuple[float, float, float, float, float] = (0.9, 0.95, 0.001, 0.5, 2.0)

    # ---- table (in tree)
    table_top_z: float = 0.8
    table_half_xy: float = 0.4
    table_friction: float = 1.0

    # ---- arm: Panda [EXT-recall A-PANDA-1] by default; `ur5e_params()` swaps in the UR5e chain (6 joints).  One tuple entry per joint.
    robot: str = "Panda"
    link_axis: Tuple[Tuple[float, float, float], ...] | None = None  # joint axes in the link frames (None: all local z)
    link_inertia_quat: Tuple[Tuple[float, float, float, float], ...] | None = None  # orientation of the principal axes (None: link frame)
    link_diaginertia3: Tuple[Tuple[float, float, float], ...] | None = None  # full principal inertias (None: `link_diaginertia` x identity)
    base_pos: Tuple[float, float, float] = (-0.56, 0.0, 0.913)
    link_pos: Tuple[Tuple[float, float, float], ...] = (
        (0, 0, 0.333),
        (0, 0, 0),
        (0, -0.316, 0),
        (0.0825, 0, 0),
        (-0.0825, 0.384, 0),
        (0, 0, 0),
        (0.088, 0, 0),
    )
    link_quat: Tuple[Tuple[float, float, float, float], ...] = (
        (1, 0, 0, 0),
        (0.7071067811865476, -0.7071067811865476, 0, 0),
        (0.7071067811865476, 0.7071067811865476, 0, 0),
        (0.7071067811865476, 0.7071067811865476, 0, 0),
        (0.7071067811865476, -0.7071067811865476, 0, 0),
        (0.7071067811865476, 0.7071067811865476, 0, 0),
        (0.7071067811865476, 0.7071067811865476, 0, 0),
    )
    link_mass: Tuple[float, ...] = (3.0, 3.0, 2.0, 2.0, 2.0, 1.5, 0.5)
    link_diaginertia: Tuple[float, ...] = (0.3, 0.3, 0.2, 0.2, 0.2, 0.1, 0.05)
    link_com: Tuple[Tuple[float, float, float], ...] = (
        (0, 0, -0.07),
        (0, -0.1, 0),
        (0.04, 0, -0.05),
        (-0.04, 0.05, 0),
        (0, 0, -0.15),
        (0.06, 0, 0),
        (0, 0, 0.08),
    )
    joint_range: Tuple[Tuple[float, float], ...] = (
        (-2.8973, 2.8973),
        (-1.7628, 1.7628),
        (-2.8973, 2.8973),
        (-3.0718, -0.0698),
        (-2.8973, 2.8973),
        (-0.0175, 3.7525),
        (-2.8973, 2.8973),
    )
    joint_damping: float = 0.1
    ctrl_range: Tuple[float, ...] = (80.0, 80.0, 80.0, 80.0, 12.0, 12.0, 12.0)
    init_qpos: Tuple[float, ...] = (
        0.0,
        np.pi / 16.0,
        0.0,
        -np.pi / 2.0 - np.pi / 3.0,
        0.0,
        np.pi - 0.2,
        np.pi / 4,
    )
    hand_pos: Tuple[float, float, float] = (0.0, 0.0, 0.1065)
    hand_quat: Tuple[float, float, float, float] = (0.9238795325112867, 0.0, 0.0, -0.3826834323650898)
    hand_mass: float = 0.5
    hand_diaginertia: float = 0.05

    # ---- probe (in tree: pose, mass, friction; shape substituted A-PROBE-1)
    probe_pos: Tuple[float, float, float] = (-0.004, -0.063, 0.128)
    probe_mass: float = 1.0
    probe_friction: float = 1e-4
    # The reference's collision mesh is missing (A-PROBE-1): one capsule is substituted, CALIBRATED against the only reference-produced
    # force / torque numbers there are -- the 192 raw post-reset observation rows of the shipped VecNormalize pickles
    # (scripts/probe_calibrate.py, Nelder-Mead on 10 parameters over float64 oracle resets; profiles/r02_probe_fit2.json,
    # profiles/r02_probe_calibration.md).  Result: a bar across the probe body's x axis (the wide, thin, rounded face of a convex-array
    # probe, docs/images/press_torso_transparent.png), 5.8 cm between the end-sphere centres, radius 2.1 cm, lowest point 2 mm below the
    # grip site, centre of mass 5 cm up the housing and ~8 mm off axis (from the F/T torque of the rows without contact).
    # Round 1 shipped an axial capsule with a 5 cm tip sphere (probe_tip_z / probe_back_z below, still available by passing
    # probe_seg_a = probe_seg_b = None): calibration loss 444 vs 40 for the bar; lateral force ratio 0.16 vs 0.29 (reference 0.38),
    # z-torque spread 0.004 vs 0.18 N m (0.33), F/T torque without contact (0.075,-0.005,0) vs (0.083,-0.029,-0.006) ((0.091,-0.033,-0.007)).
    probe_radius: float = 0.021
    probe_tip_z: float = -0.05  # axial form: tip-sphere centre (with probe_radius 0.05: tip surface at the grip_site origin)
    probe_back_z: float = -0.10
    # general form of the substituted capsule (probe-body frame, metres; body +z points out of the probe face, i.e. down at the goal
    # orientation): segment end points and centre of mass; None = the axial capsule ((0,0,probe_tip_z) .. (0,0,probe_back_z), COM at
    # the middle of the segment)
    probe_seg_a: Tuple[float, float, float] | None = (0.030, -0.0015, -0.019)
    probe_seg_b: Tuple[float, float, float] | None = (-0.028, -0.0005, -0.019)
    probe_com: Tuple[float, float, float] | None = (-0.006, -0.006, -0.052)

    # ---- soft composite (in tree): soft_box.xml (type box) or soft_human_torso.xml (type cylinder, `use_box_torso=False`)
    comp_type: str = "box"
    top_torso_offset: float = 0.039  # ultrasound.py:184 (0.041 for the cylinder)
    traj_x_range: float = 0.15       # ultrasound.py:185
    traj_y_range: float = 0.09       # ultrasound.py:186 (0.05 for the cylinder)
    comp_count: Tuple[int, int, int] = (9, 4, 11)
    comp_spacing: float = 0.035
    cap_radius: float = 0.0075
    cap_half_len: float = 0.025
    particle_mass: float = 0.01
    particle_friction: float = 0.01
    solref_smooth: Tuple[float, float] = (-1324.17, -17.59)
    torso_quat: Tuple[float, float, float, float] = (0.5, 0.5, -0.5, -0.5)
    # centre (0,0,0.8) + z_offset 0.005 + bottom-site offset 0.0522 (ultrasound.py:304-314, soft_box.xml:14)
    torso_pos: Tuple[float, float, float] = (0.0, 0.0, 0.8 + 0.005 + 0.0522)
    free_joint_damping: float = 0.0005  # [EXT-recall] robosuite object free joints

    # ---- scene switch
    soft_torso: bool = True


@dataclass
class UltrasoundModel:
    """Flat float64/int32 arrays.  ``g_*`` = generic tree view (oracle),
    ``arm_*``/``ts_*`` = specialised tables (CUDA kernels)."""

    params: SceneParams
    nq: int
    nv: int
    nbody: int
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    particle_names: List[str] = field(default_factory=list)

    def __getattr__(self, name):
        arrays = object.__getattribute__(self, "arrays")
        if name in arrays:
            return arrays[name]
        raise AttributeError(name)

    def geom_name(self, gid: int) -> str:
        if gid == GEOM_ID_FLOOR:
            return "floor"
        if gid == GEOM_ID_TABLE:
            return "table_collision"
        if gid == GEOM_ID_PROBE:
            return "gripper0_probe_collision"
        if gid == GEOM_ID_TORSO_CENTER:
            return "torso_Gcenter"
        return "torso_" + self.particle_names[gid - GEOM_ID_PARTICLE0]


def _composite_box(p: SceneParams):
    """Shell elements of the composite box in MuJoCo order (ix outer, iz inner)."""
    cx, cy, cz = p.comp_count
    pos, names, index = [], [], {}
    for ix in range(cx):
        for iy in range(cy):
            for iz in range(cz):
                if ix in (0, cx - 1) or iy in (0, cy - 1) or iz in (0, cz - 1):
                    index[(ix, iy, iz)] = len(pos)
                    e = p.comp_spacing * np.array([ix - 0.5 * (cx - 1), iy - 0.5 * (cy - 1), iz - 0.5 * (cz - 1)])
                    if p.comp_type == "cylinder":  # MuJoCo BoxProject [EXT-recall]: rescale (x, y) from the L-inf to the L2 ball
                        l0, l2 = max(abs(e[0]), abs(e[1])), np.hypot(e[0], e[1])
                        if l2 > 0:
                            e[:2] *= l0 / l2
                    elif p.comp_type != "box":
                        raise ValueError(f"composite type {p.comp_type!r} not supported")
                    pos.append(e)
                    names.append(f"G{ix}_{iy}_{iz}")
    pairs = []
    for (ix, iy, iz), a in index.items():
        for d in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
            nb = (ix + d[0], iy + d[1], iz + d[2])
            if nb in index:
                pairs.append((a, index[nb]))
    pairs.sort()
    return np.array(pos), names, np.array(pairs, dtype=np.int32)


def probe_geometry(p: SceneParams):
    """(segment end points [2,3], COM [3], inertia about the COM [3,3]) of the probe capsule in the probe-body frame."""
    a = np.asarray(p.probe_seg_a if p.probe_seg_a is not None else (0.0, 0.0, p.probe_tip_z), float)
    b = np.asarray(p.probe_seg_b if p.probe_seg_b is not None else (0.0, 0.0, p.probe_back_z), float)
    com = np.asarray(p.probe_com, float) if p.probe_com is not None else 0.5 * (a + b)
    hl = 0.5 * np.linalg.norm(b - a)
    Rc = quat2mat(z2quat(b - a)) if hl > 1e-12 else np.eye(3)
    return np.array([a, b]), com, Rc @ capsule_inertia(p.probe_mass, p.probe_radius, hl) @ Rc.T


def ur5e_params(base: SceneParams | None = None, **kw) -> SceneParams:
    """`robots="UR5e"` (ultrasound.py:137,833-839,863-864): the robosuite UR5e chain [EXT-recall, assumption A-UR5E-1: robosuite
    `models/assets/robots/ur5e/robot.xml` -- link offsets, joint axes, inertials, +-150 / +-28 Nm actuators, `init_qpos`, same
    mount and table offset as the Panda].  Six joints: the seventh arm slot of the state is an inert degree of freedom."""
    s2 = 0.7071067811865476
    return dataclasses.replace(
        base or SceneParams(), robot="UR5e",
        link_pos=((0, 0, 0.163), (0, 0.138, 0), (0, -0.131, 0.425), (0, 0, 0.392), (0, 0.127, 0), (0, 0, 0.1)),
        link_quat=((1, 0, 0, 0), (s2, 0, s2, 0), (1, 0, 0, 0), (s2, 0, s2, 0), (1, 0, 0, 0), (1, 0, 0, 0)),
        link_axis=((0, 0, 1), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 0, 1), (0, 1, 0)),
        link_mass=(3.7, 8.393, 2.33, 1.219, 1.219, 0.1889),
        link_diaginertia=(0.0, 0.0, 0.0, 0.0, 0.0, 0.0),
        link_diaginertia3=((0.0102675, 0.0102675, 0.00666), (0.133886, 0.133886, 0.0151074), (0.0312168, 0.0312168, 0.004095),
                           (0.0025599, 0.0025599, 0.0021942), (0.0025599, 0.0025599, 0.0021942), (0.000132134, 9.90863e-05, 9.90863e-05)),
        link_inertia_quat=((1, 0, 0, 0),) * 5 + ((s2, 0, 0, s2),),
        link_com=((0, 0, 0), (0, 0, 0.2125), (0, 0, 0.196), (0, 0.127, 0), (0, 0, 0.1), (0, 0.0771683, 0)),
        joint_range=((-6.28319, 6.28319), (-6.28319, 6.28319), (-3.14159, 3.14159), (-6.28319, 6.28319), (-6.28319, 6.28319), (-6.28319, 6.28319)),
        joint_damping=0.001, ctrl_range=(150.0, 150.0, 150.0, 28.0, 28.0, 28.0),
        init_qpos=(-0.470, -1.735, 2.480, -2.275, -1.590, -1.991),
        hand_pos=(0.0, 0.098, 0.0), hand_quat=(s2, -s2, 0.0, 0.0), **kw)


def _link_inertia(p: SceneParams, j: int):
    """inertia of arm link j about its COM, in the link frame"""
    if p.link_diaginertia3 is None:
        return np.eye(3) * p.link_diaginertia[j]
    Rq = quat2mat(p.link_inertia_quat[j]) if p.link_inertia_quat is not None else np.eye(3)
    return Rq @ np.diag(p.link_diaginertia3[j]) @ Rq.T


def _link_axis(p: SceneParams, j: int):
    return np.asarray(p.link_axis[j] if p.link_axis is not None else (0.0, 0.0, 1.0), float)


def cylinder_torso_params(**kw) -> SceneParams:
    """`use_box_torso=False`: soft_human_torso.xml:8-14 (composite cylinder, bottom site at -0.05) and ultrasound.py:184-186."""
    return SceneParams(comp_type="cylinder", top_torso_offset=0.041, traj_y_range=0.05, torso_pos=(0.0, 0.0, 0.8 + 0.005 + 0.05), **kw)


def build_model(params: SceneParams | None = None) -> UltrasoundModel:
    p = params or SceneParams()
    A: Dict[str, np.ndarray] = {}

    # ------------------------------------------------------------------ bodies
    parent, bpos, bquat, bmass, bipos, binert = [], [], [], [], [], []
    jtype, jaxis, qadr, dadr = [], [], [], []

    def add_body(par, pos, quat, mass, ipos, inertia, jt=JNT_NONE, axis=(0, 0, 1)):
        parent.append(par)
        bpos.append(np.asarray(pos, float))
        bquat.append(np.asarray(quat, float))
        bmass.append(float(mass))
        bipos.append(np.asarray(ipos, float))
        binert.append(np.asarray(inertia, float).reshape(3, 3))
        jtype.append(jt)
        jaxis.append(np.asarray(axis, float))
        return len(parent) - 1

    world = add_body(-1, (0, 0, 0), (1, 0, 0, 0), 0, (0, 0, 0), np.zeros((3, 3)))
    table = add_body(world, (0, 0, p.table_top_z - 0.025), (1, 0, 0, 0), 0, (0, 0, 0), np.zeros((3, 3)))
    prev = world
    link_ids = []
    nj = len(p.link_pos)
    assert nj in (6, 7), "the arm kernels are compiled for 6 or 7 joints"
    for j in range(nj):
        pos = np.asarray(p.link_pos[j], float)
        if j == 0:
            pos = pos + np.asarray(p.base_pos, float)
        b = add_body(prev, pos, p.link_quat[j], p.link_mass[j], p.link_com[j], _link_inertia(p, j), JNT_HINGE, _link_axis(p, j))
        link_ids.append(b)
        prev = b
    for j in range(nj, 7):
        # inert degree of freedom filling the arm's seventh state slot: a unit-inertia rotor on the world, coupled to nothing
        # (M = 1 on its diagonal, no gravity torque, no Jacobian column, no actuator)
        add_body(world, (0, 0, -10.0), (1, 0, 0, 0), 1.0, (0, 0, 0), np.eye(3), JNT_HINGE, (0, 0, 1))
    hand = add_body(prev, p.hand_pos, p.hand_quat, p.hand_mass, (0, 0, 0), np.eye(3) * p.hand_diaginertia)
    probe_seg, probe_c, probe_I = probe_geometry(p)
    probe = add_body(hand, p.probe_pos, (1, 0, 0, 0), p.probe_mass, probe_c, probe_I)

    npart = 0
    if p.soft_torso:
        ppos, names, pairs = _composite_box(p)
        npart = len(ppos)
        # centre geom: element geom scaled x2, same mass attribute (A-COMP-3)
        torso = add_body(
            world,
            p.torso_pos,
            p.torso_quat,
            p.particle_mass,
            (0, 0, 0),
            capsule_inertia(p.particle_mass, 2 * p.cap_radius, 2 * p.cap_half_len),
            JNT_FREE,
        )
        paxis = ppos / np.linalg.norm(ppos, axis=1, keepdims=True)
        off = p.cap_radius + p.cap_half_len
        for k in range(npart):
            Rg = quat2mat(z2quat(paxis[k]))
            Icap = capsule_inertia(p.particle_mass, p.cap_radius, p.cap_half_len)
            add_body(
                torso,
                ppos[k],
                (1, 0, 0, 0),
                p.particle_mass,
                -off * paxis[k],
                Rg @ Icap @ Rg.T,
                JNT_SLIDE,
                paxis[k],
            )
    else:
        names, pairs, ppos, paxis = [], np.zeros((0, 2), np.int32), np.zeros((0, 3)), np.zeros((0, 3))
        torso = -1

    nbody = len(parent)
    nq = nv = 0
    for b in range(nbody):
        qadr.append(nq)
        dadr.append(nv)
        if jtype[b] in (JNT_HINGE, JNT_SLIDE):
            nq += 1
            nv += 1
        elif jtype[b] == JNT_FREE:
            nq += 7
            nv += 6

    A["g_body_parent"] = np.array(parent, np.int32)
    A["g_body_pos"] = np.array(bpos)
    A["g_body_quat"] = np.array(bquat)
    A["g_body_mass"] = np.array(bmass)
    A["g_body_ipos"] = np.array(bipos)
    A["g_body_inertia"] = np.array(binert).reshape(nbody, 9)
    A["g_body_jnt_type"] = np.array(jtype, np.int32)
    A["g_body_jnt_axis"] = np.array(jaxis)
    A["g_body_qposadr"] = np.array(qadr, np.int32)
    A["g_body_dofadr"] = np.array(dadr, np.int32)

    damping = np.zeros(nv)
    damping[:nj] = p.joint_damping
    if p.soft_torso:
        damping[7:13] = p.free_joint_damping
    A["g_dof_damping"] = damping
    A["g_jnt_range"] = np.array(tuple(p.joint_range) + ((-1e9, 1e9),) * (7 - nj), float)  # arm only (7,2); inert slot: unlimited
    A["g_ctrl_range"] = np.array(tuple(p.ctrl_range) + (0.0,) * (7 - nj), float)

    qpos0 = np.zeros(nq)
    if p.soft_torso:
        qpos0[7:10] = p.torso_pos
        qpos0[10:14] = p.torso_quat
    A["g_qpos0"] = qpos0

    # ------------------------------------------------------------------ ids
    A["ids"] = np.array(
        [world, table, link_ids[0], hand, probe, torso, (torso + 1) if p.soft_torso else -1, npart],
        np.int32,
    )

    # ------------------------------------------------------------------ geoms
    # capsule geoms: body-frame segment endpoints + radius
    A["probe_seg"] = probe_seg
    A["probe_radius"] = np.array([p.probe_radius])
    if p.soft_torso:
        off = p.cap_radius + p.cap_half_len
        # in particle-body frame: outer end (sphere centre) and inner end
        A["part_seg_outer"] = -(off - p.cap_half_len) * paxis  # = -radius*axis
        A["part_seg_inner"] = -(off + p.cap_half_len) * paxis
    else:
        A["part_seg_outer"] = np.zeros((0, 3))
        A["part_seg_inner"] = np.zeros((0, 3))
    A["part_pos"] = ppos
    A["part_axis"] = paxis
    A["eq_pairs"] = pairs

    # neighbour table per particle (up to 6), -1 padded, plus index of the pair
    nbr = -np.ones((npart, 6), np.int32)
    nbr_pair = -np.ones((npart, 6), np.int32)
    cnt = np.zeros(npart, np.int32)
    for e, (a, b) in enumerate(pairs):
        nbr[a, cnt[a]] = b
        nbr_pair[a, cnt[a]] = e
        cnt[a] += 1
        nbr[b, cnt[b]] = a
        nbr_pair[b, cnt[b]] = e
        cnt[b] += 1
    A["part_nbr"] = nbr
    A["part_nbr_pair"] = nbr_pair

    model = UltrasoundModel(params=p, nq=nq, nv=nv, nbody=nbody, arrays=A, particle_names=names)

    # ------------------------------------------------------------------ invweights at qpos0
    _compute_invweight0(model)
    _arm_tables(model)
    return model


# ----------------------------------------------------------------------------
# dense kinematics/inertia at a configuration (numpy; used for invweight0 only)
# ----------------------------------------------------------------------------


def forward_kinematics(model: UltrasoundModel, qpos):
    """World pose of every body: (xpos[nbody,3], xmat[nbody,3,3])."""
    A = model.arrays
    nb = model.nbody
    xpos = np.zeros((nb, 3))
    xmat = np.zeros((nb, 3, 3))
    xmat[0] = np.eye(3)
    for b in range(1, nb):
        par = A["g_body_parent"][b]
        jt = A["g_body_jnt_type"][b]
        qa = A["g_body_qposadr"][b]
        if jt == JNT_FREE:
            xpos[b] = qpos[qa : qa + 3]
            xmat[b] = quat2mat(qpos[qa + 3 : qa + 7])
            continue
        pos = A["g_body_pos"][b].copy()
        R = quat2mat(A["g_body_quat"][b])
        if jt == JNT_SLIDE:
            pos = pos + A["g_body_jnt_axis"][b] * qpos[qa]
        elif jt == JNT_HINGE:
            ax = A["g_body_jnt_axis"][b]
            ang = qpos[qa]
            qj = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
            R = R @ quat2mat(qj)
        xpos[b] = xpos[par] + xmat[par] @ pos
        xmat[b] = xmat[par] @ R
    return xpos, xmat


def body_jacobian(model: UltrasoundModel, xpos, xmat, body, point):
    """Dense (Jp[3,nv], Jr[3,nv]) of world ``point`` attached to ``body``."""
    A = model.arrays
    Jp = np.zeros((3, model.nv))
    Jr = np.zeros((3, model.nv))
    b = body
    while b > 0:
        jt = A["g_body_jnt_type"][b]
        d = A["g_body_dofadr"][b]
        if jt == JNT_HINGE:
            ax = xmat[b] @ A["g_body_jnt_axis"][b]
            Jr[:, d] = ax
            Jp[:, d] = np.cross(ax, point - xpos[b])
        elif jt == JNT_SLIDE:
            # axis is fixed in the body frame (body does not rotate relative to its parent)
            Jp[:, d] = xmat[b] @ A["g_body_jnt_axis"][b]
        elif jt == JNT_FREE:
            Jp[:, d : d + 3] = np.eye(3)
            for k in range(3):
                ax = xmat[b][:, k]
                Jr[:, d + 3 + k] = ax
                Jp[:, d + 3 + k] = np.cross(ax, point - xpos[b])
        b = A["g_body_parent"][b]
    return Jp, Jr


def mass_matrix(model: UltrasoundModel, qpos):
    A = model.arrays
    xpos, xmat = forward_kinematics(model, qpos)
    M = np.zeros((model.nv, model.nv))
    for b in range(1, model.nbody):
        m = A["g_body_mass"][b]
        if m <= 0:
            continue
        com = xpos[b] + xmat[b] @ A["g_body_ipos"][b]
        Iw = xmat[b] @ A["g_body_inertia"][b].reshape(3, 3) @ xmat[b].T
        Jp, Jr = body_jacobian(model, xpos, xmat, b, com)
        nz = np.nonzero(np.any(Jp != 0, axis=0) | np.any(Jr != 0, axis=0))[0]
        Jp, Jr = Jp[:, nz], Jr[:, nz]
        M[np.ix_(nz, nz)] += m * Jp.T @ Jp + Jr.T @ Iw @ Jr
    return M, xpos, xmat


def _compute_invweight0(model: UltrasoundModel):
    """dof_invweight0 / body_invweight0 / tendon_invweight0 at qpos0
    (MuJoCo ``setInertia/set0`` semantics [EXT-recall, SURVEY App. C.5])."""
    A = model.arrays
    M, xpos, xmat = mass_matrix(model, A["g_qpos0"])
    Minv = np.linalg.inv(M)
    nv = model.nv
    dof_iw = np.diag(Minv).copy()
    if model.params.soft_torso:
        # free joint: average over the translational / rotational triplets
        dof_iw[7:10] = dof_iw[7:10].mean()
        dof_iw[10:13] = dof_iw[10:13].mean()
    body_iw = np.zeros((model.nbody, 2))
    for b in range(1, model.nbody):
        if A["g_body_mass"][b] <= 0 and A["g_body_jnt_type"][b] == JNT_NONE and A["g_body_parent"][b] == 0:
            continue  # static
        com = xpos[b] + xmat[b] @ A["g_body_ipos"][b]
        Jp, Jr = body_jacobian(model, xpos, xmat, b, com)
        body_iw[b, 0] = np.trace(Jp @ Minv @ Jp.T) / 3.0
        body_iw[b, 1] = np.trace(Jr @ Minv @ Jr.T) / 3.0
    A["dof_invweight0"] = dof_iw
    A["body_invweight0"] = body_iw
    if model.params.soft_torso:
        jt = np.zeros(nv)
        jt[13:] = 1.0
        A["tendon_invweight0"] = np.array([jt @ Minv @ jt])
    else:
        A["tendon_invweight0"] = np.array([0.0])


def _arm_tables(model: UltrasoundModel):
    """Specialised arm tables for the CUDA arm kernel.

    ``arm_link`` [7,22]: pos(3) R(9, row major) com(3) mass(1) inertia about
    COM in link frame (xx,yy,zz,xy,xz,yz).  Link 7 carries the hand and the
    probe welded to it (composite rigid body).
    ``arm_tool`` [33]: in link-7 frame: grip-site pos(3) R(9); hand origin(3);
    probe tip centre(3), probe back centre(3), probe radius(1); probe body
    mass(1) com(3) inertia(6) ; +1 pad.
    """
    p = model.params
    A = model.arrays
    nj = len(p.link_pos)
    link = np.zeros((7, 22))
    link[:, 3:12] = np.eye(3).reshape(9)
    # The kernels rotate every joint about its frame's z axis.  A joint about another axis a_j is brought to that convention by
    # re-orienting its frame: F'_j = F_j A_j with A_j z = a_j; everything attached to link j (child offset / orientation, COM,
    # inertia, tool) is then expressed in the rotated frame, A_j^T (.).
    Aprev = np.eye(3)
    Acur = []
    for j in range(nj):
        pos = np.asarray(p.link_pos[j], float)
        if j == 0:
            pos = pos + np.asarray(p.base_pos, float)
        Aj = quat2mat(z2quat(_link_axis(p, j)))
        link[j, 0:3] = Aprev.T @ pos
        link[j, 3:12] = (Aprev.T @ quat2mat(p.link_quat[j]) @ Aj).reshape(9)
        link[j, 12:15] = Aj.T @ np.asarray(p.link_com[j], float)
        link[j, 15] = p.link_mass[j]
        I = Aj.T @ _link_inertia(p, j) @ Aj
        link[j, 16:22] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
        Acur.append(Aj)
        Aprev = Aj
    AL = Acur[-1]
    # weld hand + probe into the last link (expressed in its rotated frame)
    Rh = AL.T @ quat2mat(p.hand_quat)
    ph = AL.T @ np.asarray(p.hand_pos, float)
    pp = ph + Rh @ np.asarray(p.probe_pos, float)  # probe body origin in the last link's frame
    _, probe_c_local, probe_I = probe_geometry(p)
    Iprobe = Rh @ probe_I @ Rh.T
    L = nj - 1
    Il = np.array([[link[L, 16], link[L, 19], link[L, 20]], [link[L, 19], link[L, 17], link[L, 21]], [link[L, 20], link[L, 21], link[L, 18]]])
    parts = [
        (p.link_mass[L], link[L, 12:15].copy(), Il),
        (p.hand_mass, ph, np.eye(3) * p.hand_diaginertia),
        (p.probe_mass, pp + Rh @ probe_c_local, Iprobe),
    ]
    mt = sum(m for m, _, _ in parts)
    ct = sum(m * c for m, c, _ in parts) / mt
    It = np.zeros((3, 3))
    for m, c, I in parts:
        d = c - ct
        It += I + m * (d @ d * np.eye(3) - np.outer(d, d))
    link[L, 12:15] = ct
    link[L, 15] = mt
    link[L, 16:22] = [It[0, 0], It[1, 1], It[2, 2], It[0, 1], It[0, 2], It[1, 2]]
    A["arm_link"] = link

    tool = np.zeros(34)
    tool[0:3] = pp  # grip_site == ft_frame == probe body origin
    tool[3:12] = Rh.reshape(9)
    tool[12:15] = ph
    tool[15:18] = pp + Rh @ A["probe_seg"][0]
    tool[18:21] = pp + Rh @ A["probe_seg"][1]
    tool[21] = p.probe_radius
    tool[22] = p.probe_mass
    tool[23:26] = pp + Rh @ probe_c_local
    tool[26:32] = [Iprobe[0, 0], Iprobe[1, 1], Iprobe[2, 2], Iprobe[0, 1], Iprobe[0, 2], Iprobe[1, 2]]
    A["arm_tool"] = tool
