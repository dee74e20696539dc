"""MJCF-subset reader for the scene files the reference keeps in tree.

Reads exactly what the hot path needs from

* ``src/my_models/assets/objects/soft_box.xml`` / ``soft_human_torso.xml`` -- ``<composite>``: type, count, spacing,
  ``solrefsmooth``; its ``<geom>``: capsule size, mass, friction, contype; the ``bottom_site`` / ``top_site`` offsets;
* ``src/my_models/assets/arenas/ultrasound_arena.xml`` -- the collision geoms (``floor``, ``table_collision``): every other geom of
  the file is ``contype=0 conaffinity=0`` (checked) -- together with the table placement rule of
  ``src/my_models/arenas/ultrasound_arena.py:21-23,55-58`` (half size = ``table_full_size / 2``, centre = offset - half height);
* ``src/my_models/assets/grippers/ultrasound_probe_gripper.xml`` -- body pose of ``gripper_base``, the collision geom's mass and
  friction, the sites and the force / torque sensors at ``ft_frame``.

and turns them into the corresponding :class:`~.model.SceneParams` fields, so that the numbers the kernels are compiled against
can be *checked against the reference's files* instead of being transcribed by hand (``tests/golden/make_golden.py`` stores the
parse of the reference checkout in ``tests/golden/mjcf_golden.json``; ``tests/test_task_golden.py`` asserts ``SceneParams`` against it).

Only the standard library is used.  This is host-side set-up code; nothing here runs per step.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from typing import Any, Dict, Optional, Tuple

from .model import SceneParams


def _floats(s: Optional[str]) -> Tuple[float, ...]:
    return tuple(float(x) for x in s.split()) if s else ()


def read_composite(path: str) -> Dict[str, Any]:
    """``<composite>`` object file (soft_box.xml:8-14, soft_human_torso.xml:8-14)."""
    root = ET.parse(path).getroot()
    obj = root.find(".//body[@name='object']")
    comp = obj.find("composite")
    geom = comp.find("geom")
    sites = {s.get("name"): _floats(s.get("pos")) for s in root.iter("site")}
    size = _floats(geom.get("size"))
    return dict(
        model=root.get("model"), comp_type=comp.get("type"), comp_count=tuple(int(x) for x in comp.get("count").split()),
        comp_spacing=float(comp.get("spacing")), solref_smooth=_floats(comp.get("solrefsmooth")),
        geom_type=geom.get("type"), cap_radius=size[0], cap_half_len=size[1], particle_mass=float(geom.get("mass")),
        particle_friction=_floats(geom.get("friction")), particle_contype=int(geom.get("contype", "1")),
        particle_conaffinity=int(geom.get("conaffinity", "1")), object_quat=_floats(obj.get("quat")),
        bottom_site=sites.get("bottom_site"), top_site=sites.get("top_site"), has_skin=comp.find("skin") is not None,
    )


def read_arena(path: str, table_full_size=(0.8, 0.8, 0.05), table_friction=(1.0, 0.005, 0.0001), table_offset=(0.0, 0.0, 0.8)) -> Dict[str, Any]:
    """Arena file + the placement rule of ``UltrasoundArena.configure_location`` (ultrasound_arena.py:21-23,55-58)."""
    root = ET.parse(path).getroot()
    wb = root.find("worldbody")
    colliding, visual = [], []
    for g in wb.iter("geom"):
        (visual if g.get("contype") == "0" and g.get("conaffinity") == "0" else colliding).append(g.get("name"))
    floor = wb.find("./geom[@name='floor']")
    table = wb.find("./body[@name='table']/geom[@name='table_collision']")
    half = tuple(0.5 * x for x in table_full_size)
    centre = (table_offset[0], table_offset[1], table_offset[2] - half[2])
    return dict(
        colliding_geoms=colliding, visual_geoms=visual, floor_type=floor.get("type"), floor_condim=int(floor.get("condim", "3")),
        table_type=table.get("type"), table_friction_xml=_floats(table.get("friction")), table_friction=tuple(table_friction),
        table_half_size=half, table_centre=centre, table_top_z=centre[2] + half[2],
        cameras=[c.get("name") for c in wb.iter("camera")],
    )


def read_gripper(path: str) -> Dict[str, Any]:
    """Probe gripper (ultrasound_probe_gripper.xml:3-18)."""
    root = ET.parse(path).getroot()
    body = root.find(".//body[@name='gripper_base']")
    col = body.find("./geom[@name='probe_collision']")
    mesh = root.find("./asset/mesh")
    sensors = root.find("sensor")
    return dict(
        body_pos=_floats(body.get("pos")), body_quat=_floats(body.get("quat")), collision_type=col.get("type"),
        collision_mesh=col.get("mesh"), mesh_file=mesh.get("file"), mesh_scale=_floats(mesh.get("scale")),
        probe_mass=float(col.get("mass")), probe_friction=_floats(col.get("friction")),
        sites={s.get("name"): _floats(s.get("pos")) or (0.0, 0.0, 0.0) for s in body.iter("site")},
        sensors={s.tag: s.get("site") for s in (sensors if sensors is not None else [])},
        contact_geoms=[g.get("name") for g in body.iter("geom") if g.get("group") == "0"],
    )


def read_assets(models_dir: str, use_box_torso: bool = True) -> Dict[str, Any]:
    """Parse the three scene files under ``<reference>/src/my_models``."""
    a = os.path.join(models_dir, "assets")
    return dict(
        composite=read_composite(os.path.join(a, "objects", "soft_box.xml" if use_box_torso else "soft_human_torso.xml")),
        arena=read_arena(os.path.join(a, "arenas", "ultrasound_arena.xml")),
        gripper=read_gripper(os.path.join(a, "grippers", "ultrasound_probe_gripper.xml")),
    )


def scene_fields(parsed: Dict[str, Any]) -> Dict[str, Any]:
    """The :class:`SceneParams` fields the parsed files determine (everything else in SceneParams is recalled third-party data)."""
    c, ar, g = parsed["composite"], parsed["arena"], parsed["gripper"]
    assert c["geom_type"] == "capsule" and c["particle_contype"] == 0, "the kernels are built for capsule particles that do not collide with each other"
    assert list(ar["colliding_geoms"]) == ["floor", "table_collision"], ar["colliding_geoms"]
    assert g["sensors"] == {"force": "ft_frame", "torque": "ft_frame"} and tuple(g["sites"]["ft_frame"]) == (0.0, 0.0, 0.0)
    z_offset = 0.005  # placement sampler z_offset (ultrasound.py:304-314)
    return dict(
        comp_type=c["comp_type"], comp_count=tuple(c["comp_count"]), comp_spacing=c["comp_spacing"], cap_radius=c["cap_radius"],
        cap_half_len=c["cap_half_len"], particle_mass=c["particle_mass"], particle_friction=c["particle_friction"][0],
        solref_smooth=tuple(c["solref_smooth"]), torso_quat=tuple(c["object_quat"]),
        torso_pos=(0.0, 0.0, ar["table_top_z"] + z_offset - c["bottom_site"][2]),
        table_top_z=ar["table_top_z"], table_half_xy=ar["table_half_size"][0], table_friction=ar["table_friction"][0],
        probe_pos=tuple(g["body_pos"]), probe_mass=g["probe_mass"], probe_friction=g["probe_friction"][0],
    )


def scene_params_from_mjcf(models_dir: str, use_box_torso: bool = True, **overrides) -> SceneParams:
    """SceneParams with every in-tree constant taken from the reference's XML files under ``models_dir`` (= ``src/my_models``)."""
    from .model import cylinder_torso_params
    base = SceneParams() if use_box_torso else cylinder_torso_params()
    import dataclasses
    return dataclasses.replace(base, **{**scene_fields(read_assets(models_dir, use_box_torso)), **overrides})


def check_scene_params(params: SceneParams, parsed: Dict[str, Any], tol: float = 1e-12) -> Dict[str, Tuple[Any, Any]]:
    """Fields of ``params`` that disagree with the parsed files: {field: (SceneParams value, file value)} (empty = consistent)."""
    bad = {}
    for k, want in scene_fields(parsed).items():
        got = getattr(params, k)
        same = got == want if isinstance(want, str) else all(abs(float(x) - float(y)) <= tol for x, y in zip(_tup(got), _tup(want))) and len(_tup(got)) == len(_tup(want))
        if not same:
            bad[k] = (got, want)
    return bad


def _tup(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,)
