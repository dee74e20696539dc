"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/usim.h declares,
the ctypes struct mirrors match the compiled structs, the config mapping follows rl_config.yaml, the model
builder reproduces the composite numbers of SURVEY App. B, and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import CC_FIXED, CC_TRACK, ROOT
from rui_b200 import _lib, abi
from rui_b200.model import SceneParams, build_model


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "usim.h")).read()
    declared = set(re.findall(r"\b(usim_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.usim_abi_version() == abi.USIM_ABI_VERSION


def test_struct_mirrors_match_compiled_layout():
    L = _lib.lib()
    assert L.usim_sizeof_model() == ctypes.sizeof(abi.UsimModel)
    assert L.usim_sizeof_config() == ctypes.sizeof(abi.UsimConfig)
    hdr = open(os.path.join(ROOT, "include", "usim.h")).read()
    assert int(re.search(r"#define USIM_TASK_DIM (\d+)", hdr).group(1)) == abi.TASK_DIM
    assert int(re.search(r"#define USIM_MAX_CONTACTS (\d+)", hdr).group(1)) == abi.MAX_CONTACTS
    assert int(re.search(r"#define USIM_OBS_DIM (\d+)", hdr).group(1)) == abi.OBS_DIM
    assert int(re.search(r"#define USIM_DIAG_DIM (\d+)", hdr).group(1)) == abi.DIAG_DIM
    assert int(re.search(r"#define USIM_ARM_RECORD_DIM (\d+)", hdr).group(1)) == abi.ARM_RECORD_DIM
    for name, val in re.findall(r"(USIM_TS_[A-Z_]+) = (\d+)", hdr):
        py = name.replace("USIM_", "")
        if hasattr(abi, py):
            assert getattr(abi, py) == int(val), name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from rui_b200.env import BatchedUltrasound, UltrasoundVecEnv, make
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedUltrasound(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        UltrasoundVecEnv(2, dict(controller_configs=CC_TRACK, control_freq=500))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        make("Ultrasound", robots="Panda", controller_configs=CC_TRACK, control_freq=500)
    # the C entry point itself refuses too
    from rui_b200.env import packed_model
    pk = packed_model(False)
    cfg = abi.make_config(2, CC_FIXED, control_freq=500)
    h = ctypes.c_void_p()
    rc = _lib.lib().usim_create(ctypes.byref(pk.struct), ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in _lib.lib().usim_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "robotic-ultrasound-imaging_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liboracle", r"#include\s+\"[^\"]*oracle", r"import_module\([^)]*oracle", r"oracle/"):
                    assert re.search(pat, src, re.M) is None, (f, pat)
    assert "#include" not in open(os.path.join(ROOT, "include", "usim.h")).read().replace("#include <stddef.h>", "").replace("#include <stdint.h>", "")


def test_config_mapping_follows_rl_config_yaml():
    import yaml
    cfg_yaml = yaml.safe_load("""
      type: "OSC_POSE"
      input_max: 1
      input_min: -1
      output_max: [0.05, 0.05, 0.05, 0.5, 0.5, 0.5]
      output_min: [-0.05, -0.05, -0.05, -0.5, -0.5, -0.5]
      kp: 300
      damping_ratio: 1
      impedance_mode: "tracking"
      kp_limits: [0, 500]
      kp_input_max: 1
      kp_input_min: 0
      damping_ratio_limits: [0, 2]
      position_limits: null
      orientation_limits: null
      uncouple_pos_ori: True
      control_delta: True
      interpolation: null
      ramp_ratio: 0.2
    """)
    c = abi.make_config(64, cfg_yaml, control_freq=500, horizon=1000, early_termination=True, torso_solref_randomization=True,
                        initial_probe_pos_randomization=True, seed=3)
    assert c.impedance_mode == abi.MODE_TRACKING and abi.action_dim(c) == 6
    lo, hi = abi.action_bounds(c)
    assert lo.tolist() == [0] * 6 and hi.tolist() == [1] * 6  # [ART] action space of tracking.zip
    assert list(c.kp) == [300] * 6 and list(c.kp_limits) == [0, 500] and c.uncouple_pos_ori == 1
    assert c.early_termination == 1 and c.solref_randomization == 1 and c.probe_pos_randomization == 1 and c.seed == 3
    c2 = abi.make_config(1, dict(cfg_yaml, impedance_mode="variable_z"), control_freq=500)
    lo, hi = abi.action_bounds(c2)
    assert abi.action_dim(c2) == 7 and lo[-1] == -1 and hi[-1] == 1
    lo, hi = abi.action_bounds(abi.make_config(1, dict(cfg_yaml, impedance_mode="wrench"), control_freq=500))
    assert lo.tolist() == [-10] * 6 and hi.tolist() == [10] * 6
    lo, hi = abi.action_bounds(abi.make_config(1, dict(cfg_yaml, impedance_mode="fixed"), control_freq=500))
    assert lo.tolist() == [-1] * 6 and hi.tolist() == [1] * 6
    with pytest.raises(AssertionError):
        abi.make_config(1, dict(cfg_yaml, type="JOINT_POSITION"))


def test_art_action_spaces_match(art):
    for model, mode in (("tracking", "tracking"), ("variable_z", "variable_z"), ("wrench", "wrench")):
        lo, hi = abi.action_bounds(abi.make_config(1, dict(CC_TRACK, impedance_mode=mode), control_freq=500))
        np.testing.assert_allclose(lo, art[model]["action_low"])
        np.testing.assert_allclose(hi, art[model]["action_high"])


def test_composite_numbers(soft_model):
    m = soft_model.model
    assert (m.nq, m.nv, m.nbody) == (284, 283, 282)
    assert len(m.part_pos) == 270 and len(m.eq_pairs) == 536  # App. B.2
    pos = m.part_pos
    np.testing.assert_allclose(np.abs(pos).max(axis=0), [0.14, 0.0525, 0.175])
    assert (pos[:, 1] == pos[:, 1].max()).sum() == 99 and (pos[:, 1] == pos[:, 1].min()).sum() == 99
    assert m.particle_names[0] == "G0_0_0" and m.particle_names[-1] == "G8_3_10"
    assert all(re.search(r"[G]\d+[_]\d+[_]\d+$", m.geom_name(4 + k)) for k in range(270))  # ultrasound.py:724
    assert not re.search(r"[G]\d+[_]\d+[_]\d+$", m.geom_name(3))
    np.testing.assert_allclose(np.linalg.norm(m.part_axis, axis=1), 1)
    # torso body rotation maps the thin local axis (y) to world z: world half extents (0.175, 0.14, 0.0525)
    from rui_b200.model import quat2mat
    R = quat2mat(m.params.torso_quat)
    np.testing.assert_allclose(np.abs(pos @ R.T).max(axis=0), [0.175, 0.14, 0.0525], atol=1e-12)
    # every particle has 2..5 neighbours, table is symmetric
    nb = m.part_nbr
    deg = (nb >= 0).sum(axis=1)
    assert deg.min() >= 3 and deg.max() <= 4 and deg.sum() == 2 * 536
    for i in range(270):
        for j in nb[i][nb[i] >= 0]:
            assert i in nb[j]
    assert abs(m.g_body_mass[m.ids[6]:].sum() - 2.7) < 1e-12


def test_rigid_model_and_arm_tables(rigid_model):
    m = rigid_model.model
    assert (m.nq, m.nv) == (7, 7)
    L = m.arm_link
    assert abs(L[6, 15] - (0.5 + 0.5 + 1.0)) < 1e-12  # link 7 + hand + probe welded
    assert abs(L[:, 15].sum() - (3 + 3 + 2 + 2 + 2 + 1.5 + 2.0)) < 1e-12
    assert m.dof_invweight0.shape == (7,) and np.all(m.dof_invweight0 > 0)


def test_shard_range_partitions_the_env_ids():
    from rui_b200.dist import shard_range
    for total, world in ((65536, 8), (65536, 2), (4096, 1), (10, 4), (7, 8)):
        seen = []
        for r in range(world):
            off, n = shard_range(total, r, world)
            seen += list(range(off, off + n))
        assert seen == list(range(total))
    with pytest.raises(ValueError):
        shard_range(8, 3, 2)


def test_solver_knobs_reach_the_config():
    """The device solver's knobs travel through make_config unchanged; the defaults are the documented ones."""
    c = abi.make_config(8, CC_TRACK, control_freq=500)
    assert c.solver_iterations == 40 and c.solver_tolerance == pytest.approx(3e-5) and c.precond_rebuilds == 0  # 0 = library default (8); tolerance: profiles/r02_solver_tolerance.md
    c = abi.make_config(8, CC_TRACK, control_freq=500, solver_iterations=12, solver_tolerance=3e-6, precond_rebuilds=3)
    assert (c.solver_iterations, c.precond_rebuilds) == (12, 3) and c.solver_tolerance == pytest.approx(3e-6)


def test_stencil_degree_fits_the_packed_table():
    """The device keeps each element's pair stencil as ONE 16-byte record of four packed (pair, sign, neighbour) entries
    (csrc/common.cuh PartTables::nb4; usim_create rejects anything wider): the shell grid of both torso shapes has degree <= 4,
    every pair appears in the neighbour lists of both of its ends, and the indices fit the bit fields."""
    from rui_b200.model import cylinder_torso_params
    for m in (build_model(), build_model(cylinder_torso_params())):
        a = m.arrays
        nbr = np.asarray(a["part_nbr"]).reshape(-1, 6)
        pairs = np.asarray(a["eq_pairs"]).reshape(-1, 2)
        npart, npair = nbr.shape[0], pairs.shape[0]
        assert ((nbr >= 0).sum(1) <= 4).all() and npart <= 272 and npair <= 543  # NPART_MAX, NPAIR_MAX - 1 (the empty slot)
        assert npart < (1 << 15) and npair < (1 << 15)
        ends = {(int(i), int(j)) for i, j in pairs} | {(int(j), int(i)) for i, j in pairs}
        listed = {(i, int(j)) for i in range(npart) for j in nbr[i] if j >= 0}
        assert listed == ends


def test_bench_roofline_helpers_read_the_committed_capture():
    """bench.py derives `roofline.traffic` and the informative issue-slot object from profiles/traffic.json (the committed ncu capture)."""
    import json

    import bench
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert t["kernel"] == "solve_kernel" and t["envs_per_launch"] == 4096
    assert bench.ncu_traffic(4096) == pytest.approx(t["dram_bytes_per_launch"]) and bench.ncu_traffic(8192) == pytest.approx(2 * t["dram_bytes_per_launch"])
    # traffic at or below the algorithmic bytes: the state stays in L2 between steps, nothing is re-read from HBM
    assert t["dram_bytes_per_launch"] <= 4096 * bench.ALG_BYTES_PER_ENV_STEP
    iss = bench.issue_slots(4096, 0.50, 1965.0, t["mean_solver_iters_of_capture"])
    assert iss["unit"] == "G warp-instr/s" and 0.2 < iss["frac"] < 1.0 and iss["peak"] == pytest.approx(4 * 148 * 1.965)
    assert iss["achieved"] == pytest.approx(t["warp_instructions_per_launch"] / 0.50e-3 / 1e9)  # same iteration count as the capture: unscaled
    # the loop share of the instructions scales with the LIVE iteration count the kernel reports
    more = bench.issue_slots(4096, 0.50, 1965.0, 2 * t["mean_solver_iters_of_capture"])
    assert more["achieved"] == pytest.approx(iss["achieved"] * (1 + t["loop_share_of_instructions"]))
    assert bench.issue_slots(4096, 0.43, None, 5.0) is None


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ carry every key of the measurement contract (a stale or hand-edited line would not)."""
    import json
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}
    for name, n in (("r01_bench.json", 1), ("r01_bench_2gpu.json", 2), ("r01_bench_4gpu.json", 4), ("r01_bench_8gpu.json", 8),
                    ("r02_bench.json", 1), ("r02_bench_2gpu.json", 2), ("r02_bench_driver_style.json", 1)):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert need <= set(d), need - set(d)
        assert d["n_gpus"] == n and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
        assert "workload" in d["config"] and d["gpu_launches"] > 0 and d["vs_baseline"] is None
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and 0 < d["e2e"]["value"] <= d["value"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["frac"] == pytest.approx(r["achieved"] / r["peak"]) and r["kernel"] == "solve_kernel"
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] > 0
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "port"


def _fake_sb3(monkeypatch):
    """stable-baselines3 is not installed here: a stand-in package with the ABSTRACT surface of SB3 1.1's
    ``stable_baselines3.common.vec_env.base_vec_env.VecEnv`` (the class ``VecNormalize`` / ``PPO`` type-check against)."""
    import abc
    import sys
    import types

    class VecEnv(abc.ABC):
        metadata = {"render.modes": ["human", "rgb_array"]}

        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

        @abc.abstractmethod
        def reset(self): ...
        @abc.abstractmethod
        def step_async(self, actions): ...
        @abc.abstractmethod
        def step_wait(self): ...
        @abc.abstractmethod
        def close(self): ...
        @abc.abstractmethod
        def get_attr(self, attr_name, indices=None): ...
        @abc.abstractmethod
        def set_attr(self, attr_name, value, indices=None): ...
        @abc.abstractmethod
        def env_method(self, method_name, *method_args, indices=None, **method_kwargs): ...
        @abc.abstractmethod
        def env_is_wrapped(self, wrapper_class, indices=None): ...
        @abc.abstractmethod
        def seed(self, seed=None): ...

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

    pkg, common, vec = types.ModuleType("stable_baselines3"), types.ModuleType("stable_baselines3.common"), types.ModuleType("stable_baselines3.common.vec_env")
    vec.VecEnv = VecEnv
    pkg.common, common.vec_env = common, vec
    for name, mod in (("stable_baselines3", pkg), ("stable_baselines3.common", common), ("stable_baselines3.common.vec_env", vec)):
        monkeypatch.setitem(sys.modules, name, mod)
    return VecEnv


def test_sb3_adapter_is_a_real_vecenv_subclass(monkeypatch):
    """rl.py:130-143 hands the vectorised env to SB3's ``VecNormalize`` and ``PPO``, which check ``isinstance(env, VecEnv)``.  The
    adapter class is created against whatever ``stable_baselines3`` is importable; it must implement the whole abstract surface."""
    from rui_b200 import env as E
    base = _fake_sb3(monkeypatch)
    cls = E.sb3_vec_env_class()
    assert issubclass(cls, base) and not getattr(cls, "__abstractmethods__", None), cls.__abstractmethods__
    for m in ("reset", "step_async", "step_wait", "step", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped", "seed"):
        assert callable(getattr(cls, m))
    # the stand-alone VecEnv-shaped class (no SB3 needed) has the same method set
    for m in ("reset", "step_async", "step_wait", "step", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped", "seed"):
        assert callable(getattr(E.UltrasoundVecEnv, m))


def test_sb3_adapter_needs_sb3():
    """without stable-baselines3 the adapter raises ImportError (the VecEnv-shaped class and the built-in PPO remain)"""
    import importlib.util
    if importlib.util.find_spec("stable_baselines3") is not None:
        pytest.skip("stable-baselines3 is installed")
    from rui_b200 import env as E
    with pytest.raises(ImportError):
        E.sb3_vec_env_class()
