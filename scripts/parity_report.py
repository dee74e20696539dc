#!/usr/bin/env python
"""Measured fp32-vs-float64 drift of the CUDA path against the oracle (the numbers the test tolerances are derived from).

  python scripts/parity_report.py [--envs 64] [--steps 300] [--out gpurun_out/parity_drift.json]

Runs the comparisons of tests/test_gpu_parity.py (same helper, tests/parity_util.py) and writes the per-quantity maxima.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CC_FIXED, CC_TRACK  # noqa: E402
from parity_util import compare_rollout, make_oracles  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_drift.json"))
    ap.add_argument("--tol", type=float, default=0.0, help="device solver tolerance (0: library default)")
    ap.add_argument("--only", default="", help="run only the cases whose name contains this")
    args = ap.parse_args()
    from oracle import oracle as O
    from rui_b200.env import BatchedUltrasound

    res = {}
    kw = dict(torso_solref_randomization=True, initial_probe_pos_randomization=True, seed=3)

    def run(name, cc, soft, n, steps, lo, hi, adim=6, **extra):
        if args.only and args.only not in name:
            return
        if args.tol:
            extra["solver_tolerance"] = args.tol
        t0 = time.time()
        env = BatchedUltrasound(n, device=0, soft_torso=soft, controller_configs=cc, **{"control_freq": 500, "horizon": 1000, **extra})
        env.reset()
        orcs = make_oracles(O, env, cc, soft=soft, **{"control_freq": 500, "horizon": 1000, **extra})
        rng = np.random.default_rng(0)
        acts = rng.uniform(lo, hi, size=(steps, n, adim))
        marks = {}

        def on_step(s, dr):
            if s + 1 in (10, 30, 60, 100, 200, 300, 500, 1000):
                marks[s + 1] = dict(dr.max)

        dr, log = compare_rollout(O, env, orcs, acts, on_step=on_step)
        res[name] = dict(envs=n, steps=log["steps"], env_steps=log["env_steps"], max=dr.max, max_incl_threshold_steps=log["drift_all"].max,
                         threshold_env_steps=log["threshold_env_steps"], threshold_events=log["threshold_events"], by_step=marks,
                         done_mismatch=log["done_mismatch"], contact_mismatch=len(log["contact_mismatch"]),
                         terminated=int(log["terminated"].sum()), overflow=env.contact_overflow_count,
                         seconds=round(time.time() - t0, 1))
        print(name, json.dumps({k: float(f"{v:.3g}") for k, v in dr.max.items()}), "done_mismatch", len(log["done_mismatch"]),
              "contact_mismatch", len(log["contact_mismatch"]), "threshold env-steps", log["threshold_env_steps"], "of", log["env_steps"],
              "all:", json.dumps({k: float(f"{v:.2g}") for k, v in log["drift_all"].max.items() if k in ("force_rel", "reward", "dfz_rel", "qvel")}), "terminated", int(log["terminated"].sum()), flush=True)
        env.close()

    run("config3_soft_tracking", CC_TRACK, True, args.envs, args.steps, 0, 1, **kw)
    if args.only and "long" in args.only:  # bounded drift beyond the short horizon (north_star): a full 1000-step episode
        run("long_1000_steps", CC_TRACK, True, 8, 1000, 0, 1, **kw)
    run("config3_early_termination", CC_TRACK, True, args.envs, 250, 0, 1, early_termination=True, **dict(kw, horizon=250))
    run("substeps_5_control_freq_100", CC_TRACK, True, 8, 30, 0, 1, **dict(kw, control_freq=100))
    run("substeps_25_control_freq_20", CC_TRACK, True, 8, 12, 0, 1, **dict(kw, control_freq=20))
    run("config2_rigid_fixed_random", CC_FIXED, False, 16, 300, -1, 1)
    # BASELINE config 2 proper: press on the table, then random actions (the sequence of tests/test_gpu_parity.py)
    if not args.only or args.only in "config2_rigid_press":
        from test_gpu_parity import _config2_actions
        env = BatchedUltrasound(64, device=0, soft_torso=False, controller_configs=CC_FIXED, control_freq=500, horizon=1000)
        env.reset()
        orcs = make_oracles(O, env, CC_FIXED, n=1, soft=False)
        dr, log = compare_rollout(O, env, orcs, _config2_actions(64))
        res["config2_rigid_press"] = dict(envs=64, steps=log["steps"], max=dr.max, max_incl_threshold_steps=log["drift_all"].max,
                                          threshold_env_steps=log["threshold_env_steps"], done_mismatch=log["done_mismatch"],
                                          contact_mismatch=len(log["contact_mismatch"]))
        print("config2_rigid_press", json.dumps({k: float(f"{v:.3g}") for k, v in dr.max.items()}), "threshold env-steps", log["threshold_env_steps"],
              "all:", json.dumps({k: float(f"{v:.2g}") for k, v in log["drift_all"].max.items() if k in ("force_rel", "reward", "qvel", "qpos")}), flush=True)
        env.close()
    run("wrench", dict(CC_TRACK, impedance_mode="wrench"), True, 8, 60, -10, 10, **kw)
    run("variable_z", dict(CC_TRACK, impedance_mode="variable_z"), True, 8, 60, np.r_[np.zeros(6), -1], np.ones(7), adim=7, **kw)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1, default=float)


if __name__ == "__main__":
    main()
