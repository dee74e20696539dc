#!/usr/bin/env python
"""Drift of the fp32 CUDA trajectories against the float64 oracle beyond the parity horizon (north_star: "bounded drift
reported beyond it").  Identical initial states and action sequences; max over envs of |dqpos|, |dqvel|, relative contact
force error and |dreward| per 50-step window, over a full 1000-step episode.

  python scripts/drift_report.py --json profiles/r01_drift.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (checker)
from rui_b200.abi import make_config  # noqa: E402
from rui_b200.env import BatchedUltrasound, packed_model  # noqa: E402

CC_FIXED = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3, kp=300,
                damping_ratio=1, impedance_mode="fixed", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0, uncouple_pos_ori=True)
CC_TRACK = dict(CC_FIXED, impedance_mode="tracking")


def run(soft, cc, steps, n, act_fn, tol=None, **kw):
    ekw = dict(kw, solver_tolerance=tol) if tol else kw  # device-side solver tolerance (the oracle always solves to 1e-10)
    env = BatchedUltrasound(n, soft_torso=soft, controller_configs=cc, control_freq=500, horizon=steps + 1, seed=3, **ekw)
    env.reset()
    q, v, w, t = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
    orcs = []
    for i in range(n):
        e = O.OracleEnv(packed_model(soft), make_config(1, cc, control_freq=500, horizon=steps + 1, seed=3, **kw), i)
        e.reset()
        e.set_state(q[i], v[i], w[i], t[i])
        orcs.append(e)
    rng = np.random.default_rng(0)
    lo, hi = env.action_spec
    rows, win = [], np.zeros(5)
    for s in range(steps):
        a = act_fn(s, n, rng, lo, hi)
        o, r, d, _ = env.step(torch.as_tensor(a, dtype=torch.float32), auto_reset=False)
        q, v = [x.cpu().numpy().astype(np.float64) for x in env.get_state()[:2]]
        o, r = o.cpu().numpy(), r.cpu().numpy()
        pairs_equal = 1.0
        for i in range(n):
            oo, orr, od = orcs[i].step(a[i])
            oq, ov = orcs[i].get_state()[:2]
            win = np.maximum(win, [np.abs(q[i] - oq).max(), np.abs(v[i] - ov).max(), abs(o[i, 2] - oo[2]) / max(1.0, abs(oo[2])), abs(r[i] - orr), 0])
            if i == 0:
                ncon, g1, g2, _ = env.contacts()
                k = int(ncon[0])
                c = orcs[0].contacts()
                same = k == len(c["geom1"]) and (g1[0, :k].cpu().numpy() == c["geom1"]).all() and (g2[0, :k].cpu().numpy() == c["geom2"]).all()
                pairs_equal = min(pairs_equal, float(same))
        win[4] = max(win[4], 1.0 - pairs_equal)
        if (s + 1) % 50 == 0:
            rows.append({"step": s + 1, "max_dqpos": win[0], "max_dqvel": win[1], "max_rel_dFz": win[2], "max_dreward": win[3],
                         "contact_list_mismatch_seen": bool(win[4])})
            print(rows[-1], flush=True)
            win = np.zeros(5)
    env.close()
    return rows


def press_then_random(s, n, rng, lo, hi):
    if s < 250:
        a = np.zeros((n, 6)); a[:, 2] = -1
        return a
    return np.repeat(rng.uniform(lo, hi, size=(1, 6)), n, 0)


def rnd(s, n, rng, lo, hi):
    return rng.uniform(lo, hi, size=(n, len(lo)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--tol", type=float, default=0.0, help="device solver tolerance (default: the library default)")
    a = ap.parse_args()
    out = {"rigid_press_config2": run(False, CC_FIXED, a.steps, 2, press_then_random, tol=a.tol or None),
           "soft_sweep_config3": run(True, CC_TRACK, a.steps, 3, rnd, tol=a.tol or None, torso_solref_randomization=True, initial_probe_pos_randomization=True)}
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
