#!/usr/bin/env python
"""Calibration of the substituted probe geometry (the reference's collision mesh is missing, A-PROBE-1) against the only
reference-produced force / torque numbers: the 192 raw post-reset observation rows of the shipped VecNormalize pickles
(tests/golden/art_stats.json: last_original_obs of the three models; same reset poses, different trajectory phase).

CPU only: float64 oracle resets (IK + one forward pass with ctrl = 0, as the reference's post-reset sim.forward()).

  python scripts/probe_calibrate.py                 # statistics of the shipped SceneParams next to the artifacts
  python scripts/probe_calibrate.py --sweep         # coarse sweep written to profiles/r02_probe_calibration.json
"""
import argparse
import dataclasses
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CC_TRACK  # noqa: E402
from oracle import oracle as O  # noqa: E402
from rui_b200 import abi  # noqa: E402
from rui_b200.abi import PackedModel  # noqa: E402
from rui_b200.model import SceneParams, build_model  # noqa: E402


def art_rows():
    with open(os.path.join(ROOT, "tests", "golden", "art_stats.json")) as f:
        a = json.load(f)
    return np.concatenate([np.array(a[k]["last_original_obs"]) for k in ("tracking", "variable_z", "wrench")])


def stats(rows):
    """summary statistics of post-reset rows (obs layout ultrasound.py:363-401)"""
    c = rows[:, 2] > 0
    nc = ~c
    lat = np.hypot(rows[c, 0], rows[c, 1]) / np.maximum(rows[c, 2], 1e-9)
    s = dict(n=len(rows), contact_frac=float(c.mean()), onset_z=float(rows[c, 14].max()) if c.any() else None,
             free_min_z=float(rows[nc, 14].min()) if nc.any() else None,
             fz_median=float(np.median(rows[c, 2])), fz_mean=float(rows[c, 2].mean()), fz_max=float(rows[c, 2].max()),
             lat_ratio_median=float(np.median(lat)), fx_mean=float(rows[c, 0].mean()), fy_mean=float(rows[c, 1].mean()),
             fx_over_fz=float(rows[c, 0].sum() / rows[c, 2].sum()), fy_over_fz=float(rows[c, 1].sum() / rows[c, 2].sum()),
             fx_std=float(rows[c, 0].std()), fy_std=float(rows[c, 1].std()),
             tq_free_mean=rows[nc, 3:6].mean(0).tolist() if nc.any() else None,
             tq_contact_mean=rows[c, 3:6].mean(0).tolist(), tq_contact_std=rows[c, 3:6].std(0).tolist())
    # force against depth: median Fz in bins of the z position error (eef above the trajectory height)
    edges = [-0.03, -0.01, 0.0, 0.005, 0.01, 0.0145]
    s["fz_by_depth"] = [float(np.median(rows[(rows[:, 14] >= lo) & (rows[:, 14] < hi) & c, 2])) if ((rows[:, 14] >= lo) & (rows[:, 14] < hi) & c).any() else None
                        for lo, hi in zip(edges[:-1], edges[1:])]
    return s


def reset_rows(params: SceneParams, n_envs=64, seeds=(3, 4, 5), threads=None):
    """post-reset observation rows of the oracle under rl_config.yaml's reset options"""
    pk = PackedModel(build_model(params))

    def one(args):
        seed, i = args
        e = O.OracleEnv(pk, abi.make_config(1, CC_TRACK, control_freq=500, seed=seed, torso_solref_randomization=True,
                                            initial_probe_pos_randomization=True), i)
        return e.reset()

    jobs = [(s, i) for s in seeds for i in range(n_envs)]
    with ThreadPoolExecutor(threads or os.cpu_count()) as ex:
        return np.array(list(ex.map(one, jobs)))


def show(name, s):
    f = lambda v: "None" if v is None else (f"{v:.3f}" if not isinstance(v, list) else "[" + ", ".join("nan" if x is None else f"{x:.3f}" for x in v) + "]")
    print(f"{name:28s} contact {s['contact_frac']:.2f} onset {f(s['onset_z'])} Fz med/mean/max {s['fz_median']:.1f}/{s['fz_mean']:.1f}/{s['fz_max']:.0f} "
          f"lat {s['lat_ratio_median']:.2f} Fx/Fz {s['fx_over_fz']:.3f} Fy/Fz {s['fy_over_fz']:.3f} sdFx {s['fx_std']:.1f} sdFy {s['fy_std']:.1f} "
          f"tq_free {f(s['tq_free_mean'])} tq_c_mean {f(s['tq_contact_mean'])} tq_c_std {f(s['tq_contact_std'])} Fz(depth) {f(s['fz_by_depth'])}", flush=True)


def loss(s, ref):
    """weighted squared mismatch of the summary statistics (log ratios for positive quantities)"""
    L = 0.0
    lr = lambda a, b: np.log(max(a, 1e-3) / max(b, 1e-3))
    L += 8 * ((s["contact_frac"] - ref["contact_frac"]) / 0.05) ** 2
    L += 4 * ((s["onset_z"] - ref["onset_z"]) / 0.002) ** 2
    for k in ("fz_median", "fz_mean", "fz_max", "fx_std", "fy_std"):
        L += (lr(s[k], ref[k]) / 0.2) ** 2
    L += 2 * (lr(s["lat_ratio_median"], ref["lat_ratio_median"]) / 0.2) ** 2
    for a, b in zip(s["fz_by_depth"], ref["fz_by_depth"]):
        if a is not None and b is not None:
            L += 0.5 * (lr(a, b) / 0.25) ** 2
    L += ((s["fx_over_fz"] - ref["fx_over_fz"]) / 0.04) ** 2 + ((s["fy_over_fz"] - ref["fy_over_fz"]) / 0.04) ** 2
    for a, b in zip(s["tq_free_mean"] or [0, 0, 0], ref["tq_free_mean"]):
        L += ((a - b) / 0.01) ** 2
    for a, b in zip(s["tq_contact_mean"], ref["tq_contact_mean"]):
        L += ((a - b) / 0.05) ** 2
    for a, b in zip(s["tq_contact_std"], ref["tq_contact_std"]):
        L += (lr(a, b) / 0.25) ** 2
    return float(L)


def unpack(x):
    """x = [ax, ay, az, bx, by, bz, radius, cx, cy, cz]"""
    return dict(probe_seg_a=tuple(x[0:3]), probe_seg_b=tuple(x[3:6]), probe_radius=float(x[6]), probe_com=tuple(x[7:10]))


def fit(ref, starts, iters, log_path):
    from scipy.optimize import minimize
    base = SceneParams()
    best = (1e30, None, None)
    hist = []

    def f(x):
        nonlocal best
        if not (0.003 <= x[6] <= 0.07) or np.abs(x[:6]).max() > 0.15 or np.abs(x[7:10]).max() > 0.2 or np.linalg.norm(x[0:3] - x[3:6]) < 1e-3:
            return 1e6
        try:
            st = stats(reset_rows(dataclasses.replace(base, **unpack(x))))
            L = loss(st, ref)
        except Exception:
            return 1e6
        hist.append((L, x.tolist()))
        if L < best[0]:
            best = (L, x.copy(), st)
            show(f"L={L:8.2f}", st)
            print("   x =", np.round(x, 4).tolist(), flush=True)
            with open(log_path, "w") as fh:
                json.dump(dict(loss=L, x=x.tolist(), params={k: (list(v) if isinstance(v, tuple) else v) for k, v in unpack(x).items()},
                               stats=st, art=ref, evaluations=len(hist)), fh, indent=1)
        return L

    for x0 in starts:
        x0 = np.asarray(x0, float)
        step = np.array([0.01] * 6 + [0.004] + [0.01] * 3)
        simplex = np.vstack([x0] + [x0 + step * e for e in np.eye(10)])
        minimize(f, x0, method="Nelder-Mead", options=dict(maxfev=iters, initial_simplex=simplex, xatol=2e-4, fatol=1e-2))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--fit", type=int, default=0, help="Nelder-Mead evaluations per start")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    ref = stats(art_rows())
    show("ART (reference artifacts)", ref)
    base = SceneParams()
    out = {"art": ref, "candidates": {}}

    def ev(name, **kw):
        s = stats(reset_rows(dataclasses.replace(base, **kw)))
        show(name, s)
        out["candidates"][name] = dict(params={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, stats=s)

    ev("shipped SceneParams")
    print("loss of the shipped parameters:", round(loss(out["candidates"]["shipped SceneParams"]["stats"], ref), 2))
    if args.fit:
        starts = json.loads(os.environ.get("PROBE_STARTS", "[]")) or [
            [0, 0, -0.05, 0, 0, -0.10, 0.05, 0, 0, -0.075],              # round-1 axial capsule
            [0.02, 0.0, -0.012, -0.02, 0.0, -0.012, 0.012, 0, 0.015, -0.05],   # bar across body x
            [0.0, 0.02, -0.012, 0.0, -0.02, -0.020, 0.012, 0, 0.015, -0.05],   # tilted bar across body y
        ]
        fit(ref, starts, args.fit, args.json or os.path.join(ROOT, "gpurun_out", "probe_fit.json"))
        return
    if args.sweep:
        for cand in json.loads(os.environ.get("PROBE_CANDIDATES", "[]")):
            nm = cand.pop("name")
            ev(nm, **{k: (tuple(v) if isinstance(v, list) else v) for k, v in cand.items()})
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
