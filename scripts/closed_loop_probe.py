#!/usr/bin/env python
"""Closed-loop behavioural probe (use 2): replay the reference's shipped `tracking` PPO policy, with its
own VecNormalize statistics, in the B200 env under the rl_config.yaml settings, and compare the observation statistics,
episode lengths and per-step reward with the 40 M-sample statistics stored in the reference's artifacts.

Not a parity gate (the physics constants of the arm/probe are recalled or substituted): reports ratios.
  python scripts/closed_loop_probe.py [--envs 4096] [--steps 3000] [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rui_b200.env import BatchedUltrasound  # noqa: E402
from rui_b200.ppo import MlpPolicy  # noqa: E402

CC = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3, kp=300,
          damping_ratio=1, impedance_mode="tracking", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0, uncouple_pos_ori=True)


def run(envs=4096, steps=3000, deterministic=False, seed=3, device=0, scene_params=None):
    fix = np.load(os.path.join(ROOT, "tests", "golden", "tracking_policy.npz"))
    art = json.load(open(os.path.join(ROOT, "tests", "golden", "art_stats.json")))["tracking"]
    dev = torch.device(f"cuda:{device}")
    pol = MlpPolicy(19, 6).to(dev)
    pol.load_state_dict({k: torch.as_tensor(fix[k]) for k in pol.state_dict().keys()})
    mean = torch.as_tensor(fix["obs_mean"], dtype=torch.float32, device=dev)
    std = torch.sqrt(torch.as_tensor(fix["obs_var"], dtype=torch.float32, device=dev) + 1e-8)
    env = BatchedUltrasound(envs, device=dev, controller_configs=CC, control_freq=500, horizon=1000, early_termination=True,
                            torso_solref_randomization=True, initial_probe_pos_randomization=True, seed=seed, scene_params=scene_params)
    obs = env.reset().clone()
    reset_obs = obs.clone()
    torch.manual_seed(seed)
    n = 0
    s1 = torch.zeros(19, dtype=torch.float64, device=dev)
    s2 = torch.zeros(19, dtype=torch.float64, device=dev)
    rsum = torch.zeros((), dtype=torch.float64, device=dev)
    ep_len = torch.zeros(envs, device=dev)
    ep_ret = torch.zeros(envs, dtype=torch.float64, device=dev)
    lens, rets = [], []
    with torch.no_grad():
        for t in range(steps):
            nobs = torch.clamp((obs - mean) / std, -10, 10)
            a = pol.act(nobs, deterministic=deterministic)[0]
            o, r, d, tobs = env.step(torch.clamp(a, 0, 1), auto_reset=True)
            # statistics over the observations the policy sees (post-step obs; terminal obs for finished envs, as VecNormalize sees them)
            seen = torch.where(d.bool().unsqueeze(1), tobs, o).to(torch.float64)
            s1 += seen.sum(0); s2 += (seen * seen).sum(0); n += envs
            rsum += r.sum().to(torch.float64)
            ep_len += 1; ep_ret += r.to(torch.float64)
            if d.any():
                idx = d.bool()
                lens += ep_len[idx].tolist(); rets += ep_ret[idx].tolist()
                ep_len[idx] = 0; ep_ret[idx] = 0
            obs = o.clone()
    m = (s1 / n).cpu().numpy()
    v = (s2 / n).cpu().numpy() - m * m
    ro = reset_obs.cpu().numpy()
    aro = np.array(art["last_original_obs"])
    out = {
        "envs": envs, "steps": steps, "episodes": len(lens),
        "obs_mean": m.tolist(), "obs_var": v.tolist(), "art_obs_mean": art["obs_mean"], "art_obs_var": art["obs_var"],
        "step_reward": float(rsum / n), "art_step_reward": float(np.sum(art["ep_returns"]) / np.sum(art["ep_lengths"])),
        "ep_len_mean": float(np.mean(lens)) if lens else None, "art_ep_len_mean": float(np.mean(art["ep_lengths"])),
        "ep_ret_mean": float(np.mean(rets)) if rets else None, "art_ep_ret_mean": float(np.mean(art["ep_returns"])),
        "full_length_fraction": float(np.mean(np.array(lens) >= 1000)) if lens else None,
        "art_full_length_fraction": float(np.mean(np.array(art["ep_lengths"]) >= 1000)),
        "reset_in_contact_fraction": float((ro[:, 2] > 0).mean()), "art_reset_in_contact_fraction": float((aro[:, 2] > 0).mean()),
        "reset_fz_median_in_contact": float(np.median(ro[ro[:, 2] > 0, 2])), "art_reset_fz_median_in_contact": float(np.median(aro[aro[:, 2] > 0, 2])),
        "reset_pos_err_mean": ro[:, 12:15].mean(0).tolist(), "art_reset_pos_err_mean": aro[:, 12:15].mean(0).tolist(),
        "reset_pos_err_std": ro[:, 12:15].std(0).tolist(), "art_reset_pos_err_std": aro[:, 12:15].std(0).tolist(),
    }
    env.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--json", default="")
    ap.add_argument("--probe-json", default="", help="SceneParams overrides (a scripts/probe_calibrate.py --fit result, or a plain dict)")
    a = ap.parse_args()
    sp = None
    if a.probe_json:
        import dataclasses

        from rui_b200.model import SceneParams
        ov = json.load(open(a.probe_json))
        ov = ov.get("params", ov)
        sp = dataclasses.replace(SceneParams(), **{k: (tuple(v) if isinstance(v, list) else v) for k, v in ov.items()})
    out = run(a.envs, a.steps, scene_params=sp)
    names = ["Fx", "Fy", "Fz", "tq_x", "tq_y", "tq_z", "vx", "vy", "vz", "Fz_mean-5", "dFz", "vel_mean-.04", "ex", "ey", "ez", "q0", "q1", "q2", "q3"]
    print(f"{'channel':14s} {'mean':>10s} {'ART mean':>10s} {'var':>12s} {'ART var':>12s}")
    for i, nme in enumerate(names):
        print(f"{nme:14s} {out['obs_mean'][i]:10.4f} {out['art_obs_mean'][i]:10.4f} {out['obs_var'][i]:12.4g} {out['art_obs_var'][i]:12.4g}")
    for k in ("step_reward", "ep_len_mean", "ep_ret_mean", "full_length_fraction", "reset_in_contact_fraction", "reset_fz_median_in_contact",
              "reset_pos_err_mean", "reset_pos_err_std"):
        print(f"{k:28s} ours {out[k]}   ART {out['art_' + k]}")
    print("episodes finished:", out["episodes"])
    rat = lambda x, y: float("nan") if y == 0 else x / y
    print("ratios ours / ART  mean:", " ".join(f"{nme}={rat(out['obs_mean'][i], out['art_obs_mean'][i]):.2f}" for i, nme in enumerate(names[:6])))
    print("ratios ours / ART   var:", " ".join(f"{nme}={rat(out['obs_var'][i], out['art_obs_var'][i]):.2f}" for i, nme in enumerate(names[:12])))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
