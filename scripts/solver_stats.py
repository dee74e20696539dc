#!/usr/bin/env python
"""Solver work per env-step under the bench workload (4096 envs, steady-state episode mix, random gains), for the shipped probe
geometry and the round-1 axial capsule: CG iterations, preconditioner rebuilds, line-search evaluations, contacts.
  python scripts/solver_stats.py [--steps 200]"""
import argparse
import dataclasses
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ENV_OPTS, SEED, desync  # noqa: E402
from rui_b200.env import BatchedUltrasound  # noqa: E402
from rui_b200.model import SceneParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--envs", type=int, default=4096)
a = ap.parse_args()
dev = torch.device("cuda:0")
for name, sp in (("shipped (calibrated bar)", None),
                 ("axial capsule r=5cm (round 1)", dataclasses.replace(SceneParams(), probe_seg_a=None, probe_seg_b=None, probe_com=None, probe_radius=0.05))):
    env = BatchedUltrasound(a.envs, device=dev, seed=SEED, scene_params=sp, **ENV_OPTS)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED)
    acts = [torch.rand(a.envs, env.action_dim, device=dev, generator=gen) for _ in range(8)]
    desync(env, acts, gen, -1)
    acc = torch.zeros(6, dtype=torch.float64, device=dev)
    hist = torch.zeros(41, dtype=torch.float64, device=dev)
    for s in range(a.steps):
        env.step(acts[s % 8])
        d = env.diag()
        ncon, g1, g2, _ = env.contacts()
        probe = ((g2 == 2) & (g1 >= 4) & (torch.arange(g1.shape[1], device=dev)[None, :] < ncon[:, None])).sum(1).double()
        acc += torch.stack([d[:, 20].double().mean(), d[:, 24].double().mean(), d[:, 25].double().mean(), d[:, 22].double().mean(), probe.mean(),
                            (probe > 0).double().mean()])
        hist += torch.bincount(d[:, 20].long().clamp(0, 40), minlength=41).double()
    acc /= a.steps
    hist /= hist.sum()
    print(f"{name}: CG iterations {acc[0]:.2f}  rebuilds {acc[1]:.2f}  line-search evals {acc[2]:.2f} ({acc[2] / max(acc[0], 1e-9):.2f}/iteration)  "
          f"contacts {acc[3]:.1f}  probe-particle contacts {acc[4]:.2f}  envs in probe contact {acc[5]:.2f}")
    print("   iteration histogram:", " ".join(f"{k}:{100 * float(hist[k]):.1f}%" for k in range(41) if hist[k] > 0.002))
    env.close()
