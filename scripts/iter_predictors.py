"""What predicts a long solve?  Per env-step of the staggered benchmark mix: CG iterations against quantities known BEFORE the launch
(iterations / rebuilds / contact count of the previous steps, warm-start shift of the arm, action change, contact flag)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import bench
from rui_b200 import abi
from rui_b200.env import BatchedUltrasound
n = 4096
env = BatchedUltrasound(n, device=0, seed=3, **bench.ENV_OPTS)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(3)
q, v, w, t = env.get_state()
t[:, abi.TS_TIMESTEP] = torch.randint(0, 1000, (n,), device="cuda").float()
env.set_state(task=t)
a_prev = torch.rand(n, 6, device="cuda", generator=gen)
for s in range(300):
    env.step(a_prev, auto_reset=True); a_prev = torch.rand(n, 6, device="cuda", generator=gen)
rows = []
d1 = env.diag().clone(); d2 = d1.clone()
for s in range(200):
    a = torch.rand(n, 6, device="cuda", generator=gen)
    env.step(a, auto_reset=True)
    d = env.diag().clone()
    rec = env.arm_record()
    ts = env.get_state()[3]
    feat = torch.stack([d[:, 20], d1[:, 20], d2[:, 20], d1[:, 24], d1[:, 22], rec[:, 172:179].abs().amax(1), (a - a_prev).abs().amax(1),
                        ts[:, abi.TS_IN_CONTACT], ts[:, abi.TS_TIMESTEP], torch.maximum(d1[:, 20], d2[:, 20]), d1[:, 25]], 1)
    rows.append(feat.cpu().numpy())
    d2, d1, a_prev = d1, d, a
X = np.concatenate(rows)
X = X[X[:, 8] > 3]  # skip the first steps of an episode (cold start, handled by the order already)
names = ["iters", "iters_prev", "iters_prev2", "rebuilds_prev", "ncon_prev", "arm_shift_max", "action_change_max", "in_contact", "timestep", "max_prev2", "ls_evals_prev"]
y = X[:, 0]
print("samples", len(y), "mean iters %.2f" % y.mean(), "P(>=10) %.4f" % (y >= 10).mean(), "P(>=15) %.5f" % (y >= 15).mean())
from scipy.stats import spearmanr
for k in range(1, len(names)):
    r = spearmanr(X[:, k], y).correlation
    # of the 5% env-steps the predictor ranks highest, how many of the long solves (>= 10 iterations) are caught
    thr = np.quantile(X[:, k], 0.95)
    caught = ((X[:, k] >= thr) & (y >= 10)).sum() / max(1, (y >= 10).sum())
    frac = (X[:, k] >= thr).mean()
    print(f"{names[k]:20s} spearman {r:+.3f}   top {100*frac:.1f}% by this feature catch {100*caught:.1f}% of the >=10-iteration solves")
# transition table
for p in (3, 4, 5, 6, 8):
    m = X[:, 1] == p
    if m.sum() > 100:
        print(f"prev {p}: n {m.sum()}, now mean {y[m].mean():.2f}, P(>=8) {(y[m] >= 8).mean():.3f}, P(>=12) {(y[m] >= 12).mean():.4f}")
