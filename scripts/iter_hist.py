"""Developer script: distribution of CG iterations / contacts per env over a random-action rollout (run under gpurun)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from rui_b200.env import BatchedUltrasound

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedUltrasound(n, device=0, seed=3, **bench.ENV_OPTS)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(3)
hist = np.zeros(82, dtype=np.int64); nc = []
for s in range(120):
    env.step(torch.rand(n, 6, device="cuda", generator=gen))
    if s >= 20:
        d = env.diag().cpu().numpy()
        hist += np.bincount(d[:, 20].astype(int), minlength=82)[:82]
        nc.append(d[:, 22])
nc = np.concatenate(nc)
tot = hist.sum()
print("iters: mean", (hist * np.arange(82)).sum() / tot)
print("hist:", {i: round(100 * h / tot, 2) for i, h in enumerate(hist) if h})
print("cum>=: ", {k: round(100 * hist[k:].sum() / tot, 3) for k in (10, 15, 20, 30, 40)})
print("ncon: mean", nc.mean(), "p50", np.percentile(nc, 50), "p99", np.percentile(nc, 99), "max", nc.max())
# iterations split by probe contact (is the missing arm<->torso coupling of the preconditioner visible?)
t = env.get_state()[3].cpu().numpy()
d = env.diag().cpu().numpy()
inc = t[:, 30] > 0
print("in contact: %.3f of envs; iters in contact %.2f, free %.2f; ncon in contact %.1f free %.1f" % (inc.mean(), d[inc, 20].mean(), d[~inc, 20].mean(), d[inc, 22].mean(), d[~inc, 22].mean()))
