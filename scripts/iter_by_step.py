"""Developer script: mean CG iterations / line-search effort as the episode proceeds (run under gpurun)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from rui_b200.env import BatchedUltrasound

n = 4096
env = BatchedUltrasound(n, device=0, seed=3, **bench.ENV_OPTS)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(3)
rows = []
for s in range(200):
    env.step(torch.rand(n, 6, device="cuda", generator=gen))
    d = env.diag()
    rows.append((float(d[:, 20].mean()), float(d[:, 22].mean()), float(d[:, 21].median())))
for lo, hi in ((0, 3), (3, 10), (10, 20), (20, 40), (40, 80), (80, 140), (140, 200)):
    r = np.array(rows[lo:hi])
    print(f"steps {lo:3d}-{hi:3d}: mean iterations {r[:, 0].mean():.2f}  mean ncon {r[:, 1].mean():.1f}  median final |grad| {r[:, 2].mean():.2e}")
# the same with a constant action (no control noise)
env.reset()
a = torch.full((n, 6), 0.5, device="cuda")
rows = []
for s in range(60):
    env.step(a)
    rows.append(float(env.diag()[:, 20].mean()))
print("constant action 0.5: mean iterations steps 0-10 %.2f, 10-30 %.2f, 30-60 %.2f" % (np.mean(rows[:10]), np.mean(rows[10:30]), np.mean(rows[30:])))
