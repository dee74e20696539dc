"""What do long solves look like?  Rebuilds and line-search evaluations per iteration, contact counts, by iteration-count class."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
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
for s in range(300):
    env.step(torch.rand(n, 6, device="cuda", generator=gen), auto_reset=True)
rows = []
prev_nc = env.diag()[:, 22].clone()
for s in range(300):
    env.step(torch.rand(n, 6, device="cuda", generator=gen), auto_reset=True)
    d = env.diag(); ts = env.get_state()[3]
    ncon, g1, g2, dist = env.contacts()
    nprobe = ((g2 == 2) & (g1 >= 4) & (torch.arange(g1.shape[1], device="cuda")[None, :] < ncon[:, None])).sum(1).float()
    rows.append(torch.stack([d[:, 20], d[:, 24], d[:, 25], d[:, 22], d[:, 22] - prev_nc, nprobe, ts[:, abi.TS_TIMESTEP], d[:, 21], d[:, 2]], 1).cpu().numpy())
    prev_nc = d[:, 22].clone()
X = np.concatenate(rows); X = X[X[:, 6] > 3]
it = X[:, 0]
print("class        n      rebuilds/it  ls_evals/it  ncon   |dncon|  probe contacts  final |grad|  Fz")
for lo, hi in ((0, 5), (5, 7), (7, 9), (9, 12), (12, 16), (16, 100)):
    m = (it >= lo) & (it < hi)
    if m.sum():
        print(f"[{lo:2d},{hi:3d}) {m.sum():8d}   {np.mean(X[m,1]/np.maximum(it[m],1)):.3f}        {np.mean(X[m,2]/np.maximum(it[m],1)):.2f}        {X[m,3].mean():.1f}   {np.abs(X[m,4]).mean():.2f}     {X[m,5].mean():.2f}          {np.median(X[m,7]):.2e}   {X[m,8].mean():.1f}")
