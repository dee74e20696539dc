"""How often does a solve reach the iteration cap?  Full-episode mix (4096 envs x 2000 steps, random actions, auto-reset): histogram of
CG iterations per env-step, residual (relative gradient) of the capped solves."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from rui_b200.env import BatchedUltrasound
env = BatchedUltrasound(4096, device=0, seed=3, **bench.ENV_OPTS)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(3)
hist = torch.zeros(64, device="cuda")
capped_grad = []
ts_at_cap = []
for s in range(2000):
    env.step(torch.rand(4096, 6, device="cuda", generator=gen), auto_reset=True)
    d = env.diag()
    it = d[:, 20].long().clamp(0, 63)
    hist += torch.bincount(it, minlength=64).float()
    m = it >= 40
    if bool(m.any()):
        capped_grad.append(d[m, 21].cpu())
        ts_at_cap.append(env.get_state()[3][m, 14].cpu())
h = hist.cpu().numpy(); tot = h.sum()
print("mean", float((h * range(64)).sum() / tot), "P(>=12)", float(h[12:].sum() / tot), "P(>=20)", float(h[20:].sum() / tot), "P(>=40)", float(h[40:].sum() / tot), "count>=40", int(h[40:].sum()))
if capped_grad:
    g = torch.cat(capped_grad); t = torch.cat(ts_at_cap)
    print("capped solves: gradient norm median %.3g max %.3g; episode step of those: median %d, fraction within 3 steps of a reset %.2f" % (float(g.median()), float(g.max()), int(t.median()), float((t <= 3).float().mean())))
