"""CG iterations of the first steps of an episode (cold start: qacc_warmstart = 0 after a reset) against the steady state, in the
staggered episode mix of the benchmark."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from rui_b200.env import BatchedUltrasound
env = BatchedUltrasound(4096, device=0, seed=3, **bench.ENV_OPTS)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(3)
# stagger the episode phases as bench.py does: run 1000 steps first (all envs reset together at 1000), then look at a window
from rui_b200 import abi
q, v, w, t = env.get_state()
t[:, abi.TS_TIMESTEP] = torch.randint(0, 1000, (4096,), device="cuda").float()
env.set_state(task=t)
for s in range(1100):
    env.step(torch.rand(4096, 6, device="cuda", generator=gen), auto_reset=True)
sums = torch.zeros(12, device="cuda"); cnts = torch.zeros(12, device="cuda"); mx = torch.zeros(12, device="cuda")
steady = []
for s in range(1000):
    env.step(torch.rand(4096, 6, device="cuda", generator=gen), auto_reset=True)
    d = env.diag(); ts = env.get_state()[3][:, abi.TS_TIMESTEP].long()  # timestep AFTER the step: k = the k-th step of the episode
    for k in range(1, 11):
        m = ts == k
        if bool(m.any()):
            sums[k] += d[m, 20].sum(); cnts[k] += m.sum(); mx[k] = torch.maximum(mx[k], d[m, 20].max())
    steady.append(float(d[ts > 50, 20].mean()))
print("steady-state mean iterations", sum(steady) / len(steady))
for k in range(1, 11):
    print("episode step", k, "mean iterations %.1f" % float(sums[k] / cnts[k].clamp(min=1)), "max", int(mx[k]), "samples", int(cnts[k]))
