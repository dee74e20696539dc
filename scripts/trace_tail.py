"""Timeline of one step launch (USIM_TRACE): when do the SMs run dry?  Per-env (start, end, SM) of the last launch of a staggered
4096-env run -> launch duration, mean resident CTAs over time, idle slot-time in the tail, duration of an env by CG iterations."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
path = "/tmp/usim_trace.bin"
os.environ["USIM_TRACE"] = path
os.environ["USIM_LIB"] = os.path.join(ROOT, "build", "libusim_trace.so")  # built with: bash scripts/build_variant.sh trace -DUSIM_TRACE=1
import torch
import bench
from rui_b200 import abi
from rui_b200.env import BatchedUltrasound
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedUltrasound(n, device=0, seed=3, **bench.ENV_OPTS)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(3)
q, v, w, t = env.get_state()
t[:, abi.TS_TIMESTEP] = torch.randint(0, 1000, (n,), device="cuda").float()
env.set_state(task=t)
for s in range(1200):
    env.step(torch.rand(n, 6, device="cuda", generator=gen), auto_reset=True)
prev_iters = env.diag()[:, 20].cpu().numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush.zero_(); torch.cuda.synchronize()
env.step(torch.rand(n, 6, device="cuda", generator=gen), auto_reset=True)
torch.cuda.synchronize()
iters = env.diag()[:, 20].cpu().numpy()
env.close()
tr = np.fromfile(path, dtype=np.uint64).reshape(n, 3).astype(np.int64)
t0, t1, sm = tr[:, 0], tr[:, 1], tr[:, 2]
ok = t1 > 0
base = t0[ok].min()
t0, t1 = (t0 - base) / 1e3, (t1 - base) / 1e3  # us
end = t1[ok].max()
dur = (t1 - t0)[ok]
print(f"launch: {ok.sum()} envs on {len(np.unique(sm[ok]))} SMs, {end:.1f} us from first start to last end; env duration mean {dur.mean():.1f} us, p10 {np.percentile(dur,10):.1f}, p90 {np.percentile(dur,90):.1f}, max {dur.max():.1f}")
# resident CTAs over time
grid = np.linspace(0, end, 41)
res = [(np.sum((t0[ok] <= g) & (t1[ok] > g))) for g in grid]
print("resident CTAs at 0..100% of the launch (2.5% steps):", res)
slots = 8 * len(np.unique(sm[ok]))
busy = dur.sum()
print(f"slot-time used {busy/ (slots*end):.3f} of {slots} slots x {end:.1f} us")
last_start = t0[ok].max()
print(f"last CTA starts at {last_start:.1f} us ({last_start/end:.2f} of the launch); per-SM finish time: min {min(t1[ok][sm[ok]==k].max() for k in np.unique(sm[ok])):.1f} median {np.median([t1[ok][sm[ok]==k].max() for k in np.unique(sm[ok])]):.1f} max {end:.1f}")
for lo, hi in ((0, 4), (4, 6), (6, 8), (8, 12), (12, 99)):
    m = ok & (iters >= lo) & (iters < hi)
    if m.any():
        print(f"iterations [{lo},{hi}): {m.sum()} envs, duration mean {(t1 - t0)[m].mean():.1f} us; started at {t0[m].mean():.1f} us on average")
fin = np.array([t1[ok][sm[ok] == k].max() for k in np.unique(sm[ok])])
print("per-SM finish time percentiles (us): p10 %.0f p50 %.0f p75 %.0f p90 %.0f p95 %.0f p99 %.0f max %.0f" % tuple(np.percentile(fin, [10, 50, 75, 90, 95, 99, 100])))
order = np.argsort(-t1)
print("the 12 envs that finish last: (end us, start us, duration us, iterations, iterations in the previous step)")
for e in order[:12]:
    print("   %.0f  %.0f  %.0f  %d  %d" % (t1[e], t0[e], t1[e] - t0[e], iters[e], prev_iters[e]))
late_long = ok & (t0 > 250) & (iters >= 8)
print("envs with >= 8 iterations started after 250 us:", int(late_long.sum()), "of", int((ok & (iters >= 8)).sum()), "; their previous-step iterations: mean %.1f" % prev_iters[late_long].mean() if late_long.any() else "")
early = ok & (t0 < 0.25 * end); late = ok & (t0 > 0.6 * end)
print(f"duration per CG pass: started in the first quarter {((t1-t0)[early] / (iters[early] + 1)).mean():.2f} us, in the last quarter {((t1-t0)[late] / (iters[late] + 1)).mean():.2f} us")
