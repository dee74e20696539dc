"""Soak: 4096 envs x N steps of random actions (rl_config.yaml options, early termination on, auto-reset): divergence and contact-overflow
counters, finiteness and ranges of everything returned, episode statistics."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rui_b200.env import BatchedUltrasound
CC = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
          kp=300, damping_ratio=1, impedance_mode="tracking", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0,
          damping_ratio_limits=[0, 2], uncouple_pos_ori=True, control_delta=True)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
out = {}
for et in (False, True):
    env = BatchedUltrasound(4096, controller_configs=CC, control_freq=500, horizon=1000, early_termination=et, torso_solref_randomization=True,
                            initial_probe_pos_randomization=True, seed=11)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(5)
    bad, ndone, rsum, omax, itmax = 0, 0, 0.0, 0.0, 0
    for s in range(steps):
        o, r, d, t = env.step(torch.rand(4096, 6, device="cuda", generator=gen), auto_reset=True)
        if s % 20 == 0:
            bad += int((~torch.isfinite(o)).sum() + (~torch.isfinite(r)).sum())
            bad += int(((r < 0) | (r > 12.0001)).sum())
            omax = max(omax, float(o[:, :3].abs().max()))
            itmax = max(itmax, int(env.diag()[:, 20].max()))
        ndone += int(d.sum()); rsum += float(r.mean())
    out["early_termination" if et else "horizon_only"] = dict(steps=steps, env_steps=steps * 4096, episodes=ndone, mean_reward=rsum / steps, non_finite_or_out_of_range=bad,
                                                             max_abs_force=omax, max_cg_iterations=itmax, divergence_count=env.divergence_count,
                                                             contact_overflow_count=env.contact_overflow_count)
    env.close()
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
