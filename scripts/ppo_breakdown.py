"""Where a PPO iteration (BASELINE config 5, one GPU) spends its time: rollout vs update (CUDA events), and the kernels of the update
(torch profiler, one epoch)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rui_b200.env import BatchedUltrasound
from rui_b200.ppo import PPO
CC = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
          kp=300, damping_ratio=1, impedance_mode="tracking", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0,
          damping_ratio_limits=[0, 2], uncouple_pos_ori=True, control_delta=True)
N = int(os.environ.get("ENVS", 8192))
env = BatchedUltrasound(N, controller_configs=CC, control_freq=500, horizon=1000, early_termination=True, torso_solref_randomization=True,
                        initial_probe_pos_randomization=True, seed=3)
m = PPO(env, n_steps=32, seed=1)
m._setup()
for _ in range(2):
    m.train(m.collect_rollouts())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tr, tu, cpu_r, cpu_u = 0.0, 0.0, 0.0, 0.0
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev[0].record(); b = m.collect_rollouts(); ev[1].record(); t1 = time.perf_counter()
    m.train(b); ev[2].record(); t2 = time.perf_counter(); torch.cuda.synchronize()
    tr += ev[0].elapsed_time(ev[1]); tu += ev[1].elapsed_time(ev[2]); cpu_r += t1 - t0; cpu_u += t2 - t1
print(f"rollout {tr/3:.1f} ms (cpu enqueue {1e3*cpu_r/3:.1f})  update {tu/3:.1f} ms (cpu enqueue {1e3*cpu_u/3:.1f}) per iteration; {tu/3/320*1e3:.0f} us per minibatch step")
b = m.collect_rollouts()
m.n_epochs = 1
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.train(b); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
