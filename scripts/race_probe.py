import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rui_b200.env import BatchedUltrasound
CC = dict(type="OSC_POSE", impedance_mode="tracking", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0, uncouple_pos_ori=True)
env = BatchedUltrasound(2, controller_configs=CC, control_freq=500, torso_solref_randomization=True, initial_probe_pos_randomization=True, seed=3, solver_iterations=6)
o = env.reset()
for s in range(2):
    o, r, d, _ = env.step(torch.full((2, 6), 0.5))
torch.cuda.synchronize()
print("obs", o[0, :3].tolist(), "iters", env.diag()[0, 20].item())
