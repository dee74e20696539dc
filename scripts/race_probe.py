import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rui_b200.env import BatchedUltrasound
CC = dict(type="OSC_POSE", impedance_mode="tracking", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0, uncouple_pos_ori=True)
# small workload for compute-sanitizer: reset, steps with auto-reset across a horizon (slot take-over + prepare launches on the side
# stream), an explicit masked reset, a host-buffer step
env = BatchedUltrasound(3, controller_configs=CC, control_freq=500, horizon=2, torso_solref_randomization=True, initial_probe_pos_randomization=True,
                        seed=3, solver_iterations=6)
o = env.reset()
for s in range(5):
    o, r, d, _ = env.step(torch.full((3, 6), 0.5), auto_reset=True)
env.reset(torch.tensor([1, 0, 1], dtype=torch.uint8))
import numpy as np
env.step_host(np.full((3, 6), 0.4, np.float32))
torch.cuda.synchronize()
print("obs", o[0, :3].tolist(), "iters", env.diag()[0, 20].item(), "episodes", env.get_state()[3][:, 25].tolist())
