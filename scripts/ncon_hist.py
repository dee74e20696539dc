"""Distribution of the per-env contact count (uncapped, diag[22]) over full random-action episodes (sizing of the device contact cap)."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rui_b200.env import BatchedUltrasound
CCF = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
           kp=300, damping_ratio=1, impedance_mode="fixed", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0,
           damping_ratio_limits=[0, 2], uncouple_pos_ori=True, control_delta=True)
for name, cc, lo in (("tracking", dict(CCF, impedance_mode="tracking"), 0.0), ("fixed", CCF, -1.0)):
    env = BatchedUltrasound(4096, controller_configs=cc, control_freq=500, seed=3, torso_solref_randomization=True, initial_probe_pos_randomization=True)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(11)
    hist = torch.zeros(257, device="cuda")
    for s in range(1000):
        env.step(lo + (1 - lo) * torch.rand(4096, 6, device="cuda", generator=gen), auto_reset=True)
        c = env.diag()[:, 22].long().clamp(0, 256)
        hist += torch.bincount(c, minlength=257).float()
    h = hist.cpu().numpy()
    nz = [(i, int(v)) for i, v in enumerate(h) if v > 0]
    cum = h.cumsum() / h.sum()
    print(name, "max", nz[-1][0], "min", nz[0][0], "p50", int((cum >= 0.5).argmax()), "p99", int((cum >= 0.99).argmax()), "p9999", int((cum >= 0.9999).argmax()),
          "tail", nz[-8:], "overflow", env.contact_overflow_count)
    env.close()
