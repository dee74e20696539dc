"""Calibration sweep of the substituted probe geometry (A-PROBE-1) against the reference's post-reset statistics [ART]."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rui_b200.env import BatchedUltrasound
from rui_b200.model import SceneParams
CC = dict(type="OSC_POSE", impedance_mode="tracking", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0, uncouple_pos_ori=True)
art = json.load(open(os.path.join(ROOT, "tests", "golden", "art_stats.json")))["tracking"]
aro = np.array(art["last_original_obs"])
print("ART: in-contact %.3f  Fz median %.2f  Fz max %.1f" % ((aro[:, 2] > 0).mean(), np.median(aro[aro[:, 2] > 0, 2]), aro[:, 2].max()))
for r in (0.02, 0.03, 0.04, 0.05, 0.06):
    for back in (-0.10,):
        sp = SceneParams(probe_radius=r, probe_tip_z=-r, probe_back_z=min(back, -r - 0.02))
        env = BatchedUltrasound(2048, device=0, controller_configs=CC, control_freq=500, horizon=1000, torso_solref_randomization=True,
                                initial_probe_pos_randomization=True, seed=3, scene_params=sp)
        o = env.reset().cpu().numpy()
        ncon = env.contacts()[0].float().mean().item()
        inc = o[:, 2] > 0
        print(f"r={r:.3f}: in-contact {inc.mean():.3f}  Fz median {np.median(o[inc, 2]):.2f}  Fz mean {o[inc,2].mean():.2f} Fz max {o[:, 2].max():.1f}  mean ncon {ncon:.1f}")
        env.close()
