"""A/B of the two arm kernels on the GPU: thread per env (USIM_ARM_THREAD=1) vs lane per link (default).
Same seed, same actions: prints the largest deviation per field of the arm record after the first step (identical inputs) and of
the state / observations after a rollout, for the Panda (modes tracking, fixed) and the UR5e."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from rui_b200.env import BatchedUltrasound

FIELDS = dict(M=(0, 49), qfrc_smooth=(49, 56), tau=(56, 63), Jsite=(63, 105), Jhand=(105, 126), site=(126, 129), Rsite=(129, 138),
              ptip=(138, 141), pback=(141, 144), tau0=(144, 147), Jft=(147, 168), quat=(168, 172), dx=(172, 179))
CCF = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
           kp=300, damping_ratio=1, impedance_mode="fixed", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0,
           damping_ratio_limits=[0, 2], uncouple_pos_ori=False, control_delta=True)
CCT = dict(CCF, impedance_mode="tracking", uncouple_pos_ori=True)


def run(thread, cc, n=256, steps=40, **kw):
    os.environ["USIM_ARM_THREAD"] = "1" if thread else "0"
    env = BatchedUltrasound(n, controller_configs=cc, control_freq=kw.pop("control_freq", 500), seed=5, torso_solref_randomization=True,
                            initial_probe_pos_randomization=True, **kw)
    env.reset()
    g = torch.Generator(device="cpu").manual_seed(1)
    lo = 0.0 if cc["impedance_mode"] != "fixed" else -1.0
    acts = (lo + (1 - lo) * torch.rand(steps, n, env.action_dim, generator=g)).cuda()
    o, r, d, _ = env.step(acts[0])
    rec1 = env.arm_record().clone()
    for s in range(1, steps):
        o, r, d, _ = env.step(acts[s])
    q, v, w, t = env.get_state()
    out = dict(rec1=rec1, rec=env.arm_record().clone(), obs=o.clone(), q=q.clone(), v=v.clone(), rew=r.clone())
    torch.cuda.synchronize()
    env.close()
    return out


def report(name, cc, **kw):
    a, b = run(True, cc, **kw), run(False, cc, **kw)
    res = {}
    for f, (i0, i1) in FIELDS.items():
        x, y = a["rec1"][:, i0:i1], b["rec1"][:, i0:i1]
        res[f] = [float((x - y).abs().max()), float(x.abs().max())]
    res["rollout"] = {k: float((a[k] - b[k]).abs().max()) for k in ("obs", "q", "v", "rew")}
    print(name, json.dumps(res))
    return res


if __name__ == "__main__":
    report("panda_tracking", CCT)
    report("panda_fixed_coupled", CCF)
    report("panda_tracking_20hz", CCT, control_freq=20, steps=6)
    from rui_b200.model import ur5e_params
    report("ur5e_tracking", CCT, scene_params=ur5e_params())
    report("rigid_fixed", CCF, soft_torso=False)
    # timing of one step with each kernel, 4096 envs
    for thread in (1, 0):
        os.environ["USIM_ARM_THREAD"] = str(thread)
        env = BatchedUltrasound(4096, controller_configs=CCT, control_freq=500, seed=5)
        env.reset()
        a = torch.rand(4096, 6, device="cuda")
        for _ in range(30):
            env.step(a, auto_reset=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            env.step(a, auto_reset=True)
        e1.record()
        torch.cuda.synchronize()
        print("arm_thread" if thread else "arm_warp", "ms/step", e0.elapsed_time(e1) / 200)
        env.close()
