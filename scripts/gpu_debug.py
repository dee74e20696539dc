"""Developer script: stage-by-stage comparison of the CUDA path with the float64 oracle (run under gpurun)."""
import sys, time, os
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rui_b200.env import BatchedUltrasound, packed_model
from rui_b200.abi import make_config
from oracle import oracle as O

np.set_printoptions(precision=5, suppress=True, linewidth=220)
CC_FIXED = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
                kp=300, damping_ratio=1, impedance_mode="fixed", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0, uncouple_pos_ori=True)
CC_TRACK = dict(CC_FIXED, impedance_mode="tracking")


def compare(tag, soft, cc, nsteps, seed=3, n=4, act_fn=None, **kw):
    print(f"===== {tag}: soft={soft} mode={cc['impedance_mode']} steps={nsteps}")
    env = BatchedUltrasound(n, soft_torso=soft, controller_configs=cc, control_freq=500, horizon=1000, seed=seed, **kw)
    pk = packed_model(soft)
    obs = env.reset().cpu().numpy().copy()
    q, v, w, t = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
    orc = []
    for i in range(n):
        cfg = make_config(1, cc, control_freq=500, horizon=1000, seed=seed, **{k: v for k, v in kw.items() if k not in ("solver_iterations", "solver_tolerance")})
        e = O.OracleEnv(pk, cfg, i)
        oobs = e.reset()
        oq, ov, ow, ot = e.get_state()
        if i == 0:
            print("reset: arm q gpu", q[0, :7], "\n       arm q orc", oq[:7])
            print("reset: task gpu", t[0, :18], "\n       task orc", ot[:18])
            print("reset: obs gpu", obs[0], "\n       obs orc", oobs)
        print(f"  env{i} reset diffs: q {np.abs(q[i]-oq).max():.2e} task {np.abs(t[i,:32]-ot[:32]).max():.2e} obs {np.abs(obs[i]-oobs).max():.2e} ncon orc {e.ncon}")
        # continue the oracle from the GPU's reset state so that the trajectories start identical
        e.set_state(q[i], v[i], w[i], t[i])
        orc.append(e)
    d = env.diag().cpu().numpy()
    print("  gpu diag[0]", d[0])
    rng = np.random.default_rng(0)
    lo, hi = env.action_spec
    for s in range(nsteps):
        a = act_fn(s, n) if act_fn else rng.uniform(lo, hi, size=(n, env.action_dim))
        o, r, dn, _ = env.step(torch.as_tensor(a, dtype=torch.float32), auto_reset=False)
        o, r, dn = o.cpu().numpy(), r.cpu().numpy(), dn.cpu().numpy()
        q, v, w, t = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
        d = env.diag().cpu().numpy()
        worst = np.zeros(6)
        for i in range(n):
            oo, orr, od = orc[i].step(a[i])
            oq, ov, ow, ot = orc[i].get_state()
            od_ = orc[i].diag()
            worst = np.maximum(worst, [np.abs(q[i] - oq).max(), np.abs(v[i] - ov).max(), np.abs(o[i, :3] - oo[:3]).max(), abs(r[i] - orr),
                                       np.abs(d[i, 13:20] - od_[13:20]).max(), np.abs(o[i] - oo).max()])
            if i == 0 and (s < 3 or s % 10 == 0 or s == nsteps - 1):
                print(f"  step {s}: ncon gpu {int(d[0,22])} orc {orc[0].ncon} | iters gpu {int(d[0,20])} gnorm {d[0,21]:.2e} orc {orc[0].solver_iter} | Fz gpu {o[0,2]:.4f} orc {oo[2]:.4f} | r gpu {r[0]:.5f} orc {orr:.5f} | done {dn[0]} {od}")
        if s < 3 or s % 10 == 0 or s == nsteps - 1:
            print(f"  step {s}: max|dq| {worst[0]:.2e} max|dv| {worst[1]:.2e} max|dF| {worst[2]:.2e} |dr| {worst[3]:.2e} |dtau| {worst[4]:.2e} |dobs| {worst[5]:.2e}")
    env.close()


def press(s, n):
    a = np.zeros((n, 6))
    a[:, 2] = -1.0
    return a


def sweep():
    """Accuracy vs solver tolerance / iteration cap: rigid press + random, and soft tracking."""
    def run(soft, cc, nsteps, act_fn, **kw):
        n = 2
        env = BatchedUltrasound(n, soft_torso=soft, controller_configs=cc, control_freq=500, horizon=1000, seed=3, **kw)
        pk = packed_model(soft)
        env.reset()
        q, v, w, t = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
        okw = {k: x for k, x in kw.items() if k not in ("solver_iterations", "solver_tolerance")}
        e = O.OracleEnv(pk, make_config(1, cc, control_freq=500, horizon=1000, seed=3, **okw), 0)
        e.reset(); e.set_state(q[0], v[0], w[0], t[0])
        rng = np.random.default_rng(0)
        lo, hi = env.action_spec
        mdq = mdv = mdf = 0.0; its = []
        for s in range(nsteps):
            a = act_fn(s, n, rng, lo, hi)
            o, r, dn, _ = env.step(torch.as_tensor(a, dtype=torch.float32), auto_reset=False)
            oo, orr, od = e.step(a[0])
            q, v, _, _ = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
            oq, ov, _, _ = e.get_state()
            mdq = max(mdq, np.abs(q[0] - oq).max()); mdv = max(mdv, np.abs(v[0] - ov).max())
            mdf = max(mdf, abs(float(o[0, 2]) - oo[2]) / max(1.0, abs(oo[2])))
            its.append(float(env.diag()[0, 20]))
        env.close()
        return mdq, mdv, mdf, np.mean(its), np.max(its)

    def press_then_random(s, n, rng, lo, hi):
        if s < 250:
            a = np.zeros((n, 6)); a[:, 2] = -1; return a
        return np.repeat(rng.uniform(lo, hi, size=(1, 6)), n, 0)

    def rnd(s, n, rng, lo, hi):
        return rng.uniform(lo, hi, size=(n, 6))

    for tol in (1e-5, 3e-6, 1e-6, 3e-7, 1e-7):
        for cap in (20, 40, 80):
            r = run(False, CC_FIXED, 400, press_then_random, solver_tolerance=tol, solver_iterations=cap)
            sft = run(True, CC_TRACK, 60, rnd, solver_tolerance=tol, solver_iterations=cap, torso_solref_randomization=True, initial_probe_pos_randomization=True)
            print(f"tol {tol:.0e} cap {cap}: rigid dq {r[0]:.2e} dv {r[1]:.2e} dF {r[2]:.2e} it {r[3]:.1f}/{r[4]:.0f} | soft dq {sft[0]:.2e} dv {sft[1]:.2e} dF {sft[2]:.2e} it {sft[3]:.1f}/{sft[4]:.0f}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    t0 = time.time()
    if which == "sweep":
        sweep()
    if which in ("all", "rigid"):
        compare("rigid free-space random", False, CC_FIXED, 30)
        compare("rigid press", False, CC_FIXED, 260, act_fn=press)
    if which in ("all", "soft"):
        compare("soft tracking", True, CC_TRACK, 40, torso_solref_randomization=True, initial_probe_pos_randomization=True)
    print("total time", time.time() - t0)
