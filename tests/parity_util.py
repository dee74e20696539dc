"""Step-by-step comparison of a batch of CUDA envs with float64 oracle envs started from IDENTICAL states.

Used by tests/test_gpu_parity.py (asserts the stated tolerances) and scripts/parity_report.py (measures the maxima the
tolerances are derived from; profiles/r02_parity_drift.json).  Test infrastructure: the only place the oracle meets the GPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from rui_b200 import abi

# channel groups of the 19-float observation (ultrasound.py:363-401)
OBS_GROUPS = {
    "force_xy": slice(0, 2), "force_z": slice(2, 3), "torque": slice(3, 6), "eef_vel": slice(6, 9), "fz_mean": slice(9, 10),
    "dfz": slice(10, 11), "vel_mean": slice(11, 12), "pos_err": slice(12, 15), "quat_err": slice(15, 19),
}


def make_oracles(O, env, cc, n=None, soft=True, **kw):
    """Oracle envs continuing from the GPU env's current state (north_star: identical initial states)."""
    from rui_b200.env import packed_model
    drop = ("solver_iterations", "solver_tolerance", "precond_rebuilds", "scene_params")
    okw = {k: v for k, v in kw.items() if k not in drop}
    okw.setdefault("control_freq", 500)
    pk = env.packed if kw.get("scene_params") is not None else packed_model(soft)
    n = env.num_envs if n is None else n
    q, v, w, t = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
    out = []
    for i in range(n):
        e = O.OracleEnv(pk, abi.make_config(1, cc, **okw), i)
        e.reset()
        e.set_state(q[i], v[i], w[i], t[i])
        out.append(e)
    return out


def termination_flags(q7, ts, jnt_range, horizon):
    """The OR-ed conditions of _check_terminated (ultrasound.py:635-670) + the horizon, evaluated on a (qpos, task record) pair:
    (joint limit, trajectory deviation, orientation in contact, lost contact, horizon)."""
    lo, hi = jnt_range[:7, 0], jnt_range[:7, 1]
    qlim = bool(np.any(~((lo + 0.1 < q7) & (q7 < hi - 0.1))))
    pe = float(np.hypot(ts[abi.TS_POS_ERR], ts[abi.TS_POS_ERR + 1]))
    inc, touched = bool(ts[abi.TS_IN_CONTACT]), bool(ts[abi.TS_TOUCHED])
    return (qlim, pe > 1.0, inc and ts[abi.TS_ORI_ERR] > 0.10, touched and not inc, ts[abi.TS_TIMESTEP] >= horizon)


def termination_margin(q7, ts, jnt_range):
    """distance of every thresholded quantity from its threshold (a fp32/fp64 flip is legitimate only inside round-off of one)"""
    lo, hi = jnt_range[:7, 0], jnt_range[:7, 1]
    m = [np.abs(q7 - lo - 0.1).min(), np.abs(hi - 0.1 - q7).min(), abs(np.hypot(ts[abi.TS_POS_ERR], ts[abi.TS_POS_ERR + 1]) - 1.0),
         abs(ts[abi.TS_ORI_ERR] - 0.10)]
    return float(min(m))


class Drift:
    """running maxima of |gpu - oracle| per quantity"""

    def __init__(self):
        self.max = {}

    def add(self, name, val):
        val = float(val)
        if not np.isfinite(val):
            val = float("inf")
        self.max[name] = max(self.max.get(name, 0.0), val)

    def __getitem__(self, k):
        return self.max.get(k, 0.0)


def compare_rollout(O, env, orcs, actions, *, threads=None, check_contacts=True, on_step=None, settle=5):
    """Step env (auto_reset off) and the oracles with the same actions [steps][n][adim]; returns (Drift, log).

    Compared per step, for every env still running on both sides: qpos, qvel, all 19 observation channels, reward, done,
    the task record entries, contact-pair lists.  Force-like channels are reported as absolute AND relative deviations
    (relative to max(1 N, |oracle|)).

    The two trajectories run freely from identical initial states, so they are NOT in identical states later (fp32 drift ~1e-5 m):
    a contact whose distance crosses zero within that drift appears one step apart on the two sides.  The contact's damping force
    is there from its first step (aref = -B v), so for that step -- and while the kick decays, `settle` steps -- forces differ by
    O(B v m) although both sides are right.  Such env-steps are the ones where the two contact lists differ (only in pairs within
    2e-6 m of the threshold: anything else is logged as a mismatch); they are counted (`log["threshold_env_steps"]`) and kept out of
    the returned Drift, which covers every other env-step; `log["drift_all"]` covers all of them."""
    n = len(orcs)
    dr, dr_all = Drift(), Drift()
    log = {"done_mismatch": [], "contact_mismatch": [], "steps": 0, "env_steps": 0, "terminated": np.zeros(n, bool),
           "threshold_env_steps": 0, "threshold_events": 0, "drift_all": dr_all}
    alive = np.ones(n, bool)
    last_event = np.full(n, -10 ** 9)
    jr = np.asarray(env.model.g_jnt_range, dtype=np.float64)
    for s, a in enumerate(actions):
        if not alive.any():
            break
        o, r, d, _ = env.step(torch.as_tensor(np.asarray(a), dtype=torch.float32), auto_reset=False)
        q, v, _, t = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
        o, r, d = o.cpu().numpy().astype(np.float64)[:n], r.cpu().numpy().astype(np.float64)[:n], d.cpu().numpy().astype(bool)[:n]
        idx = np.nonzero(alive)[0]
        oo, orr, od = np.zeros((n, abi.OBS_DIM)), np.zeros(n), np.zeros(n, bool)
        oo[idx], orr[idx], od[idx], stepped = O.step_batch([orcs[i] for i in idx], np.asarray(a, dtype=np.float64)[idx], threads)
        assert stepped.all(), "oracle / bookkeeping out of sync"
        if check_contacts:
            ncon, g1, g2, _ = [x.cpu().numpy() for x in env.contacts()]
        for i in idx:
            oq, ov, _, ot = orcs[i].get_state()
            if check_contacts:
                k = int(ncon[i])
                got = list(zip(g1[i, :k].tolist(), g2[i, :k].tolist()))
                c = orcs[i].contacts()
                want = list(zip(c["geom1"].tolist(), c["geom2"].tolist()))
                if got != want:
                    last_event[i] = s
                    log["threshold_events"] += 1
                    dist = {p: dd for p, dd in zip(want, c["dist"])}
                    bad = [p for p in set(got) ^ set(want) if abs(dist.get(p, 0.0)) > 2e-6]
                    order_ok = [p for p in got if p in dist] == [p for p in want if p in set(got)]
                    if bad or not order_ok or len([p for p in got if p not in dist]) > 2:
                        log["contact_mismatch"].append(dict(step=s, env=int(i), bad=bad))
            clean = s - last_event[i] > settle
            log["threshold_env_steps"] += not clean
            for D in ((dr, dr_all) if clean else (dr_all,)):
                D.add("qpos", np.abs(q[i] - oq).max())
                D.add("qvel", np.abs(v[i] - ov).max())
                D.add("reward", abs(r[i] - orr[i]))
                for name, sl in OBS_GROUPS.items():
                    D.add("obs_" + name, np.abs(o[i, sl] - oo[i, sl]).max())
                D.add("force_rel", (np.abs(o[i, 0:3] - oo[i, 0:3]) / np.maximum(1.0, np.abs(oo[i, 0:3]).max())).max())
                D.add("torque_rel", (np.abs(o[i, 3:6] - oo[i, 3:6]) / np.maximum(0.1, np.abs(oo[i, 3:6]).max())).max())
                D.add("fz_mean_rel", abs(o[i, 9] - oo[i, 9]) / max(1.0, abs(oo[i, 9] + 5.0)))
                D.add("dfz_rel", abs(o[i, 10] - oo[i, 10]) / max(500.0, abs(oo[i, 10])))  # dFz = dF * control_freq: 1 N of force = 500 units
                for name, kk, wd in (("traj_pt", abi.TS_TRAJ_PT, 3), ("vel_mean", abi.TS_VEL_MEAN, 1), ("fz_mean", abi.TS_FZ_MEAN, 1),
                                     ("pos_err", abi.TS_POS_ERR, 2), ("ori_err", abi.TS_ORI_ERR, 1)):
                    D.add("ts_" + name, np.abs(t[i, kk:kk + wd] - ot[kk:kk + wd]).max())
            exact = (t[i, abi.TS_TOUCHED] == ot[abi.TS_TOUCHED] and t[i, abi.TS_TIMESTEP] == ot[abi.TS_TIMESTEP] and
                     t[i, abi.TS_IN_CONTACT] == ot[abi.TS_IN_CONTACT])
            fg = termination_flags(q[i, :7], t[i], jr, env.horizon)
            fo = termination_flags(oq[:7], ot, jr, env.horizon)
            if bool(d[i]) != bool(od[i]) or fg != fo or not exact:
                log["done_mismatch"].append(dict(step=s, env=int(i), gpu=(bool(d[i]),) + fg, oracle=(bool(od[i]),) + fo,
                                                 margin=termination_margin(oq[:7], ot, jr)))
            log["env_steps"] += 1
            if d[i] or od[i]:
                alive[i] = False
                log["terminated"][i] = True
        log["steps"] = s + 1
        if on_step is not None:
            on_step(s, dr)
    return dr, log
