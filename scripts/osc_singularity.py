"""OSC torques near a kinematic singularity: device (Cholesky inverse of J M^-1 J^T, fp32) against the oracle (Jacobi pinv, float64).

robosuite computes Lambda = pinv(J M^-1 J^T).  Away from singularities inverse and pseudo-inverse coincide; this sweep
drives the elbow towards full extension (q4 -> its upper limit, where the arm loses the radial translation) and reports, per pose,
the condition number of J M^-1 J^T and the largest torque deviation, coupled and uncoupled Lambda.  Rigid scene (no contacts)."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from rui_b200 import abi
from rui_b200.env import BatchedUltrasound, packed_model
from oracle import oracle as O

CCF = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
           kp=300, damping_ratio=1, impedance_mode="fixed", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0,
           damping_ratio_limits=[0, 2], uncouple_pos_ori=True, control_delta=True)
Q4 = [-2.0, -1.0, -0.5, -0.3, -0.2, -0.15, -0.1, -0.08, -0.0705]


def sweep(uncouple):
    cc = dict(CCF, uncouple_pos_ori=uncouple)
    n = len(Q4)
    env = BatchedUltrasound(n, soft_torso=False, controller_configs=cc, control_freq=500, seed=1)
    env.reset()
    q, v, w, t = [x.clone() for x in env.get_state()]
    for i, q4 in enumerate(Q4):
        q[i, :7] = torch.tensor([0.0, 0.3, 0.0, q4, 0.0, 1.2, 0.785])
    env.set_state(qpos=q, task=t)
    qn, vn, wn, tn = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
    act = np.tile(np.array([0.3, -0.2, 0.25, 0.1, -0.1, 0.05]), (n, 1))
    rows = []
    orcs = []
    for i in range(n):
        e = O.OracleEnv(packed_model(False), abi.make_config(1, cc, control_freq=500), i)
        e.reset()
        e.set_state(qn[i], vn[i], wn[i], tn[i])
        e.step(act[i])
        orcs.append(e)
    env.step(torch.as_tensor(act, dtype=torch.float32, device="cuda"))
    tau = env.diag()[:, 13:20].cpu().numpy().astype(np.float64)
    for i, e in enumerate(orcs):
        J, _, _ = e.eef()
        A = J @ np.linalg.solve(e.M[:7, :7], J.T)
        rows.append(dict(q4=Q4[i], cond=float(np.linalg.cond(A)), dtau=float(np.abs(tau[i] - e.tau).max()), tau_max=float(np.abs(e.tau).max()),
                         clipped=bool(np.any(np.abs(e.tau) >= 79.99) or np.any(np.abs(e.tau[4:]) >= 11.99))))
    env.close()
    return rows


if __name__ == "__main__":
    out = dict(uncoupled=sweep(True), coupled=sweep(False))
    for k, rows in out.items():
        for r in rows:
            print(k, json.dumps(r))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
