#!/usr/bin/env python
"""CPU analysis (numpy, float64; no GPU): spectrum of  P^-1 H  for the device solver's preconditioner.

H = M + J^T D J is the Hessian of MuJoCo's primal problem at a post-reset state of the soft scene (active contacts taken in their
quadratic zone), assembled densely in the unknown ordering of the CUDA kernel; P^-1 is the kernel's preconditioner (csrc/soft.cuh:
7x7 arm block, arrow torso block with the exact 6x6 Schur complement over the slider diagonal, slider block D - W inverted by a
polynomial in N = D^-1 W), restated with dense linear algebra.  Prints the condition number for the polynomial variants the
optimisation log compares (Jacobi, second order, cubic Neumann, cubic Chebyshev) -- CG iterations scale with sqrt(kappa) -- and for
the arm <-> torso coupled dense block the kernel does not have yet.

  python scripts/precond_spectrum.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402  (analysis tool: test infrastructure, not the product path)
from rui_b200 import abi  # noqa: E402
from rui_b200.env import packed_model  # noqa: E402

CC_TRACK = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
                kp=300, damping_ratio=1, impedance_mode="tracking", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0, uncouple_pos_ori=True)


def skew(r):
    return np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])


def quat2mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def assemble(seed=5, steps=5):
    pk = packed_model(True)
    m = pk.model
    P, A = m.params, m.arrays
    cfg = abi.make_config(1, CC_TRACK, control_freq=500, horizon=1000, seed=seed, torso_solref_randomization=True,
                          initial_probe_pos_randomization=True)
    e = O.OracleEnv(pk, cfg, 0)
    e.reset()
    rng = np.random.default_rng(1)
    for _ in range(steps):
        e.step(rng.uniform(0, 1, 6))
    q, v, _, ts = e.get_state()
    e.forward(e.tau)
    M, c = e.M, e.contacts()
    dmin, dmax, width, mid, power = P.solimp

    def imped(pos):
        x = abs(pos) / width
        y = 1.0 if x >= 1 else (x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1))
        return min(max(dmin + (0.0 if x == 0 else y) * (dmax - dmin), 1e-4), 0.9999)

    R, Pt = quat2mat(q[10:14]), q[7:10]
    ax, ppos, iw = np.asarray(A["part_axis"]), np.asarray(A["part_pos"]), np.asarray(A["dof_invweight0"])
    biw = np.asarray(A["body_invweight0"])
    qs = q[14:]
    n = 283
    H = M.copy()
    Dfix = np.array([1 / ((1 - imped(x)) / imped(x) * iw[13 + i]) for i, x in enumerate(qs)])
    H[13:, 13:] += np.diag(Dfix)
    pairs = np.asarray(A["eq_pairs"])
    Dp = np.array([1 / ((1 - imped(qs[a] - qs[b])) / imped(qs[a] - qs[b]) * (iw[13 + a] + iw[13 + b])) for a, b in pairs])
    W = np.zeros((270, 270))
    for (a, b), d in zip(pairs, Dp):
        W[a, b] += d
        W[b, a] += d
    H[13:, 13:] += np.diag(W.sum(1)) - W
    it = imped(qs.sum())
    Dt = 1 / ((1 - it) / it * float(np.asarray(A["tendon_invweight0"])[0]))
    H[13:, 13:] += Dt * np.ones((270, 270))
    J, spos, _ = e.eef()
    Ka = np.zeros((7, 7))            # arm block of the preconditioner
    Kf = np.zeros((6, 6))            # free-body block (world force, body torque)
    sk = np.zeros((270, 3))          # per slider: sum of K a_w over its contacts
    dgc = np.zeros(270)
    ids = A["ids"]
    for k in range(len(c["dist"])):
        if np.abs(c["force"][k]).max() == 0:
            continue  # top zone: inactive
        nrm, pos = c["frame"][k][0], c["pos"][k]
        g1, g2 = c["geom1"][k], c["geom2"][k]
        diagA = (biw[ids[6] + (g1 - 4), 0] if g1 >= 4 else 0.0) + (biw[ids[4], 0] if g2 == 2 else 0.0)
        imp = imped(c["dist"][k])
        Dn = 1 / ((1 - imp) / imp * diagA)
        K = Dn * P.impratio * np.eye(3) + (Dn - Dn * P.impratio) * np.outer(nrm, nrm)  # quadratic zone
        Jk = np.zeros((3, n))
        if g2 == 2:
            Jc = J[:3] + np.cross(J[3:].T, pos - spos).T
            Jk[:, :7] += Jc
            Ka += Jc.T @ K @ Jc
        if g1 >= 4:
            i, r = g1 - 4, pos - Pt
            aw = R @ ax[i]
            Jk[:, 7:10] -= np.eye(3)
            Jk[:, 10:13] -= -skew(r) @ R
            Jk[:, 13 + i] -= aw
            Af = np.hstack([np.eye(3), -skew(r) @ R])
            Kf += Af.T @ K @ Af
            sk[i] += K @ aw
            dgc[i] += aw @ K @ aw
        H += Jk.T @ K @ Jk
    return dict(H=H, M=M, R=R, Pt=Pt, ax=ax, ppos=ppos, qs=qs, W=W, Dfix=Dfix, Dt=Dt, Ka=Ka, Kf=Kf, sk=sk, dgc=dgc, cap_r=P.cap_radius,
                ncon=len(c["dist"]))


def precond_inverse(s, order, c):
    """Dense P^-1 of the kernel's preconditioner.  order 0: Jacobi slider block, 1: D^-1 (I + N), 2: D^-1 (I + c (N + N^2))."""
    H, M, R, ax = s["H"], s["M"], s["R"], s["ax"]
    n = 283
    Pa = M[:7, :7] + s["Ka"]
    D = np.diag(M)[13:] + s["Dfix"] + s["W"].sum(1) + s["Dt"] + s["dgc"]
    aw = ax @ R.T
    cr = (s["ppos"] + (s["qs"] - s["cap_r"])[:, None] * ax) @ R.T       # lever the kernel uses: outer end of the capsule, world frame
    B = np.hstack([M[13:, 7:10] + s["sk"], np.cross(cr, s["sk"]) @ R])   # [270, 6]: translation (world), rotation (body frame)
    Aff = M[7:13, 7:13] + s["Kf"]
    Sf = Aff - B.T @ (B / D[:, None])
    N = s["W"] / D[:, None]
    Q = {0: np.eye(270), 1: np.eye(270) + N, 2: np.eye(270) + c * (N + N @ N)}[order] / D[None, :]
    Pinv = np.zeros((n, n))
    Pinv[:7, :7] = np.linalg.inv(Pa)
    Sfi = np.linalg.inv(Sf)
    DB = B / D[:, None]
    Pinv[7:13, 7:13] = Sfi
    Pinv[7:13, 13:] = -Sfi @ DB.T
    Pinv[13:, 7:13] = -DB @ Sfi
    Pinv[13:, 13:] = Q + DB @ Sfi @ DB.T
    return Pinv


def coupled_inverse(s, order, c):
    """What the kernel does NOT do yet: arm (7) and free body (6) as ONE dense 13x13 block, every slider eliminated against all 13
    (the probe contacts couple the arm to the sliders they touch and, through them, to the torso)."""
    H = s["H"]
    D = np.diag(H)[13:].copy()
    N = s["W"] / D[:, None]
    Q = {0: np.eye(270), 1: np.eye(270) + N, 2: np.eye(270) + c * (N + N @ N)}[order] / D[None, :]
    B = H[13:, :13]
    DB = B / D[:, None]
    Si = np.linalg.inv(H[:13, :13] - B.T @ DB)
    Pinv = np.zeros_like(H)
    Pinv[:13, :13], Pinv[:13, 13:], Pinv[13:, :13], Pinv[13:, 13:] = Si, -Si @ DB.T, -DB @ Si, Q + DB @ Si @ DB.T
    return Pinv


def kappa(s, order, c=1.6, coupled=False):
    Pinv, H = (coupled_inverse if coupled else precond_inverse)(s, order, c), s["H"]
    assert np.abs(Pinv - Pinv.T).max() < 1e-9 * np.abs(Pinv).max()
    L = np.linalg.cholesky(Pinv + 0)           # P^-1 is symmetric positive definite
    ev = np.linalg.eigvalsh(L.T @ H @ L)
    return ev[-1] / ev[0], ev


if __name__ == "__main__":
    st = assemble()
    rho = np.abs(np.linalg.eigvals(st["W"] / (np.diag(st["M"])[13:] + st["Dfix"] + st["W"].sum(1) + st["Dt"] + st["dgc"])[:, None])).max()
    print(f"contacts {st['ncon']}, spectral radius of N = D^-1 W: {rho:.3f}  ->  c = 4/(4 - 3 rho^2) = {4 / (4 - 3 * rho * rho):.2f}")
    k_raw = np.linalg.cond(st["H"])
    print(f"kappa(H) = {k_raw:.3g}")
    for name, order, c in (("Jacobi slider block", 0, 0), ("second order D^-1 (I + N)", 1, 0), ("cubic Neumann c = 1", 2, 1.0),
                           ("cubic Chebyshev c = 1.6", 2, 1.6), ("cubic c = 2.4", 2, 2.4)):
        k, ev = kappa(st, order, c)
        print(f"{name:28s} kappa(P^-1 H) = {k:6.2f}   sqrt = {np.sqrt(k):5.2f}   eigenvalues in [{ev[0]:.3f}, {ev[-1]:.3f}], {np.sum(ev > 1.5 * np.median(ev))} above 1.5 x median")
    print("arm and free body as one dense 13x13 block, sliders eliminated against all 13 (not implemented on the device yet):")
    for name, order, c in (("Jacobi slider block", 0, 0), ("second order D^-1 (I + N)", 1, 0), ("cubic Chebyshev c = 1.4", 2, 1.4)):
        k, ev = kappa(st, order, c, coupled=True)
        print(f"{name:28s} kappa(P^-1 H) = {k:6.2f}   sqrt = {np.sqrt(k):5.2f}   eigenvalues in [{ev[0]:.3f}, {ev[-1]:.3f}]")
