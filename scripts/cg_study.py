#!/usr/bin/env python
"""CPU study (numpy, float64; no GPU) of the device solver's iteration count.

Assembles MuJoCo's primal problem  min_a 1/2 (a-a0)^T M (a-a0) + sum_eq 1/2 D (Ja-aref)^2 + sum_contacts cone(Ja-aref)  of a soft-scene
state from the oracle's data (the formulation of tests/test_oracle_physics.py::test_soft_scene_solution_satisfies_the_optimality_
conditions) and runs the kernel's algorithm -- nonlinear CG, Polak-Ribiere+, exact line search, preconditioner rebuilt when contact
zones change -- with different preconditioners, from the warm start the device uses (previous qacc), to the device's stopping rule.
Answers: how many iterations are due to conditioning, how many to the changing active set?

  python scripts/cg_study.py [--states 6]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import oracle as O  # noqa: E402
from rui_b200 import abi  # noqa: E402
from rui_b200.env import packed_model  # noqa: E402
from precond_spectrum import CC_TRACK, quat2mat, skew  # noqa: E402


class Problem:
    def __init__(self, e, pk):
        m = pk.model
        P, A = m.params, m.arrays
        q, v, warm, ts = e.get_state()
        e.forward(e.tau)
        self.M, self.a0, self.a_star, self.warm = e.M, e.qacc_smooth, e.qacc, warm
        c = e.contacts()
        dmin, dmax, width, mid, power = P.solimp

        def imped(pos):
            x = abs(pos) / width
            y = 1.0 if x >= 1 else (x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1))
            return min(max(dmin + (0.0 if x == 0 else y) * (dmax - dmin), 1e-4), 0.9999)

        def KB(sr):
            if sr[0] > 0:
                tc = max(sr[0], 2 * P.timestep)
                return 1 / (dmax ** 2 * tc ** 2 * sr[1] ** 2), 2 / (dmax * tc)
            return -sr[0] / dmax ** 2, -sr[1] / dmax

        n = 283
        R, Pt = quat2mat(q[10:14]), q[7:10]
        ax, iw, biw, ids = np.asarray(A["part_axis"]), np.asarray(A["dof_invweight0"]), np.asarray(A["body_invweight0"]), A["ids"]
        qs, vs = q[14:], v[13:]
        rows, D, aref = [], [], []
        K, B = KB(P.solref)
        for i in range(270):
            r = np.zeros(n); r[13 + i] = 1
            imp = imped(qs[i])
            rows.append(r); D.append(1 / ((1 - imp) / imp * iw[13 + i])); aref.append(-B * vs[i] - K * imp * qs[i])
        K2, B2 = KB((-ts[abi.TS_STIFFNESS], -ts[abi.TS_DAMPING]))
        for ia, ib in np.asarray(A["eq_pairs"]):
            r = np.zeros(n); r[13 + ia], r[13 + ib] = 1, -1
            pos, vel = qs[ia] - qs[ib], vs[ia] - vs[ib]
            imp = imped(pos)
            rows.append(r); D.append(1 / ((1 - imp) / imp * (iw[13 + ia] + iw[13 + ib]))); aref.append(-B2 * vel - K2 * imp * pos)
        r = np.zeros(n); r[13:] = 1
        imp = imped(qs.sum())
        rows.append(r); D.append(1 / ((1 - imp) / imp * float(np.asarray(A["tendon_invweight0"])[0]))); aref.append(-B * vs.sum() - K * imp * qs.sum())
        self.Je, self.De, self.ae = np.array(rows), np.array(D), np.array(aref)
        # contacts: world 3-row blocks, frame-free cone
        J, spos, _ = e.eef()
        self.cJ, self.cn, self.cD, self.ca, self.cfr, self.arm = [], [], [], [], [], []
        for k in range(len(c["dist"])):
            nrm, pos, g1, g2 = c["frame"][k][0], c["pos"][k], c["geom1"][k], c["geom2"][k]
            Jk = np.zeros((3, n))
            diagA = 0.0
            if g2 == 2:
                Jk[:, :7] += J[:3] + np.cross(J[3:].T, pos - spos).T
                diagA += biw[ids[4], 0]
            if g1 >= 4:
                i, rr = g1 - 4, pos - Pt
                Jk[:, 7:10] -= np.eye(3); Jk[:, 10:13] -= -skew(rr) @ R; Jk[:, 13 + i] -= R @ ax[i]
                diagA += biw[ids[6] + i, 0]
            imp = imped(c["dist"][k])
            vrel = Jk @ v
            self.cJ.append(Jk); self.cn.append(nrm); self.cD.append(1 / ((1 - imp) / imp * diagA))
            self.ca.append(-B * vrel - K * imp * c["dist"][k] * nrm)
            fr = max(P.table_friction if g2 == 1 or g1 == 1 else 0, P.probe_friction if g2 == 2 else 0, P.particle_friction if g1 >= 4 else 0)
            self.cfr.append(fr); self.arm.append(g2 == 2)
        self.impratio = P.impratio
        self.n = n

    # cone pieces in world form (csrc/soft.cuh cone_force / build_precond)
    def _cone(self, j, nrm, Dn, fr):
        mu, Dt = fr / np.sqrt(self.impratio), Dn * self.impratio
        jn = j @ nrm
        jt = j - jn * nrm
        N, T = jn * mu, fr * np.linalg.norm(jt)
        if N >= mu * T or (T <= 0 and N >= 0):
            return 0.0, np.zeros(3), np.zeros((3, 3)), 0
        if mu * N + T <= 0 or (T <= 0 and N < 0):
            Kq = Dt * np.eye(3) + (Dn - Dt) * np.outer(nrm, nrm)
            return 0.5 * (Dn * jn * jn + Dt * jt @ jt), -(Dn * jn * nrm + Dt * jt), Kq, 2
        Dm = Dn / (mu * mu * (1 + mu * mu))
        g = N - mu * T
        u = jt / max(np.linalg.norm(jt), 1e-300)
        gv = mu * nrm - mu * fr * u
        kk = -Dm * mu * g / T * fr * fr
        Kq = Dm * np.outer(gv, gv) + kk * (np.eye(3) - np.outer(nrm, nrm) - np.outer(u, u))
        return 0.5 * Dm * g * g, -Dm * g * gv, Kq, 1

    def evaluate(self, a, hess=False):
        """cost, gradient, zones (and the Hessian at the current zones)."""
        d = a - self.a0
        cost = 0.5 * d @ self.M @ d
        grad = self.M @ d
        je = self.Je @ a - self.ae
        cost += 0.5 * (self.De * je * je).sum()
        grad += self.Je.T @ (self.De * je)
        H = self.M + (self.Je.T * self.De) @ self.Je if hess else None
        zones = []
        for Jk, nrm, Dn, ar, fr in zip(self.cJ, self.cn, self.cD, self.ca, self.cfr):
            cst, f, Kq, z = self._cone(Jk @ a - ar, nrm, Dn, fr)
            cost += cst
            grad -= Jk.T @ f
            zones.append(z)
            if hess and z:
                H += Jk.T @ Kq @ Jk
        return cost, grad, zones, H

    def line_search(self, a, s):
        """exact minimiser of the convex, piecewise-smooth phi(alpha) = cost(a + alpha s) (safeguarded Newton on phi')"""
        def dphi(al):
            _, g, _, _ = self.evaluate(a + al * s)
            return g @ s
        d0 = dphi(0.0)
        lo, hi, al = 0.0, None, 0.0
        h = 1e-7 * max(1.0, np.abs(a).max()) / max(np.abs(s).max(), 1e-300)
        for _ in range(60):
            d1 = dphi(al)
            if abs(d1) <= 1e-10 * abs(d0):
                break
            if d1 < 0:
                lo = al
            else:
                hi = al
            d2 = (dphi(al + h) - d1) / h
            an = al - d1 / d2 if d2 > 0 else (2 * al + 1e-6)
            if hi is None:
                an = an if an > lo else 2 * al + 1e-6
            elif not (lo < an < hi):
                an = 0.5 * (lo + hi)
            al = an
        return al


def blocks_precond(p, H, coupled, exact_sliders=False):
    """P^-1 (dense) in the spirit of the kernel: dense arm / free blocks, sliders by their diagonal + cubic Chebyshev polynomial in
    the pair couplings; `coupled`: arm + free body as one 13x13 block, sliders eliminated against all 13."""
    n = p.n
    D = np.diag(H)[13:].copy()
    W = -(H[13:, 13:] - np.diag(D))
    Dt = W.min()  # the tendon's uniform off-diagonal (-Dt in W): not part of the pair couplings
    W = np.where(np.abs(W - Dt) < 1e-9 * max(1.0, abs(Dt)), 0.0, W - Dt * 0)
    W[np.abs(W) < 1e-12] = 0
    W = np.maximum(W, 0)
    N = W / D[:, None]
    Q = np.linalg.inv(H[13:, 13:]) if exact_sliders else (np.eye(270) + 1.6 * (N + N @ N)) / D[None, :]
    Pinv = np.zeros((n, n))
    dense = list(range(13)) if coupled else None
    if coupled:
        Bm = H[13:, :13]
        DB = Bm / D[:, None]
        Si = np.linalg.inv(H[:13, :13] - Bm.T @ DB)
        Pinv[:13, :13], Pinv[:13, 13:], Pinv[13:, :13], Pinv[13:, 13:] = Si, -Si @ DB.T, -DB @ Si, Q + DB @ Si @ DB.T
    else:
        Bm = H[13:, 7:13]
        DB = Bm / D[:, None]
        Si = np.linalg.inv(H[7:13, 7:13] - Bm.T @ DB)
        Pinv[:7, :7] = np.linalg.inv(H[:7, :7])
        Pinv[7:13, 7:13], Pinv[7:13, 13:], Pinv[13:, 7:13], Pinv[13:, 13:] = Si, -Si @ DB.T, -DB @ Si, Q + DB @ Si @ DB.T
    return Pinv


def pcg(p, kind, tol=1e-5, maxit=60):
    a = p.warm.copy()
    _, g, zones, H = p.evaluate(a, hess=True)
    rhsn = np.linalg.norm(p.M @ p.a0 + p.Je.T @ (p.De * p.ae))

    def make(H):
        if kind == "newton":
            return np.linalg.inv(H)
        return blocks_precond(p, H, coupled=(kind == "coupled"))

    Pinv = make(H)
    s = np.zeros(p.n)
    pg_old, gpg = np.zeros(p.n), 1.0
    rebuilds = 0
    for it in range(maxit + 1):
        hxn = np.linalg.norm(g)  # scale of the cancelling terms is dominated by rhs here
        if np.linalg.norm(g) <= tol * (1 + rhsn + hxn):
            return it, rebuilds
        pg = Pinv @ g
        beta = 0.0 if it == 0 else max(0.0, (g @ pg - g @ pg_old) / max(gpg, 1e-300))
        gpg, pg_old = g @ pg, pg
        s = -pg + beta * s
        a = a + p.line_search(a, s) * s
        _, g, z2, H = p.evaluate(a, hess=True)
        if z2 != zones:
            zones, Pinv = z2, make(H)
            rebuilds += 1
    return maxit, rebuilds


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=6)
    args = ap.parse_args()
    pk = packed_model(True)
    rng = np.random.default_rng(2)
    out = {k: [] for k in ("device", "coupled", "newton")}
    for sidx in range(args.states):
        cfg = abi.make_config(1, CC_TRACK, control_freq=500, horizon=1000, seed=7 + sidx, torso_solref_randomization=True,
                              initial_probe_pos_randomization=True)
        e = O.OracleEnv(pk, cfg, sidx)
        e.reset()
        for _ in range(20 + 7 * sidx):
            e.step(rng.uniform(0, 1, 6))
        p = Problem(e, pk)
        _, g, _, _ = p.evaluate(p.a_star)
        row = [f"state {sidx}: ncon {len(p.cJ):3d}  |grad(a*)| {np.linalg.norm(g):.1e}"]
        for kind in out:
            it, rb = pcg(p, kind)
            out[kind].append(it)
            row.append(f"{kind} {it} it / {rb} rebuilds")
        print("   ".join(row), flush=True)
    print("mean iterations:", {k: float(np.mean(v)) for k, v in out.items()})
