// arm_warp.cuh — K1/K2 with one LANE per link: 8 lanes (a "group") cooperate on one env, four envs per warp.
//
// The thread-per-env arm kernel (arm.cuh) walks the 7-link chain serially: ~5 500 straight-line instructions per env, one warp per
// SM, bound by instruction fetch and dependent-issue latency (25 us for 4096 envs, 5 % of the step).  Here every recursion over the
// chain is a SCAN over the lanes of a group, exchanged with width-8 shuffles:
//   forward kinematics        inclusive prefix PRODUCT of the local transforms (R, t), 3 Hillis-Steele steps; lane NJ carries the
//                             tool transform, so it ends up holding the grip-site pose
//   link velocities / vp acc  three prefix SUMS (omega, alpha, origin acceleration)
//   RNE bias, CRBA composites one SUFFIX sum of 16 values: link force, moment, mass, first moment and inertia, all taken about
//                             the grip site (a common reference point makes them additive; the site keeps the wrist composites'
//                             parallel-axis terms small, so the small wrist inertias do not cancel)
//   M                         lane j broadcasts the spatial force of a unit acceleration of joint j on its composite; lane i
//                             projects it on its own axis: row i of the upper triangle
//   7x7 / 6x6 Cholesky        one matrix row per lane, columns exchanged by shuffles (chol7_group)
//   M^-1 J^T                  lane r solves for task row r (six right-hand sides at once)
// Same record layout as the thread version (common.cuh AB_*), assembled in SHARED memory and written out coalesced.
// The reset kernel keeps the thread-per-env routine (arm.cuh arm_forward, ctrl = 0): it runs off the critical path.
#pragma once
#include "arm.cuh"

#define GW 8                      // lanes per env
#define ARM_SCR 160               // floats of shared scratch per env (layout below)
enum { SC_L = 0,                  // 49 Cholesky factor of M (diagonal inverted)
       SC_LF = 49,                // 49 J M^-1 J^T padded to 7x7, then its factor
       SC_QD = 98,                // 8  joint velocities
       SC_POSE = 106,             // 8  null-space posture torque before M
       SC_MP = 114,               // 8  M pose
       SC_FD = 122,               // 8  desired task wrench (PD)
       SC_W = 130,                // 8  Lambda Fd
       SC_T6 = 138,               // 8  (M^-1 J^T)^T M pose
       SC_DX = 146 };             // 8  qfrc_smooth - previous qfrc_smooth

#define GFULL 0xffffffffu
__device__ __forceinline__ float gsh(float v, int src) { return __shfl_sync(GFULL, v, src, GW); }
__device__ __forceinline__ v3 gsh3(v3 a, int src) { return mk(gsh(a.x, src), gsh(a.y, src), gsh(a.z, src)); }
// value of the lane d below (zero for the first d lanes of the group)
__device__ __forceinline__ v3 gup3(v3 a, int d, int j) {
  v3 t = mk(__shfl_up_sync(GFULL, a.x, d, GW), __shfl_up_sync(GFULL, a.y, d, GW), __shfl_up_sync(GFULL, a.z, d, GW));
  return j >= d ? t : mk(0, 0, 0);
}
// inclusive prefix sum over the lanes of a group
__device__ __forceinline__ v3 gscan_up3(v3 a, int j) {
#pragma unroll
  for (int d = 1; d < GW; d <<= 1) a = a + gup3(a, d, j);
  return a;
}
// inclusive suffix sum (lane j: sum over lanes >= j)
__device__ __forceinline__ float gscan_down(float a, int j) {
#pragma unroll
  for (int d = 1; d < GW; d <<= 1) {
    float t = __shfl_down_sync(GFULL, a, d, GW);
    if (j + d < GW) a += t;
  }
  return a;
}

// In-place Cholesky of a 7x7 (row-major, pitch 7, shared memory, both triangles valid) by the lanes of one group: row j in lane j.
// Factor in the lower triangle, diagonal INVERTED (chol7_solve multiplies).  The caller syncs the warp before and after.
__device__ __forceinline__ void chol7_group(float* A, int j) {
  const bool act = j < 7;
  float row[7];
#pragma unroll
  for (int k = 0; k < 7; k++) row[k] = act ? A[j * 7 + k] : 0.f;
#pragma unroll
  for (int c = 0; c < 7; c++) {
    float d = gsh(row[c], c);
    if (!(d > 0.f)) d = 1e-20f;
    const float inv = rsqrtf(d);
    const float l = row[c] * inv;
#pragma unroll
    for (int k = 0; k < 7; k++)
      if (k > c) row[k] -= l * gsh(l, k);
    row[c] = j == c ? inv : l;
  }
#pragma unroll
  for (int k = 0; k < 7; k++)
    if (act && k <= j) A[j * 7 + k] = row[k];
}

// One env step of the arm (controller runs), by the 8 lanes of a group; j = lane within the group.
//   q, qd      joint position / velocity of joint j (lanes j < 7; anything elsewhere)
//   act        the env's action row (global)
//   ts         the env's task record (global or shared); written only by lane 0, only if `live`
//   ab         the env's arm record in SHARED memory; on entry ab[AB_QS..+7] holds the previous physics step's qfrc_smooth
//   sc         ARM_SCR floats of shared scratch
//   live       false: a frozen / padding env -- everything is computed (the shuffles need every lane), nothing outside ab / sc is written
// Every lane of the WARP must call this (full-mask shuffles); the caller syncs the warp before reading ab.
template <int NJ>
__device__ __forceinline__ void arm_forward_group(bool policy_step, float q, float qd, const float* __restrict__ act, float* ts, float* ab,
                                                  float* sc, int j, bool live) {
  const bool on = j < NJ;
  const int jl = j < 7 ? j : 6; // table row (clamped: lane 7 reads a valid row and ignores it)
  if (!on) { q = 0.f; qd = 0.f; }

  // ---------------- forward kinematics: prefix product of (R_loc, t_loc); lane NJ = tool transform -> grip-site pose
  float R[9];
  v3 t;
  if (on) {
    float s, c;
    sincosf(q, &s, &c);
    const float* L = dm.link_R[jl];
#pragma unroll
    for (int r = 0; r < 3; r++) { // link_R Rz(q)
      R[3 * r] = c * L[3 * r] + s * L[3 * r + 1];
      R[3 * r + 1] = c * L[3 * r + 1] - s * L[3 * r];
      R[3 * r + 2] = L[3 * r + 2];
    }
    t = ld3(dm.link_pos[jl]);
  } else if (j == NJ) {
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = dm.tool[3 + i];
    t = ld3(dm.tool);
  } else {
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.f : 0.f;
    t = mk(0, 0, 0);
  }
#pragma unroll
  for (int d = 1; d < GW; d <<= 1) {
    float Rp[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Rp[i] = __shfl_up_sync(GFULL, R[i], d, GW);
    v3 tp = mk(__shfl_up_sync(GFULL, t.x, d, GW), __shfl_up_sync(GFULL, t.y, d, GW), __shfl_up_sync(GFULL, t.z, d, GW));
    if (j >= d) {
      t = tp + mv(Rp, t);
      mm3(Rp, R, R);
    }
  }
  const v3 p = t;                                              // link origin == joint anchor (lane NJ: the site)
  const v3 z = on ? mk(R[2], R[5], R[8]) : mk(0, 0, 0);        // joint axis (world)
  const v3 site = gsh3(p, NJ);
  constexpr int LL = NJ - 1;                                   // the link the hand and the probe are welded to
  const v3 hand = gsh3(p + mv(R, ld3(dm.tool + 12)), LL);

  // ---------------- velocities and velocity-product accelerations (gravity folded in: a_base = -g): three prefix sums
  const v3 zq = qd * z;
  const v3 w = gscan_up3(zq, j), wp = gup3(w, 1, j);
  const v3 al = gscan_up3(cross(wp, zq), j), alp = gup3(al, 1, j);
  const v3 r = p - gup3(p, 1, j);
  const v3 ac = gscan_up3(cross(alp, r) + cross(wp, cross(wp, r)), j) - ld3(dm.g);

  // ---------------- per-link force / moment / inertia about the site, then ONE suffix sum of 16 values
  const float m = on ? dm.link_mass[jl] : 0.f;
  const v3 cl = mv(R, ld3(dm.link_com[jl]));
  float Iw[9];
  world_inertia(R, dm.link_I[jl], Iw);
  v3 Fs, Ns, hs;
  float cm, I6[6]; // composite inertia about the site: xx, yy, zz, xy, xz, yz
  {
    const v3 acom = ac + cross(al, cl) + cross(w, cross(w, cl));
    const v3 f = m * acom;
    const v3 nn = mv(Iw, al) + cross(w, mv(Iw, w));
    const v3 d = p + cl - site;
    const v3 n0 = on ? nn + cross(d, f) : mk(0, 0, 0);
    const float dd = dot(d, d);
    float v[16] = {f.x, f.y, f.z, n0.x, n0.y, n0.z, m, m * d.x, m * d.y, m * d.z,
                   Iw[0] + m * (dd - d.x * d.x), Iw[4] + m * (dd - d.y * d.y), Iw[8] + m * (dd - d.z * d.z),
                   Iw[1] - m * d.x * d.y, Iw[2] - m * d.x * d.z, Iw[5] - m * d.y * d.z};
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = gscan_down(on ? v[k] : 0.f, j);
    Fs = mk(v[0], v[1], v[2]); Ns = mk(v[3], v[4], v[5]); cm = v[6]; hs = mk(v[7], v[8], v[9]);
#pragma unroll
    for (int k = 0; k < 6; k++) I6[k] = v[10 + k];
  }
  const v3 a = p - site;
  const float bias = dot(z, Ns - cross(a, Fs)); // moment about the joint anchor, projected on the axis
  // ---------------- M: unit acceleration of joint c moves composite c; lane i projects its spatial force on axis i
  {
    const v3 fj = cross(z, hs - cm * a);
    const v3 nj = symv(I6, z) - cross(hs, cross(z, a));
#pragma unroll
    for (int c = 0; c < 7; c++) {
      const v3 fb = gsh3(fj, c), nb = gsh3(nj, c);
      float v = dot(z, nb - cross(a, fb));
      if (!(on && c < NJ)) v = j == c ? 1.f : 0.f; // identity in the slots of absent joints
      if (j <= c) { ab[AB_M + j * 7 + c] = v; ab[AB_M + c * 7 + j] = v; }
    }
  }
  // ---------------- Jacobians (column j), site pose, probe capsule ends
  const v3 jp = cross(z, site - p);
  if (j < 7) {
    const v3 jh = cross(z, hand - p);
    ab[AB_JSITE + 0 * 7 + j] = jp.x; ab[AB_JSITE + 1 * 7 + j] = jp.y; ab[AB_JSITE + 2 * 7 + j] = jp.z;
    ab[AB_JSITE + 3 * 7 + j] = z.x; ab[AB_JSITE + 4 * 7 + j] = z.y; ab[AB_JSITE + 5 * 7 + j] = z.z;
    ab[AB_JHAND + 0 * 7 + j] = jh.x; ab[AB_JHAND + 1 * 7 + j] = jh.y; ab[AB_JHAND + 2 * 7 + j] = jh.z;
    sc[SC_QD + j] = qd;
    sc[SC_POSE + j] = on ? 10.f * (ts[USIM_TS_INIT_JOINT + j] - q) - 6.3245553203f * qd : 0.f;
  }
  if (j == NJ) {
    st3(ab + AB_EEFPOS, p);
#pragma unroll
    for (int i = 0; i < 9; i++) ab[AB_EEFR + i] = R[i];
    float qx[4];
    mat2quat_xyzw(R, qx);
    ab[AB_QUAT] = qx[0]; ab[AB_QUAT + 1] = qx[1]; ab[AB_QUAT + 2] = qx[2]; ab[AB_QUAT + 3] = qx[3];
  }
  // ---------------- F/T sensor pieces for the probe body (welded to link LL): lane LL owns the probe, every lane its column
  {
    float Ip[9];
    world_inertia(R, dm.tool + 26, Ip);
    const float mp = dm.tool[22];
    const v3 clp = mv(R, ld3(dm.tool + 23));
    const v3 rc_l = p + clp - site; // probe COM relative to the site
    if (j == LL) {
      const v3 acom = ac + cross(al, clp) + cross(w, cross(w, clp));
      const v3 t0 = mv(Ip, al) + cross(w, mv(Ip, w)) + cross(rc_l, mp * acom);
      st3(ab + AB_TAU0, t0);
      st3(ab + AB_PTIP, p + mv(R, ld3(dm.tool + 15)));
      st3(ab + AB_PBACK, p + mv(R, ld3(dm.tool + 18)));
    }
    const v3 rc = gsh3(rc_l, LL);
    const float Is[6] = {gsh(Ip[0], LL), gsh(Ip[4], LL), gsh(Ip[8], LL), gsh(Ip[1], LL), gsh(Ip[2], LL), gsh(Ip[5], LL)};
    if (j < 7) {
      const v3 tt = symv(Is, z) + cross(rc, mp * (jp + cross(z, rc))); // zero for an absent joint (z = 0, jp = 0)
      ab[AB_JFT + 0 * 7 + j] = tt.x; ab[AB_JFT + 1 * 7 + j] = tt.y; ab[AB_JFT + 2 * 7 + j] = tt.z;
    }
  }
  __syncwarp();

  // ---------------- OSC_POSE torques [SURVEY App. C.2/C.3]
  float av[7];
  for (int i = 0; i < dm.adim; i++) av[i] = act[i];
  if (j == 0 && policy_step && live) osc_set_goal(av, ld3(ab + AB_EEFPOS), ab + AB_EEFR, ts);
  // factor of M (row per lane) while lane 0 writes the goal
  if (j < 7) {
#pragma unroll
    for (int k = 0; k < 7; k++) sc[SC_L + j * 7 + k] = ab[AB_M + j * 7 + k];
  }
  __syncwarp();
  chol7_group(sc + SC_L, j);
  const int rr = j < 6 ? j : 5; // task row of this lane (lanes 6, 7 shadow row 5 and write nothing)
  float Fd, Mp = 0.f;
  {
    float kp, kd;
    if (dm.mode == USIM_MODE_FIXED) { kp = dm.kp[rr]; kd = 2.f * sqrtf(kp) * dm.dr[rr]; }
    else { kp = scale1(av[rr], dm.kp_in_min, dm.kp_in_max, dm.kp_lim[0], dm.kp_lim[1]); kd = 2.f * sqrtf(kp); }
    float vel = 0.f;
#pragma unroll
    for (int k = 0; k < 7; k++) vel += ab[AB_JSITE + rr * 7 + k] * sc[SC_QD + k];
    if (dm.mode == USIM_MODE_WRENCH) {
      Fd = fminf(fmaxf(av[rr], -10.f), 10.f);
    } else {
      const v3 eo = ori_error(ts + USIM_TS_GOAL_ORI, ab + AB_EEFR);
      const float err = rr < 3 ? ts[USIM_TS_GOAL_POS + rr] - ab[AB_EEFPOS + rr] : (rr == 3 ? eo.x : rr == 4 ? eo.y : eo.z);
      Fd = kp * err - kd * vel;
    }
    if (j < 6) sc[SC_FD + j] = Fd;
    if (j < 7) { // M pose (row j)
#pragma unroll
      for (int k = 0; k < 7; k++) Mp += ab[AB_M + j * 7 + k] * sc[SC_POSE + k];
      sc[SC_MP + j] = Mp;
    }
  }
  __syncwarp();
  // M^-1 J^T: lane r solves for task row r; then column r of J M^-1 J^T and entry r of (M^-1 J^T)^T M pose
  {
    float col[7];
#pragma unroll
    for (int k = 0; k < 7; k++) col[k] = ab[AB_JSITE + rr * 7 + k];
    chol7_solve<7>(sc + SC_L, col);
    float t6 = 0.f;
#pragma unroll
    for (int k = 0; k < 7; k++) t6 += col[k] * sc[SC_MP + k];
    if (j < 6) {
      sc[SC_T6 + j] = t6;
#pragma unroll
      for (int r2 = 0; r2 < 6; r2++) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 7; k++) s += ab[AB_JSITE + r2 * 7 + k] * col[k];
        sc[SC_LF + r2 * 7 + j] = s;
      }
    } else if (j == 6) {
#pragma unroll
      for (int k = 0; k < 7; k++) { sc[SC_LF + 6 * 7 + k] = k == 6 ? 1.f : 0.f; sc[SC_LF + k * 7 + 6] = k == 6 ? 1.f : 0.f; }
    }
  }
  __syncwarp();
  // Lambda Fd with the position / orientation blocks uncoupled: two 3x3 solves, lanes 0 and 1
  if (dm.mode == USIM_MODE_WRENCH) {
    if (j < 6) sc[SC_W + j] = Fd;
  } else if (dm.uncouple && j < 2) {
    const int o = 3 * j;
    float B[9], x3[3];
#pragma unroll
    for (int r2 = 0; r2 < 3; r2++) {
      x3[r2] = sc[SC_FD + o + r2];
#pragma unroll
      for (int s2 = 0; s2 < 3; s2++) B[r2 * 3 + s2] = sc[SC_LF + (o + r2) * 7 + o + s2];
    }
    chol<3>(B);
    chol_solve<3>(B, x3);
    sc[SC_W + o] = x3[0]; sc[SC_W + o + 1] = x3[1]; sc[SC_W + o + 2] = x3[2];
  }
  __syncwarp();
  chol7_group(sc + SC_LF, j);
  __syncwarp();
  // every lane: Lambda_full t6 (null-space projection) and, coupled, Lambda_full Fd; then its own torque
  float tau, qs;
  {
    float t6[6], W[6];
#pragma unroll
    for (int k = 0; k < 6; k++) t6[k] = sc[SC_T6 + k];
    chol7_solve<6>(sc + SC_LF, t6);
    if (dm.mode != USIM_MODE_WRENCH && !dm.uncouple) {
#pragma unroll
      for (int k = 0; k < 6; k++) W[k] = sc[SC_FD + k];
      chol7_solve<6>(sc + SC_LF, W);
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) W[k] = sc[SC_W + k];
    }
    float s = bias + Mp;
#pragma unroll
    for (int r2 = 0; r2 < 6; r2++) s += ab[AB_JSITE + r2 * 7 + jl] * (W[r2] - t6[r2]);
    tau = fminf(fmaxf(s, -dm.ctrl[jl]), dm.ctrl[jl]);
    if (!on) tau = 0.f;
    qs = tau - bias - dm.arm_damp * qd;
  }
  // warm-start shift for the solve kernel: M^-1 (qfrc_smooth - qfrc_smooth of the previous physics step); the record still holds the
  // previous value here.  On the first step of an episode it is stale (the warm start is zero then anyway).
  const bool shift = !(policy_step && ts[USIM_TS_TIMESTEP] == 0.f);
  if (j < 7) sc[SC_DX + j] = shift ? qs - ab[AB_QS + j] : 0.f;
  __syncwarp();
  if (j < 7) { ab[AB_QS + j] = qs; ab[AB_TAU + j] = tau; }
  {
    float dx[7];
#pragma unroll
    for (int k = 0; k < 7; k++) dx[k] = sc[SC_DX + k];
    chol7_solve<7>(sc + SC_L, dx);
    if (j == 0) {
#pragma unroll
      for (int k = 0; k < 7; k++) ab[AB_DX + k] = dx[k];
      ab[AB_DX + 7] = 0.f; // (pad of the record)
    }
  }
}

// Four envs per warp.  Also clears `done` (frozen envs report done = 0; the solve kernel sets it for the envs it steps) and flattens
// the iteration-count bins the previous solve launch filled into the launch order of the next one (highest bin first; soft.cuh NBIN).
#ifndef ARMW_BLOCK
#define ARMW_BLOCK 32
#endif
template <int NJ>
__global__ void __launch_bounds__(ARMW_BLOCK) arm_kernel_w(int n, const float* __restrict__ qpos, const float* __restrict__ qvel,
                                                          const float* __restrict__ act, float* __restrict__ task, float* __restrict__ armbuf,
                                                          uint8_t* __restrict__ done, int policy_step, int nbin,
                                                          const int* __restrict__ bin_cnt_prev, const int* __restrict__ bin_items_prev,
                                                          int* __restrict__ bin_cnt_next, int* __restrict__ order, int* __restrict__ queue) {
  if (queue && blockIdx.x == 0 && threadIdx.x == 0) *queue = 0; // work queue of the solve launch that follows
  __shared__ __align__(16) float sm[ARMW_BLOCK / GW][ARMBUF + ARM_SCR];
  const int g = threadIdx.x / GW, j = threadIdx.x % GW;
  const int env = blockIdx.x * (ARMW_BLOCK / GW) + g;
  const bool inr = env < n;
  const int e = inr ? env : n - 1;
  if (inr && j == 0) {
    if (order) {
      int off = 0, b = nbin - 1;
      for (; b > 0; b--) {
        const int c = bin_cnt_prev[b];
        if (env < off + c) break;
        off += c;
      }
      order[env] = bin_items_prev[(size_t)b * n + min(env - off, n - 1)];
      for (int k = env; k < nbin; k += n) bin_cnt_next[k] = 0; // (n may be smaller than the number of bins)
    }
    if (done) done[env] = 0;
  }
  float* ts = task + (size_t)e * USIM_TASK_DIM;
  float* ab = sm[g];
  float* sc = ab + ARMBUF;
  const bool live = inr && ts[USIM_TS_DONE] == 0.f; // terminated env: frozen until reset, its record stays
  float q = 0.f, qd = 0.f;
  if (j < 7) {
    q = qpos[(size_t)e * QPAD + j]; qd = qvel[(size_t)e * QPAD + j];
    ab[AB_QS + j] = armbuf[(size_t)e * ARMBUF + AB_QS + j];
  }
  arm_forward_group<NJ>(policy_step != 0, q, qd, act + (size_t)e * dm.adim, ts, ab, sc, j, live);
  __syncwarp();
  if (live) {
    float* dst = armbuf + (size_t)env * ARMBUF;
    for (int i = j; i < ARMBUF; i += GW) dst[i] = ab[i];
  }
}
