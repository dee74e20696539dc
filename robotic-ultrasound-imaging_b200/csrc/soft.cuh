// soft.cuh — K3..K9 fused: one WARP per env, the whole env state staged in shared memory.
//
//   K3 torso kinematics + inertia + bias   (free body + 270 radial sliders, arrow-structured M)
//   K4 collision: table<->particle, probe<->particle, table<->probe, deterministic MuJoCo ordering
//   K5 constraint parameters: solref/solimp -> aref, D; equality rows folded into a 270-node stencil
//   K6 primal nonlinear CG (Polak-Ribiere, exact Newton line search, elliptic cones) with an
//      arrow preconditioner (6x6 Schur complement over the slider diagonal) + 7x7 arm block
//   K7 semi-implicit Euler, quaternion integration      K8 cfrc_ext / F-T torque of the probe
//   K9 reward, running statistics, termination, observation (ultrasound.py:230-269,363-401,512-550,635-670)
//
// Unknown vector layout (nv = 283): [0..6] arm, [7..9] free translation (world), [10..12] free
// rotation (body frame), [13+i] slider i.  No constraint Jacobian is ever materialised: contact
// rows are evaluated as rigid-body point velocities / wrenches, equality rows as a graph stencil.
#pragma once
#include "common.cuh"

// One env per CTA (no CTA waits for a slower env; measured 1/2/4/8 envs per CTA: 2.17/2.12/1.96/1.82 M steps/s).
// WPE warps cooperate on that env over the same shared-memory image: more resident warps per SM without more shared memory.
#ifndef WPE
#define WPE 1
#endif
#define NT (32 * WPE)
#define RED_MAX 24
#define NPAIR_MAX 544
// unroll factor of the hot per-slider / per-contact loops of the CG iteration: trades ILP against the size of the loop body
// (the body must stay inside the instruction cache; `no_instruction` was the top stall of the first profile)
// measured 4096 envs: compiler default 3.66 M steps/s, forced unroll 1 / 2 / 4: 3.20 / 2.74 / 2.64 M -> leave it to the compiler
#define USIM_STR2(x) #x
#define USIM_STR(x) USIM_STR2(x)
#ifdef HOT_UNROLL
#define PRAGMA_HOT _Pragma(USIM_STR(unroll HOT_UNROLL))
#else
#define PRAGMA_HOT
#endif

struct __align__(16) WS {
  float qs[NPART_MAX];                                 // slider position (slider velocity lives in hs[13..] until the CG loop starts)
  // Hx holds (M+E)x - rhs; during set-up `grad` accumulates rhs (= qfrc_smooth + J_eq^T D aref)
  float x[QPAD], Hx[QPAD], grad[QPAD], pg[QPAD], s[QPAD], hs[QPAD];
  float dg[NPART_MAX], dg0[NPART_MAX], kx[NPART_MAX], ky[NPART_MAX], kz[NPART_MAX], df[NPART_MAX]; // dg0: diagonal without contacts
  float Dp[NPAIR_MAX];                                 // D of each "smooth" pair
  float cpos[3][DEV_MAXC], cn[3][DEV_MAXC], cjar[3][DEV_MAXC], cjv[3][DEV_MAXC]; // cjv holds aref until the first J*x
  float cD[DEV_MAXC], cdist[DEV_MAXC];
  short cpart[DEV_MAXC];
  unsigned char ctype[DEV_MAXC], czone[DEV_MAXC];
  float ab[ARMBUF];
  float ts[USIM_TASK_DIM];
  float R[9], p[3], vf[6], qdarm[7];
  float Mff[36], Sf[36], Pa[49];
  float mc[3], dv[12], red[24];
  float red2[2][WPE][RED_MAX]; // double-buffered cross-warp reduction scratch
  int cnt2[2][WPE];
  float lsign[7], lD[7], laref[7];
  float Dt, areft, mtot;
  int ncon;
};

__device__ __forceinline__ void env_sync() {
  if (WPE == 1) __syncwarp(); else __syncthreads();
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void cone_force(float j0, float j1, float j2, float Dn, float Dt, float mu, float fr, float& f0, float& f1,
                                           float& f2, int& zone) {
  float N = j0 * mu, U1 = j1 * fr, U2 = j2 * fr, T = sqrtf(U1 * U1 + U2 * U2);
  if (N >= mu * T || (T <= 0.f && N >= 0.f)) { zone = 0; f0 = f1 = f2 = 0.f; }
  else if (mu * N + T <= 0.f || (T <= 0.f && N < 0.f)) { zone = 2; f0 = -Dn * j0; f1 = -Dt * j1; f2 = -Dt * j2; }
  else {
    zone = 1;
    float Dm = Dn / (mu * mu * (1.f + mu * mu)), NmT = N - mu * T;
    f0 = -Dm * NmT * mu;
    f1 = -f0 / T * U1 * fr;
    f2 = -f0 / T * U2 * fr;
  }
}
// first / second directional derivative of the cone cost at jar along jv
__device__ __forceinline__ void cone_ls(float j0, float j1, float j2, float v0, float v1, float v2, float Dn, float Dt, float mu, float fr,
                                        float& d1, float& d2) {
  float N = j0 * mu, U1 = j1 * fr, U2 = j2 * fr, T = sqrtf(U1 * U1 + U2 * U2);
  if (N >= mu * T || (T <= 0.f && N >= 0.f)) { d1 = 0.f; d2 = 0.f; }
  else if (mu * N + T <= 0.f || (T <= 0.f && N < 0.f)) {
    d1 = Dn * j0 * v0 + Dt * (j1 * v1 + j2 * v2);
    d2 = Dn * v0 * v0 + Dt * (v1 * v1 + v2 * v2);
  } else {
    float Dm = Dn / (mu * mu * (1.f + mu * mu)), g = N - mu * T;
    float Np = mu * v0, U1p = fr * v1, U2p = fr * v2;
    float Tp = (U1 * U1p + U2 * U2p) / T;
    float Tpp = (U1p * U1p + U2p * U2p - Tp * Tp) / T;
    float gp = Np - mu * Tp;
    d1 = Dm * g * gp;
    d2 = Dm * (gp * gp - g * mu * Tpp);
  }
}

__device__ __forceinline__ void contact_params(int type, float& fr, float& mu) {
  fr = type == 0 ? dm.fr_table_part : (type == 1 ? dm.fr_probe_part : dm.fr_table_probe);
  mu = fr * rsqrtf(dm.impratio);
}

// closest points between segments (p1,q1) and (p2,q2) — same branch structure as the oracle
__device__ __forceinline__ void seg_seg(v3 p1, v3 q1, v3 p2, v3 q2, v3& c1, v3& c2) {
  v3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2), den = a * e - b * b, s, t;
  if (den > 1e-14f * a * e) s = fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f); else s = 0.f;
  t = (b * s + f) / e;
  if (t < 0.f) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
  else if (t > 1.f) { t = 1.f; s = fminf(fmaxf((b - c) / a, 0.f), 1.f); }
  c1 = p1 + s * d1;
  c2 = p2 + t * d2;
}

// mode: 0 = env step, 1 = reset forward (no integration; initialises the running statistics)
__global__ void __launch_bounds__(NT, 8) solve_kernel(
    int n, int mode, const uint8_t* __restrict__ mask, float* __restrict__ qpos, float* __restrict__ qvel, float* __restrict__ warm,
    float* __restrict__ task, const float* __restrict__ armbuf, PartTables pt, const int* __restrict__ eq_pairs,
    const short* __restrict__ nbr_pair, float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ done,
    float* __restrict__ diag, int* __restrict__ ncon_out, int* __restrict__ geom1_out, int* __restrict__ geom2_out,
    float* __restrict__ dist_out, int* __restrict__ diverged) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int env = blockIdx.x;
  if (env >= n) return;
  if (mask && !mask[env]) return;
  WS& w = *reinterpret_cast<WS*>(smem_raw);
  int rphase = 0;
  // block-wide sums of K values: warp shuffles, then one barrier over double-buffered scratch (identical result in every thread)
  auto bsumk = [&](auto& v) {
    constexpr int K = sizeof(v) / sizeof(float);
    static_assert(K <= RED_MAX, "reduction batch too large");
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = wsum(v[k]);
    if (WPE > 1) {
      float(*buf)[RED_MAX] = w.red2[rphase];
      rphase ^= 1;
      if (lane == 0) { // lane 0 of EVERY warp publishes its warp's partial sums
#pragma unroll
        for (int k = 0; k < K; k++) buf[wrp][k] = v[k];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < K; k++) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < WPE; q++) t += buf[q][k];
        v[k] = t;
      }
    }
  };
  auto bsum = [&](float x) -> float {
    float v1[1] = {x};
    bsumk(v1);
    return v1[0];
  };
  float* ts_g = task + (size_t)env * USIM_TASK_DIM;
  if (mode == 0 && ts_g[USIM_TS_DONE] != 0.f) return;
  const int np = dm.soft ? dm.npart : 0;
  const int nv = 7 + (dm.soft ? 6 + np : 0);
  const float h = dm.h;
  float* qp_g = qpos + (size_t)env * QPAD;
  float* qv_g = qvel + (size_t)env * QPAD;
  float* wm_g = warm + (size_t)env * QPAD;

  // ------------------------------------------------------------------ load
  for (int i = tid; i < ARMBUF; i += NT) w.ab[i] = armbuf[(size_t)env * ARMBUF + i];
  for (int i = tid; i < USIM_TASK_DIM; i += NT) w.ts[i] = ts_g[i];
  for (int i = tid; i < QPAD; i += NT) { w.x[i] = i < nv ? wm_g[i] : 0.f; w.grad[i] = 0.f; }
  for (int i = tid; i < np; i += NT) {
    w.qs[i] = qp_g[14 + i]; w.hs[13 + i] = qv_g[13 + i];
  }
  if (tid < 7) w.qdarm[tid] = qv_g[tid];
  float quat[4] = {1, 0, 0, 0};
  if (dm.soft) {
    if (tid < 6) w.vf[tid] = qv_g[7 + tid];
    if (tid < 3) w.p[tid] = qp_g[7 + tid];
    quat[0] = qp_g[10]; quat[1] = qp_g[11]; quat[2] = qp_g[12]; quat[3] = qp_g[13];
    float nq = rsqrtf(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
    quat[0] *= nq; quat[1] *= nq; quat[2] *= nq; quat[3] *= nq;
    if (tid == 0) quat2mat(quat, w.R);
  }
  env_sync();
  const float off = dm.cap_r + dm.cap_hl;
  const v3 site = ld3(w.ab + AB_EEFPOS), ptip = ld3(w.ab + AB_PTIP), pback = ld3(w.ab + AB_PBACK);
  float R[9];
  v3 P = mk(0, 0, 0), vlin = mk(0, 0, 0), wl = mk(0, 0, 0), ww = mk(0, 0, 0);
  if (dm.soft) {
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = w.R[i];
    P = ld3(w.p); vlin = ld3(w.vf); wl = ld3(w.vf + 3); ww = mv(R, wl);
  }

  // ------------------------------------------------------------------ K3: torso inertia + bias, equality parameters
  const float ksm = -w.ts[USIM_TS_STIFFNESS], bsm = -w.ts[USIM_TS_DAMPING];
  if (dm.soft) {
    v3 gl = mtv(R, ld3(dm.g));
    float a_mc[3] = {0, 0, 0}, a_I[6] = {0, 0, 0, 0, 0, 0}, a_F[3] = {0, 0, 0}, a_T[3] = {0, 0, 0}, a_q = 0.f, a_v = 0.f;
    const float m = dm.part_mass;
    for (int i = tid; i < np; i += NT) {
      v3 ah = ld3(pt.axis + 3 * i), r0 = ld3(pt.pos + 3 * i);
      float q = w.qs[i], sd = w.hs[13 + i];
      v3 c = r0 + (q - off) * ah;
      a_mc[0] += m * c.x; a_mc[1] += m * c.y; a_mc[2] += m * c.z;
      float cc = dot(c, c);
      a_I[0] += m * (cc - c.x * c.x); a_I[1] += m * (cc - c.y * c.y); a_I[2] += m * (cc - c.z * c.z);
      a_I[3] -= m * c.x * c.y; a_I[4] -= m * c.x * c.z; a_I[5] -= m * c.y * c.z;
      v3 avp = cross(wl, cross(wl, c)) + 2.f * sd * cross(wl, ah);
      v3 F = m * (avp - gl);
      v3 T = cross(c, F);
      a_F[0] += F.x; a_F[1] += F.y; a_F[2] += F.z; a_T[0] += T.x; a_T[1] += T.y; a_T[2] += T.z;
      a_q += q; a_v += sd;
      w.grad[13 + i] = -dot(ah, F);
      // "fix" equality of this slider
      float K, B, imp;
      kbi(dm.solref[0], dm.solref[1], q, &K, &B, &imp);
      float D = 1.f / fmaxf(1e-15f, (1.f - imp) / imp * pt.iw_dof[i]);
      w.df[i] = D;
      w.dg[i] = m + D;
      w.grad[13 + i] += D * (-B * sd - K * imp * q);
    }
    {
      float r17[17] = {a_mc[0], a_mc[1], a_mc[2], a_F[0], a_F[1], a_F[2], a_T[0], a_T[1], a_T[2],
                       a_I[0], a_I[1], a_I[2], a_I[3], a_I[4], a_I[5], a_q, a_v};
      bsumk(r17);
#pragma unroll
      for (int k = 0; k < 3; k++) { a_mc[k] = r17[k]; a_F[k] = r17[3 + k]; a_T[k] = r17[6 + k]; }
#pragma unroll
      for (int k = 0; k < 6; k++) a_I[k] = r17[9 + k];
      a_q = r17[15]; a_v = r17[16];
    }
    // tendon equality: sum q = 0
    float K, B, imp;
    kbi(dm.solref[0], dm.solref[1], a_q, &K, &B, &imp);
    float Dt = 1.f / fmaxf(1e-15f, (1.f - imp) / imp * dm.tendon_iw);
    float areft = -B * a_v - K * imp * a_q;
    float mtot = np * m + dm.center_mass;
    if (tid == 0) {
      w.Dt = Dt; w.areft = areft; w.mtot = mtot;
      w.mc[0] = a_mc[0]; w.mc[1] = a_mc[1]; w.mc[2] = a_mc[2];
      // M_ff: [v (world); omega (body)]
      // rotational inertia: parallel-axis part (depends on q) + constant part (capsules about their COM, centre geom)
      float It[9] = {a_I[0] + dm.rot_I[0], a_I[3] + dm.rot_I[3], a_I[4] + dm.rot_I[4], a_I[3] + dm.rot_I[3], a_I[1] + dm.rot_I[1],
                     a_I[5] + dm.rot_I[5], a_I[4] + dm.rot_I[4], a_I[5] + dm.rot_I[5], a_I[2] + dm.rot_I[2]};
      for (int a = 0; a < 36; a++) w.Mff[a] = 0.f;
      for (int a = 0; a < 3; a++) w.Mff[a * 6 + a] = mtot;
      // M_v,omega = -R [mc]x  ;  [mc]x = [[0,-z,y],[z,0,-x],[-y,x,0]]
      float X[9] = {0, -a_mc[2], a_mc[1], a_mc[2], 0, -a_mc[0], -a_mc[1], a_mc[0], 0}, RX[9];
      mm3(R, X, RX);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) { w.Mff[a * 6 + 3 + b] = -RX[3 * a + b]; w.Mff[(3 + b) * 6 + a] = -RX[3 * a + b]; }
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) w.Mff[(3 + a) * 6 + 3 + b] = It[3 * a + b];
      // bias of the free body
      v3 sF = mk(a_F[0], a_F[1], a_F[2]), sT = mk(a_T[0], a_T[1], a_T[2]);
      v3 bv = mv(R, sF) - dm.center_mass * ld3(dm.g);
      v3 Iw = symv(dm.rot_I, wl);
      v3 bw = sT + cross(wl, Iw);
      w.grad[7] = -bv.x - dm.free_damp * vlin.x; w.grad[8] = -bv.y - dm.free_damp * vlin.y; w.grad[9] = -bv.z - dm.free_damp * vlin.z;
      w.grad[10] = -bw.x - dm.free_damp * wl.x; w.grad[11] = -bw.y - dm.free_damp * wl.y; w.grad[12] = -bw.z - dm.free_damp * wl.z;
    }
    env_sync();
    // "smooth" pair equalities (carry solrefsmooth = (-stiffness, -damping) of this episode)
    for (int pr = tid; pr < dm.npair; pr += NT) {
      int a = eq_pairs[2 * pr], b = eq_pairs[2 * pr + 1];
      float pos = w.qs[a] - w.qs[b], vel = w.hs[13 + a] - w.hs[13 + b], K2, B2, imp2;
      kbi(ksm, bsm, pos, &K2, &B2, &imp2);
      float D = 1.f / fmaxf(1e-15f, (1.f - imp2) / imp2 * (pt.iw_dof[a] + pt.iw_dof[b]));
      w.Dp[pr] = D;
      if (pr == 0) w.Dp[dm.npair] = 0.f; // the slot empty neighbour entries point at
      float ar = D * (-B2 * vel - K2 * imp2 * pos);
      atomicAdd(&w.grad[13 + a], ar); atomicAdd(&w.grad[13 + b], -ar);
      atomicAdd(&w.dg[a], D); atomicAdd(&w.dg[b], D);
    }
    env_sync();
    for (int i = tid; i < np; i += NT) { w.grad[13 + i] += w.Dt * w.areft; w.dg0[i] = w.dg[i] + w.Dt; }
  }
  if (tid < 7) {
    w.grad[lane] = w.ab[AB_QS + lane];
    // joint limits (margin 0)
    float q = qp_g[lane], lo = dm.jnt_lo[lane], hi = dm.jnt_hi[lane], dist = 0.f, sg = 0.f;
    if (q - lo < 0.f) { dist = q - lo; sg = 1.f; } else if (hi - q < 0.f) { dist = hi - q; sg = -1.f; }
    float K, B, imp;
    kbi(dm.solref[0], dm.solref[1], dist, &K, &B, &imp);
    w.lsign[lane] = sg;
    w.lD[lane] = 1.f / fmaxf(1e-15f, (1.f - imp) / imp * dm.iw_arm[lane]);
    w.laref[lane] = -B * sg * w.qdarm[tid] - K * imp * dist;
  }
  env_sync();

  // ------------------------------------------------------------------ K4: collision, MuJoCo contact order
  int ncon = 0;
  { // (table, probe): geom1 = table, geom2 = probe
    if (tid == 0) {
      v3 ends[2] = {ptip, pback};
      for (int e = 0; e < 2; e++) {
        float dist = ends[e].z - dm.probe_r - dm.table_z;
        if (dist < 0.f && fabsf(ends[e].x) <= dm.table_half && fabsf(ends[e].y) <= dm.table_half) {
          w.cpos[0][ncon] = ends[e].x; w.cpos[1][ncon] = ends[e].y; w.cpos[2][ncon] = dm.table_z + 0.5f * dist;
          w.cn[0][ncon] = 0.f; w.cn[1][ncon] = 0.f; w.cn[2][ncon] = 1.f;
          w.cdist[ncon] = dist; w.cpart[ncon] = -1; w.ctype[ncon] = 2;
          ncon++;
        }
      }
    }
    if (tid == 0) w.ncon = ncon;
    env_sync();
    ncon = w.ncon;
  }
  if (dm.soft) {
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) { // pass 0: (table, particle k); pass 1: (probe, particle k)
      for (int base = 0; base < np; base += NT) {
        int i = base + tid;
        bool h0 = false, h1 = false;
        v3 p0 = mk(0, 0, 0), p1 = mk(0, 0, 0), n0 = mk(0, 0, -1);
        float d0 = 0.f, d1 = 0.f;
        if (i < np) {
          v3 ah = ld3(pt.axis + 3 * i), r0 = ld3(pt.pos + 3 * i);
          float q = w.qs[i];
          v3 eo = P + mv(R, r0 + (q - dm.cap_r) * ah), ei = P + mv(R, r0 + (q - dm.cap_r - 2.f * dm.cap_hl) * ah);
          if (pass == 0) {
            d0 = eo.z - dm.cap_r - dm.table_z;
            h0 = d0 < 0.f && fabsf(eo.x) <= dm.table_half && fabsf(eo.y) <= dm.table_half;
            p0 = mk(eo.x, eo.y, dm.table_z + 0.5f * d0);
            d1 = ei.z - dm.cap_r - dm.table_z;
            h1 = d1 < 0.f && fabsf(ei.x) <= dm.table_half && fabsf(ei.y) <= dm.table_half;
            p1 = mk(ei.x, ei.y, dm.table_z + 0.5f * d1);
          } else {
            v3 c1, c2;
            seg_seg(eo, ei, ptip, pback, c1, c2);
            v3 d = c2 - c1;
            float len = norm(d);
            d0 = len - dm.cap_r - dm.probe_r;
            h0 = d0 < 0.f;
            n0 = len > 1e-12f ? (1.f / len) * d : mk(0, 0, 1);
            p0 = c1 + (dm.cap_r + 0.5f * d0) * n0;
          }
        }
        unsigned b0 = __ballot_sync(0xffffffffu, h0), b1 = __ballot_sync(0xffffffffu, h1);
        int before = 0, total = __popc(b0) + __popc(b1);
        if (WPE > 1) { // cross-warp exclusive prefix of the per-warp hit counts (keeps the particle order)
          int(*cb) = w.cnt2[rphase];
          rphase ^= 1;
          if (lane == 0) cb[wrp] = total;
          __syncthreads();
          total = 0;
#pragma unroll
          for (int q = 0; q < WPE; q++) { if (q < wrp) before += cb[q]; total += cb[q]; }
        }
        int slot = ncon + before + __popc(b0 & lt) + __popc(b1 & lt);
        if (h0 && slot < DEV_MAXC) {
          w.cpos[0][slot] = p0.x; w.cpos[1][slot] = p0.y; w.cpos[2][slot] = p0.z;
          w.cn[0][slot] = n0.x; w.cn[1][slot] = n0.y; w.cn[2][slot] = n0.z;
          w.cdist[slot] = d0; w.cpart[slot] = (short)i; w.ctype[slot] = (unsigned char)pass;
        }
        if (h0) slot++;
        if (h1 && slot < DEV_MAXC) {
          w.cpos[0][slot] = p1.x; w.cpos[1][slot] = p1.y; w.cpos[2][slot] = p1.z;
          w.cn[0][slot] = 0.f; w.cn[1][slot] = 0.f; w.cn[2][slot] = -1.f;
          w.cdist[slot] = d1; w.cpart[slot] = (short)i; w.ctype[slot] = 0;
        }
        ncon += total;
      }
    }
  }
  const int ncon_found = ncon;
  if (ncon > DEV_MAXC) ncon = DEV_MAXC;
  env_sync();

  // ------------------------------------------------------------------ K5: contact parameters (aref, D)
  v3 Vs, Ws; // site velocity (linear, angular)
  {
    float s6[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 7; j++) s += w.ab[AB_JSITE + r * 7 + j] * w.qdarm[j];
      s6[r] = s;
    }
    Vs = mk(s6[0], s6[1], s6[2]); Ws = mk(s6[3], s6[4], s6[5]);
  }
  for (int c = tid; c < ncon; c += NT) {
    int type = w.ctype[c], i = w.cpart[c];
    v3 pos = mk(w.cpos[0][c], w.cpos[1][c], w.cpos[2][c]), nn = mk(w.cn[0][c], w.cn[1][c], w.cn[2][c]), t1, t2;
    make_frame(nn, &t1, &t2);
    v3 rel = mk(0, 0, 0);
    float diagA = 0.f;
    if (type != 2) {
      v3 aw = mv(R, ld3(pt.axis + 3 * i));
      rel = rel - (vlin + cross(ww, pos - P) + w.hs[13 + i] * aw);
      diagA += pt.iw_body[i];
    }
    if (type != 0) { rel = rel + Vs + cross(Ws, pos - site); diagA += dm.iw_probe; }
    float K, B, imp;
    kbi(dm.solref[0], dm.solref[1], w.cdist[c], &K, &B, &imp);
    w.cD[c] = 1.f / fmaxf(1e-15f, (1.f - imp) / imp * diagA);
    w.cjv[0][c] = -B * dot(nn, rel) - K * imp * w.cdist[c];
    w.cjv[1][c] = -B * dot(t1, rel);
    w.cjv[2][c] = -B * dot(t2, rel);
  }
  env_sync();

  // ------------------------------------------------------------------ helpers (lambdas over the warp)
  // out = (M + E) in
  auto applyH = [&](const float* in, float* out) {
    float sx = 0.f, cl[3] = {0, 0, 0};
    v3 ivl = mk(0, 0, 0);
    if (dm.soft) {
      ivl = mtv(R, ld3(in + 7)); // R^T in_v
      PRAGMA_HOT
      for (int i = tid; i < np; i += NT) {
        float xi = in[13 + i];
        v3 ah = ld3(pt.axis + 3 * i);
        sx += xi;
        cl[0] += dm.part_mass * ah.x * xi; cl[1] += dm.part_mass * ah.y * xi; cl[2] += dm.part_mass * ah.z * xi;
      }
      {
        float r4[4] = {sx, cl[0], cl[1], cl[2]};
        bsumk(r4);
        sx = r4[0]; cl[0] = r4[1]; cl[1] = r4[2]; cl[2] = r4[3];
      }
      PRAGMA_HOT
      for (int i = tid; i < np; i += NT) {
        float xi = in[13 + i];
        v3 ah = ld3(pt.axis + 3 * i);
        float acc = dm.part_mass * (dot(ah, ivl) + xi) + w.df[i] * xi + w.Dt * sx;
        // 6 packed (pair << 16 | neighbour) entries, three 8-byte loads issued back to back, no data-dependent branch
        const int2* row = reinterpret_cast<const int2*>(pt.nbrpk + 6 * i);
        int2 e01 = row[0], e23 = row[1], e45 = row[2];
        int ee[6] = {e01.x, e01.y, e23.x, e23.y, e45.x, e45.y};
#pragma unroll
        for (int k = 0; k < 6; k++) acc += w.Dp[ee[k] >> 16] * (xi - in[13 + (ee[k] & 0xffff)]);
        out[13 + i] = acc;
      }
    }
    if (tid < 7) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 7; j++) s += w.ab[AB_M + lane * 7 + j] * in[j];
      out[lane] = s;
    } else if (tid < 13 && dm.soft) {
      int r = lane - 7;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 6; c++) s += w.Mff[r * 6 + c] * in[7 + c];
      if (r < 3) s += R[3 * r] * cl[0] + R[3 * r + 1] * cl[1] + R[3 * r + 2] * cl[2];
      out[lane] = s;
    }
    env_sync();
  };
  // dv[0..5] = Jsite in_arm ; dv[6..8] = in_v ; dv[9..11] = R in_omega
  auto dense_vel = [&](const float* in) {
    if (tid < 6) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 7; j++) s += w.ab[AB_JSITE + lane * 7 + j] * in[j];
      w.dv[lane] = s;
    } else if (tid < 9) {
      w.dv[lane] = dm.soft ? in[7 + lane - 6] : 0.f;
    } else if (tid < 12) {
      int r = lane - 9;
      w.dv[lane] = dm.soft ? R[3 * r] * in[10] + R[3 * r + 1] * in[11] + R[3 * r + 2] * in[12] : 0.f;
    }
    env_sync();
  };
  // out[j][c] = (J in)_c for the 3 rows of each contact (needs dense_vel(in) first)
  auto contactJ = [&](const float* in, float (*out)[DEV_MAXC], bool sub_aref) {
    v3 V = ld3(w.dv), W = ld3(w.dv + 3), iv = ld3(w.dv + 6), iw = ld3(w.dv + 9);
    PRAGMA_HOT
    for (int c = tid; c < ncon; c += NT) {
      int type = w.ctype[c], i = w.cpart[c];
      v3 pos = mk(w.cpos[0][c], w.cpos[1][c], w.cpos[2][c]), nn = mk(w.cn[0][c], w.cn[1][c], w.cn[2][c]), t1, t2;
      make_frame(nn, &t1, &t2);
      v3 rel = mk(0, 0, 0);
      if (type != 2) rel = rel - (iv + cross(iw, pos - P) + in[13 + i] * mv(R, ld3(pt.axis + 3 * i)));
      if (type != 0) rel = rel + V + cross(W, pos - site);
      float o0 = dot(nn, rel), o1 = dot(t1, rel), o2 = dot(t2, rel);
      if (sub_aref) { o0 -= w.cjv[0][c]; o1 -= w.cjv[1][c]; o2 -= w.cjv[2][c]; }
      out[0][c] = o0; out[1][c] = o1; out[2][c] = o2;
    }
    env_sync();
  };
  // grad = Hx - rhs - J^T f(jar); also returns probe wrench (force, torque about the site) via w.red[12..17]
  auto update_grad = [&]() -> bool {
    PRAGMA_HOT
    for (int i = tid; i < QPAD; i += NT) w.grad[i] = i < nv ? w.Hx[i] : 0.f;
    env_sync();
    bool changed = false;
    float g[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // particle-side (force, torque about P), probe-side (force, torque about site)
    PRAGMA_HOT
    for (int c = tid; c < ncon; c += NT) {
      int type = w.ctype[c], i = w.cpart[c], zone;
      float fr, mu, f0, f1, f2, Dn = w.cD[c];
      contact_params(type, fr, mu);
      cone_force(w.cjar[0][c], w.cjar[1][c], w.cjar[2][c], Dn, Dn * dm.impratio, mu, fr, f0, f1, f2, zone);
      changed |= (zone != (int)w.czone[c]);
      w.czone[c] = (unsigned char)zone;
      if (zone == 0) continue;
      v3 pos = mk(w.cpos[0][c], w.cpos[1][c], w.cpos[2][c]), nn = mk(w.cn[0][c], w.cn[1][c], w.cn[2][c]), t1, t2;
      make_frame(nn, &t1, &t2);
      v3 Fw = f0 * nn + f1 * t1 + f2 * t2; // force on geom2
      if (type != 2) {
        v3 Fp = -Fw, T = cross(pos - P, Fp);
        g[0] += Fp.x; g[1] += Fp.y; g[2] += Fp.z; g[3] += T.x; g[4] += T.y; g[5] += T.z;
        atomicAdd(&w.grad[13 + i], -dot(mv(R, ld3(pt.axis + 3 * i)), Fp));
      }
      if (type != 0) {
        v3 T = cross(pos - site, Fw);
        g[6] += Fw.x; g[7] += Fw.y; g[8] += Fw.z; g[9] += T.x; g[10] += T.y; g[11] += T.z;
      }
    }
    float g13[13];
#pragma unroll
    for (int k = 0; k < 12; k++) g13[k] = g[k];
    g13[12] = changed ? 1.f : 0.f;
    bsumk(g13);
#pragma unroll
    for (int k = 0; k < 12; k++) g[k] = g13[k];
    changed = g13[12] > 0.f;
    if (tid < 7) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 6; r++) s += w.ab[AB_JSITE + r * 7 + lane] * g[6 + r];
      // joint limit row
      float sg = w.lsign[lane];
      if (sg != 0.f) {
        float jar = sg * w.x[lane] - w.laref[lane];
        if (jar < 0.f) s += sg * (-w.lD[lane] * jar);
      }
      w.grad[lane] -= s;
    } else if (tid < 10 && dm.soft) {
      w.grad[lane] -= g[lane - 7];
    } else if (tid < 13 && dm.soft) {
      int r = lane - 10; // R^T torque
      w.grad[lane] -= R[r] * g[3] + R[3 + r] * g[4] + R[6 + r] * g[5];
    }
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) w.red[12 + k] = g[6 + k];
    }
    env_sync();
    return changed;
  };
  // pg = P^-1 grad  (arm: dense 7x7 Cholesky; torso: arrow with the 6x6 Schur complement Sf)
  auto precond = [&]() {
    float t[6] = {0, 0, 0, 0, 0, 0};
    if (dm.soft) {
      PRAGMA_HOT
      for (int i = tid; i < np; i += NT) {
        v3 ah = ld3(pt.axis + 3 * i), aw = mv(R, ah);
        v3 kk = mk(w.kx[i], w.ky[i], w.kz[i]);
        v3 bv = kk + dm.part_mass * aw;
        v3 cr = mv(R, ld3(pt.pos + 3 * i) + (w.qs[i] - dm.cap_r) * ah);
        v3 bw = mtv(R, cross(cr, kk));
        float gi = w.grad[13 + i] / w.dg[i];
        t[0] += bv.x * gi; t[1] += bv.y * gi; t[2] += bv.z * gi; t[3] += bw.x * gi; t[4] += bw.y * gi; t[5] += bw.z * gi;
      }
      bsumk(t);
      float y[6];
#pragma unroll
      for (int k = 0; k < 6; k++) y[k] = w.grad[7 + k] - t[k];
      chol_solve<6>(w.Sf, y);
      if (tid < 6) w.pg[7 + lane] = y[lane];
      PRAGMA_HOT
      for (int i = tid; i < np; i += NT) {
        v3 ah = ld3(pt.axis + 3 * i), aw = mv(R, ah);
        v3 kk = mk(w.kx[i], w.ky[i], w.kz[i]);
        v3 bv = kk + dm.part_mass * aw;
        v3 cr = mv(R, ld3(pt.pos + 3 * i) + (w.qs[i] - dm.cap_r) * ah);
        v3 bw = mtv(R, cross(cr, kk));
        float by = bv.x * y[0] + bv.y * y[1] + bv.z * y[2] + bw.x * y[3] + bw.y * y[4] + bw.z * y[5];
        w.pg[13 + i] = (w.grad[13 + i] - by) / w.dg[i];
      }
    }
    {
      float ya[7];
#pragma unroll
      for (int j = 0; j < 7; j++) ya[j] = w.grad[j];
      chol_solve<7>(w.Pa, ya);
      if (tid < 7) w.pg[lane] = ya[lane];
    }
    env_sync();
  };
  auto vdot = [&](const float* a, const float* b) {
    float s = 0.f;
PRAGMA_HOT
    for (int i = tid; i < nv; i += NT) s += a[i] * b[i];
    return bsum(s);
  };

  // ------------------------------------------------------------------ K6: nonlinear PCG; pass -1 evaluates the warm start
  for (int c = tid; c < ncon; c += NT) w.czone[c] = 255; // "unknown": the first update always reports a change
  for (int i = tid; i < QPAD; i += NT) w.s[i] = 0.f;
  env_sync();

  // preconditioner from the current active set (contact zones); rebuilt when the zones change
  auto build_precond = [&]() {
    // Runs 1-3 times per solve: written for SMALL CODE (rolled loops, stack arrays), not for speed, so that it does not
    // evict the CG loop body from the instruction cache.
#pragma unroll 1
    for (int i = tid; i < np; i += NT) { w.dg[i] = w.dg0[i]; w.kx[i] = 0.f; w.ky[i] = 0.f; w.kz[i] = 0.f; }
    env_sync();
    float acc[42]; // packed upper triangles of the 6x6 wrench-space Hessians: [0..20] torso side (about P), [21..41] probe side (about the site)
#pragma unroll
    for (int k = 0; k < 42; k++) acc[k] = 0.f;
#pragma unroll 1
    for (int c = tid; c < ncon; c += NT) {
      int zone = w.czone[c], type = w.ctype[c], i = w.cpart[c];
      if (zone == 0) continue;
      float Dn = w.cD[c];
      v3 pos = mk(w.cpos[0][c], w.cpos[1][c], w.cpos[2][c]);
      float F[9]; // contact frame rows: normal, tangent 1, tangent 2
      {
        v3 nn = mk(w.cn[0][c], w.cn[1][c], w.cn[2][c]), t1, t2;
        make_frame(nn, &t1, &t2);
        st3(F, nn); st3(F + 3, t1); st3(F + 6, t2);
      }
      float Hc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // Hessian of the cone cost wrt (jar_n, jar_t1, jar_t2)
      if (zone == 2) { // bottom zone: quadratic
        Hc[0] = Dn; Hc[4] = Dn * dm.impratio; Hc[8] = Dn * dm.impratio;
      } else { // middle zone: exact Hessian of 0.5 Dm (N - mu T)^2
        float fr, mu;
        contact_params(type, fr, mu);
        float N = w.cjar[0][c] * mu, U1 = w.cjar[1][c] * fr, U2 = w.cjar[2][c] * fr, T = fmaxf(sqrtf(U1 * U1 + U2 * U2), 1e-20f);
        float Dm = Dn / (mu * mu * (1.f + mu * mu)), NmT = N - mu * T, u1 = U1 / T, u2 = U2 / T;
        float g3[3] = {mu, -mu * u1 * fr, -mu * u2 * fr}, kk = -Dm * mu * NmT / T * fr * fr;
#pragma unroll
        for (int a2 = 0; a2 < 3; a2++)
#pragma unroll
          for (int b2 = 0; b2 < 3; b2++) Hc[3 * a2 + b2] = Dm * g3[a2] * g3[b2];
        Hc[4] += kk * (1.f - u1 * u1); Hc[5] -= kk * u1 * u2; Hc[7] -= kk * u1 * u2; Hc[8] += kk * (1.f - u2 * u2);
      }
      if (type != 2) { // slider of this particle: k_i += K a_i, dg_i += a_i^T K a_i with K = F^T Hc F
        v3 aw = mv(R, ld3(pt.axis + 3 * i));
        v3 fa = mv(F, aw), hf = mv(Hc, fa), ka = mtv(F, hf);
        atomicAdd(&w.kx[i], ka.x); atomicAdd(&w.ky[i], ka.y); atomicAdd(&w.kz[i], ka.z);
        atomicAdd(&w.dg[i], dot(fa, hf));
      }
      // wrench-space Hessian A^T K A with A = [I, -[r]x]: W[a] = [f_a ; r x f_a], U = Hc W, acc += W^T U (upper triangle)
      auto accum = [&](float* A21, v3 r) {
        float W[18], U[18];
#pragma unroll
        for (int a2 = 0; a2 < 3; a2++) {
          v3 f = ld3(F + 3 * a2), m3 = cross(r, f);
          st3(W + 6 * a2, f); st3(W + 6 * a2 + 3, m3);
        }
#pragma unroll
        for (int a2 = 0; a2 < 3; a2++)
#pragma unroll
          for (int q = 0; q < 6; q++) U[6 * a2 + q] = Hc[3 * a2] * W[q] + Hc[3 * a2 + 1] * W[6 + q] + Hc[3 * a2 + 2] * W[12 + q];
        int idx = 0;
#pragma unroll
        for (int p2 = 0; p2 < 6; p2++)
#pragma unroll
          for (int q = p2; q < 6; q++) { A21[idx] += W[p2] * U[q] + W[6 + p2] * U[6 + q] + W[12 + p2] * U[12 + q]; idx++; }
      };
      if (type != 2) accum(acc, pos - P);
      if (type != 0) accum(acc + 21, pos - site);
    }
    {
      float(&lo21)[21] = *reinterpret_cast<float(*)[21]>(acc);
      float(&hi21)[21] = *reinterpret_cast<float(*)[21]>(acc + 21);
      bsumk(lo21);
      bsumk(hi21);
    }
    env_sync();
    // arm block: Pa = M + Jsite^T Kp Jsite + limits   (lanes 0..6, column `lane`)
    if (tid < 7) {
      float Kp6[36];
      {
        int idx = 21;
#pragma unroll
        for (int a2 = 0; a2 < 6; a2++)
#pragma unroll
          for (int b2 = a2; b2 < 6; b2++) { Kp6[a2 * 6 + b2] = acc[idx]; Kp6[b2 * 6 + a2] = acc[idx]; idx++; }
      }
      float KJ[6]; // (Kp Jsite)[:, lane]
#pragma unroll
      for (int a2 = 0; a2 < 6; a2++) {
        float sacc = 0.f;
#pragma unroll
        for (int b2 = 0; b2 < 6; b2++) sacc += Kp6[a2 * 6 + b2] * w.ab[AB_JSITE + b2 * 7 + lane];
        KJ[a2] = sacc;
      }
#pragma unroll
      for (int r2 = 0; r2 < 7; r2++) {
        float sacc = w.ab[AB_M + r2 * 7 + lane];
#pragma unroll
        for (int a2 = 0; a2 < 6; a2++) sacc += w.ab[AB_JSITE + a2 * 7 + r2] * KJ[a2];
        if (r2 == lane && w.lsign[lane] != 0.f) sacc += w.lD[lane];
        w.Pa[r2 * 7 + lane] = sacc;
      }
    }
    // torso block: Sf = Mff + Kf(local) - sum_i b_i b_i^T / dg_i
    if (dm.soft) {
      float sb[21];
#pragma unroll
      for (int k = 0; k < 21; k++) sb[k] = 0.f;
#pragma unroll 1
      for (int i = tid; i < np; i += NT) {
        v3 ah = ld3(pt.axis + 3 * i), aw = mv(R, ah);
        v3 kk = mk(w.kx[i], w.ky[i], w.kz[i]);
        v3 bv = kk + dm.part_mass * aw;
        v3 cr = mv(R, ld3(pt.pos + 3 * i) + (w.qs[i] - dm.cap_r) * ah);
        v3 bw = mtv(R, cross(cr, kk));
        float b6[6] = {bv.x, bv.y, bv.z, bw.x, bw.y, bw.z}, inv = 1.f / w.dg[i];
        int idx = 0;
#pragma unroll
        for (int a2 = 0; a2 < 6; a2++)
#pragma unroll
          for (int c2 = a2; c2 < 6; c2++) { sb[idx] += b6[a2] * b6[c2] * inv; idx++; }
      }
      bsumk(sb);
      if (tid == 0) {
        // rotate the angular part of Kf to the body frame: T = diag(I, R); Kl = T^T Kf T.  Scratch: hs[0..95] (dead here)
        float* Kf = w.hs; float* Kl = w.hs + 36; float* sbs = w.hs + 72;
        {
          int idx = 0;
#pragma unroll
          for (int a2 = 0; a2 < 6; a2++)
#pragma unroll
            for (int b2 = a2; b2 < 6; b2++) { Kf[a2 * 6 + b2] = acc[idx]; Kf[b2 * 6 + a2] = acc[idx]; sbs[idx] = sb[idx]; idx++; }
        }
        int idx = 0;
#pragma unroll 1
        for (int a2 = 0; a2 < 3; a2++)
#pragma unroll 1
          for (int b2 = 0; b2 < 3; b2++) {
            Kl[a2 * 6 + b2] = Kf[a2 * 6 + b2];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int k = 0; k < 3; k++) s1 += Kf[a2 * 6 + 3 + k] * w.R[3 * k + b2];
            Kl[a2 * 6 + 3 + b2] = s1; Kl[(3 + b2) * 6 + a2] = s1;
#pragma unroll 1
            for (int k = 0; k < 3; k++)
#pragma unroll 1
              for (int l = 0; l < 3; l++) s2 += w.R[3 * k + a2] * Kf[(3 + k) * 6 + 3 + l] * w.R[3 * l + b2];
            Kl[(3 + a2) * 6 + 3 + b2] = s2;
          }
#pragma unroll 1
        for (int attempt = 0; attempt < 2; attempt++) {
          idx = 0;
#pragma unroll 1
          for (int a2 = 0; a2 < 6; a2++)
#pragma unroll 1
            for (int b2 = a2; b2 < 6; b2++) {
              float v = w.Mff[a2 * 6 + b2] + Kl[a2 * 6 + b2] - (attempt == 0 ? sbs[idx] : 0.f);
              idx++;
              w.Sf[a2 * 6 + b2] = v; w.Sf[b2 * 6 + a2] = v;
            }
          if (chol_rolled(w.Sf, 6)) break;
          // fall back to the free block without the slider coupling
#pragma unroll 1
          for (int i = 0; i < np; i++) { w.kx[i] = 0.f; w.ky[i] = 0.f; w.kz[i] = 0.f; }
        }
      }
    }
    env_sync();
    if (tid == 0) chol_rolled(w.Pa, 7);
    env_sync();
  };

  float gpg = 1.f, gnorm = 0.f, rhsn = 0.f;
  int iters = 0, rebuilds = 0;
  const int maxit = mode == 1 ? 2 * dm.iters : dm.iters;
  // every helper has exactly ONE call site (code size: the loop body must stay inside the instruction cache)
#pragma unroll 1
  for (int it = -1; it < maxit; it++) {
    const bool init = it < 0;
    if (!init) {
      // fp32 floor of the gradient is ~eps * (|Hx| + |rhs|): the terms that cancel in it
      if (gnorm <= dm.tol * (1.f + rhsn + sqrtf(vdot(w.Hx, w.Hx)))) break;
      iters = it + 1;
    }
    const float* vin = init ? w.x : w.s;
    float* vout = init ? w.Hx : w.hs;
    applyH(vin, vout);
    dense_vel(vin);
    contactJ(vin, init ? w.cjar : w.cjv, init);
    if (init) {
      rhsn = sqrtf(vdot(w.grad, w.grad)); // grad holds rhs until here
      PRAGMA_HOT
      for (int i = tid; i < nv; i += NT) w.Hx[i] -= w.grad[i];
      env_sync();
    } else {
      // ---- exact line search: Newton on phi'(alpha)
      float q1 = 0.f, q2 = 0.f;
      PRAGMA_HOT
      for (int i = tid; i < nv; i += NT) { q1 += w.s[i] * w.Hx[i]; q2 += w.s[i] * w.hs[i]; }
      {
        float r2[2] = {q1, q2};
        bsumk(r2);
        q1 = r2[0]; q2 = r2[1];
      }
      float alpha = 0.f, lo = 0.f, hi = -1.f, d0abs = 0.f;
#pragma unroll 1
      for (int ls = 0; ls < 8; ls++) {
        float d1 = 0.f, d2 = 0.f;
        PRAGMA_HOT
        for (int c = tid; c < ncon; c += NT) {
          float fr, mu, a1, a2, Dn = w.cD[c];
          contact_params(w.ctype[c], fr, mu);
          cone_ls(w.cjar[0][c] + alpha * w.cjv[0][c], w.cjar[1][c] + alpha * w.cjv[1][c], w.cjar[2][c] + alpha * w.cjv[2][c],
                  w.cjv[0][c], w.cjv[1][c], w.cjv[2][c], Dn, Dn * dm.impratio, mu, fr, a1, a2);
          d1 += a1; d2 += a2;
        }
        if (tid < 7 && w.lsign[lane] != 0.f) {
          float sg = w.lsign[lane], jar = sg * (w.x[lane] + alpha * w.s[lane]) - w.laref[lane], jv = sg * w.s[lane];
          if (jar < 0.f) { d1 += w.lD[lane] * jar * jv; d2 += w.lD[lane] * jv * jv; }
        }
        {
          float r2[2] = {d1, d2};
          bsumk(r2);
          d1 = r2[0] + q1 + alpha * q2;
          d2 = r2[1] + q2;
        }
        if (ls == 0) d0abs = fabsf(d1);
        if (fabsf(d1) <= 1e-5f * d0abs || !(d2 > 0.f)) break;
        if (d1 < 0.f) lo = alpha; else hi = alpha;
        float an = alpha - d1 / d2;
        if (hi < 0.f) { if (an <= lo) an = 2.f * alpha + 1e-6f; }
        else if (an <= lo || an >= hi) an = 0.5f * (lo + hi);
        if (an == alpha) break;
        alpha = an;
      }
      PRAGMA_HOT
      for (int i = tid; i < nv; i += NT) { w.x[i] += alpha * w.s[i]; w.Hx[i] += alpha * w.hs[i]; }
      PRAGMA_HOT
      for (int c = tid; c < ncon; c += NT) {
        w.cjar[0][c] += alpha * w.cjv[0][c]; w.cjar[1][c] += alpha * w.cjv[1][c]; w.cjar[2][c] += alpha * w.cjv[2][c];
      }
      env_sync();
    }
    bool changed = update_grad();
    float gpo = init ? 0.f : vdot(w.grad, w.pg); // with the previous pg (Polak-Ribiere)
    bool restart = init;
    // soft scene: rebuild when the active set moved; rigid scene (7 unknowns): exact Hessian every iteration = Newton
    if (init || (changed && rebuilds < dm.max_rebuilds) || (!dm.soft && ncon > 0)) {
      build_precond();
      rebuilds += init ? 0 : 1;
      restart = true;
    }
    precond();
    float gpn = vdot(w.grad, w.pg);
    gnorm = sqrtf(vdot(w.grad, w.grad));
    float beta = restart ? 0.f : fmaxf(0.f, (gpn - gpo) / fmaxf(gpg, 1e-30f));
    gpg = gpn;
    PRAGMA_HOT
    for (int i = tid; i < nv; i += NT) w.s[i] = -w.pg[i] + beta * w.s[i];
    env_sync();
  }

  // ------------------------------------------------------------------ K8: probe wrench, F/T torque
  v3 cfrc = mk(w.red[12], w.red[13], w.red[14]), ctq = mk(w.red[15], w.red[16], w.red[17]);
  v3 ft;
  {
    float t3[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      float s = w.ab[AB_TAU0 + r];
#pragma unroll
      for (int j = 0; j < 7; j++) s += w.ab[AB_JFT + r * 7 + j] * w.x[j];
      t3[r] = s;
    }
    ft = mtv(w.ab + AB_EEFR, mk(t3[0], t3[1], t3[2]) - ctq);
  }
  bool in_contact = false;
  for (int c = tid; c < ncon; c += NT) in_contact |= (w.ctype[c] == 1);
  in_contact = bsum(in_contact ? 1.f : 0.f) > 0.f;

  // ------------------------------------------------------------------ K7: integrate (mj_Euler) and write the state back
  if (mode == 0) {
    // arm: implicit joint damping, (M + h D) qacc' = M qacc, through a dense 7x7 Cholesky on lane 0
    if (tid == 0) {
      float A[49], b[7];
#pragma unroll
      for (int r = 0; r < 7; r++) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 7; j++) { A[r * 7 + j] = w.ab[AB_M + r * 7 + j] + (r == j ? h * dm.arm_damp : 0.f); s += w.ab[AB_M + r * 7 + j] * w.x[j]; }
        b[r] = s;
      }
      chol<7>(A);
      chol_solve<7>(A, b);
#pragma unroll
      for (int j = 0; j < 7; j++) w.qdarm[j] += h * b[j];
    }
    env_sync();
    for (int i = tid; i < nv; i += NT) wm_g[i] = w.x[i];
    if (tid < 7) { qv_g[tid] = w.qdarm[tid]; qp_g[lane] += h * w.qdarm[tid]; }
    if (dm.soft) {
      for (int i = tid; i < np; i += NT) {
        float v = qv_g[13 + i] + h * w.x[13 + i];
        qv_g[13 + i] = v;
        qp_g[14 + i] = w.qs[i] + h * v;
      }
      if (tid == 0) {
        v3 vn = vlin + h * ld3(w.x + 7), wn = wl + h * ld3(w.x + 10);
        st3(qv_g + 7, vn); st3(qv_g + 10, wn);
        st3(qp_g + 7, P + h * vn);
        float wnm = norm(wn), ang = h * wnm;
        if (ang > 0.f) {
          float s = sinf(0.5f * ang) / wnm, qr[4] = {cosf(0.5f * ang), s * wn.x, s * wn.y, s * wn.z}, qn[4];
          quatmul(quat, qr, qn);
          float nq = rsqrtf(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
          qp_g[10] = qn[0] * nq; qp_g[11] = qn[1] * nq; qp_g[12] = qn[2] * nq; qp_g[13] = qn[3] * nq;
        } else {
          qp_g[10] = quat[0]; qp_g[11] = quat[1]; qp_g[12] = quat[2]; qp_g[13] = quat[3];
        }
      }
    }
  }
  env_sync();

  // ------------------------------------------------------------------ contact list / diagnostics
  if (ncon_out) {
    if (tid == 0) ncon_out[env] = ncon;
    for (int c = tid; c < ncon; c += NT) {
      int type = w.ctype[c], g1, g2;
      if (type == 2) { g1 = 1; g2 = 2; } else { g1 = 4 + w.cpart[c]; g2 = type == 0 ? 1 : 2; }
      geom1_out[(size_t)env * DEV_MAXC + c] = g1;
      geom2_out[(size_t)env * DEV_MAXC + c] = g2;
      if (dist_out) dist_out[(size_t)env * DEV_MAXC + c] = w.cdist[c];
    }
  }

  // divergence guard (MuJoCo resets on bad qacc): a non-finite solution ends the episode, the reset wipes the state
  const float xnorm2 = vdot(w.x, w.x);
  // ------------------------------------------------------------------ K9: task epilogue (lane 0)
  if (tid == 0) {
    float* ts = w.ts;
    v3 hv = mk(0, 0, 0);
    {
      float t3[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 7; j++) s += w.ab[AB_JHAND + r * 7 + j] * w.qdarm[j];
        t3[r] = s;
      }
      hv = mk(t3[0], t3[1], t3[2]);
    }
    v3 eef = ld3(w.ab + AB_EEFPOS);
    const float* quat_e = w.ab + AB_QUAT;
    float reward = 0.f;
    int dn = 0;
    const bool bad = !isfinite(xnorm2) || !isfinite(cfrc.x + cfrc.y + cfrc.z) || !isfinite(ft.x + ft.y + ft.z) || !isfinite(hv.x + hv.y + hv.z) ||
                     !isfinite(eef.x + eef.y + eef.z);
    if (bad) { // keep the outputs finite; the episode ends below and the reset wipes the state
      cfrc = mk(0, 0, 0); ft = mk(0, 0, 0); hv = mk(0, 0, 0);
      if (!isfinite(eef.x + eef.y + eef.z)) eef = ld3(ts + USIM_TS_TRAJ_PT);
      if (diverged) atomicAdd(diverged, 1);
    }
    if (mode == 0) {
      ts[USIM_TS_TIMESTEP] += 1.f;
      if (in_contact) ts[USIM_TS_TOUCHED] = 1.f;
      float pe[2], oe;
      reward = reward_fn(eef, quat_e, ld3(ts + USIM_TS_TRAJ_PT), ts[USIM_TS_VEL_MEAN], ts[USIM_TS_FZ_MEAN], ts[USIM_TS_DFZ], in_contact, pe, &oe);
      ts[USIM_TS_POS_ERR] = pe[0]; ts[USIM_TS_POS_ERR + 1] = pe[1]; ts[USIM_TS_ORI_ERR] = oe;
      float t = ts[USIM_TS_TIMESTEP];
      dn = t >= (float)dm.horizon;
      float u = fminf(fmaxf(t / (float)dm.horizon + ts[USIM_TS_U0], 0.f), 1.f); // ultrasound.py:528-532
#pragma unroll
      for (int k = 0; k < 3; k++) ts[USIM_TS_TRAJ_PT + k] = ts[USIM_TS_TRAJ_START + k] + u * (ts[USIM_TS_TRAJ_END + k] - ts[USIM_TS_TRAJ_START + k]);
      ts[USIM_TS_VEL_MEAN] += (norm(hv) - ts[USIM_TS_VEL_MEAN]) / t;   // :538
      ts[USIM_TS_DFZ] = (cfrc.z - ts[USIM_TS_FZ_PREV]) * dm.ctrl_freq;  // :542
      ts[USIM_TS_FZ_PREV] = cfrc.z;
      ts[USIM_TS_FZ_MEAN] = 0.1f * cfrc.z + 0.9f * ts[USIM_TS_FZ_MEAN]; // :546
      if (dm.early_term) { // :635-670
        int term = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
          float qn = qp_g[j];
          if (!(dm.jnt_lo[j] + 0.1f < qn && qn < dm.jnt_hi[j] - 0.1f)) term = 1;
        }
        if (sqrtf(pe[0] * pe[0] + pe[1] * pe[1]) > 1.0f) term = 1;
        if (in_contact && oe > 0.10f) term = 1;
        if (ts[USIM_TS_TOUCHED] != 0.f && !in_contact) term = 1;
        dn = dn || term;
      }
      if (bad) dn = 1;
      ts[USIM_TS_DONE] = (float)dn;
    } else {
      ts[USIM_TS_FZ_PREV] = 0.f; ts[USIM_TS_DFZ] = 0.f;
      ts[USIM_TS_VEL_MEAN] = norm(hv);   // :474
      ts[USIM_TS_FZ_MEAN] = cfrc.z;      // :477
      ts[USIM_TS_TOUCHED] = 0.f; ts[USIM_TS_TIMESTEP] = 0.f; ts[USIM_TS_DONE] = 0.f;
    }
    ts[USIM_TS_IN_CONTACT] = in_contact ? 1.f : 0.f;
    if (obs) { // ultrasound.py:363-401
      float* o = obs + (size_t)env * USIM_OBS_DIM;
      o[0] = cfrc.x; o[1] = cfrc.y; o[2] = cfrc.z; o[3] = ft.x; o[4] = ft.y; o[5] = ft.z; o[6] = hv.x; o[7] = hv.y; o[8] = hv.z;
      o[9] = ts[USIM_TS_FZ_MEAN] - 5.f; o[10] = ts[USIM_TS_DFZ]; o[11] = ts[USIM_TS_VEL_MEAN] - 0.04f;
      o[12] = eef.x - ts[USIM_TS_TRAJ_PT]; o[13] = eef.y - ts[USIM_TS_TRAJ_PT + 1]; o[14] = eef.z - ts[USIM_TS_TRAJ_PT + 2];
      float gq[4] = {GQX, GQY, GQZ, GQW};
      difference_quat(quat_e, gq, o + 15); // xyzw arrays through a wxyz routine (:390)
    }
    if (mode == 0) {
      if (rew) rew[env] = reward;
      if (done) done[env] = (uint8_t)dn;
    }
    if (diag) {
      float* d = diag + (size_t)env * USIM_DIAG_DIM;
      d[0] = cfrc.x; d[1] = cfrc.y; d[2] = cfrc.z; d[3] = ft.x; d[4] = ft.y; d[5] = ft.z; d[6] = eef.x; d[7] = eef.y; d[8] = eef.z;
      d[9] = quat_e[0]; d[10] = quat_e[1]; d[11] = quat_e[2]; d[12] = quat_e[3];
#pragma unroll
      for (int j = 0; j < 7; j++) d[13 + j] = w.ab[AB_TAU + j];
      int nlim = 0;
      for (int j = 0; j < 7; j++) nlim += w.lsign[j] != 0.f;
      d[20] = (float)iters; d[21] = gnorm; d[22] = (float)ncon_found;
      d[23] = (float)(3 * ncon + nlim + (dm.soft ? 2 * 0 + np + dm.npair + 1 : 0));
    }
  }
  env_sync();
  for (int i = tid; i < USIM_TASK_DIM; i += NT) ts_g[i] = w.ts[i];
}
