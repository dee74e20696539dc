// common.cuh — device constants, small vector math, Philox, task-layer device functions.
// fp32 throughout; sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/usim.h"

#define DEV_MAXC USIM_MAX_CONTACTS
#define NPART_MAX 272
#define QPAD 288   // row pitch (floats) of qpos / qvel / warm in HBM: 284 -> 288 = 9 x 128 B
#define ARMBUF 180 // row pitch of the arm intermediate record (a multiple of 4 floats: rows are moved by 16-byte bulk copies)

// arm-record layout (floats), written by the arm kernel, read by the solve kernel
enum {
  AB_M = 0,      // 49 joint-space inertia, row major
  AB_QS = 49,    // 7  qfrc_smooth = tau - bias - damping*qvel
  AB_TAU = 56,   // 7  clipped controller torques
  AB_JSITE = 63, // 42 grip-site Jacobian (rows 0-2 Jp, 3-5 Jr)
  AB_JHAND = 105, // 21 hand-body linear Jacobian
  AB_EEFPOS = 126, // 3
  AB_EEFR = 129,   // 9
  AB_PTIP = 138,   // 3 probe capsule tip-sphere centre (world)
  AB_PBACK = 141,  // 3
  AB_TAU0 = 144,   // 3 F/T torque: velocity-product + gravity part about the site (world)
  AB_JFT = 147,    // 21 F/T torque: linear map from arm qacc (world)
  AB_QUAT = 168,   // 4 eef quaternion xyzw, w >= 0
  AB_DX = 172      // 7 warm-start shift of the arm: M^-1 (qfrc_smooth - previous qfrc_smooth)
};

struct DevModel {
  // arm
  float link_pos[7][3], link_R[7][9], link_com[7][3], link_mass[7], link_I[7][6];
  float tool[34];
  float jnt_lo[7], jnt_hi[7], ctrl[7], init_qpos[7], iw_arm[7];
  float arm_damp;
  // options
  float h, g[3], impratio, solref[2], solimp[5], solref_smooth[2];
  float table_z, table_half, fr_table_probe, fr_table_part, fr_probe_part, probe_r, cap_r, cap_hl;
  float iw_probe, iw_table; // body_invweight0 (translational)
  // torso
  int soft, npart, npair;
  float tendon_iw, free_damp, part_mass, center_mass;
  float torso_qpos0[7];
  float top_offset, traj_xr, traj_yr; // trajectory grid on the torso top (ultrasound.py:184-186)
  float rot_I[6];          // constant rotational inertia (xx,yy,zz,xy,xz,yz): capsules about their COM + centre geom
  // config
  int mode, horizon, early_term, solref_rand, pos_rand, det_traj, uncouple, iters, adim, env_off, nq, nv, max_rebuilds, ignore_done;
  unsigned seed_lo, seed_hi;
  float ctrl_freq, kp[6], dr[6], in_max, in_min, out_max[6], out_min[6], kp_lim[2], kp_in_max, kp_in_min, tol;
  float eef_bias[3];
};

// per-particle tables in global memory (read only, shared by every env: L1/L2 resident), one 16-byte load per record
struct PartTables {
  const float4* ax4; // [npart] slider axis (torso frame) xyz, w = dof_invweight0 of the slider
  const float4* ps4; // [npart] rest position (torso frame) xyz, w = body_invweight0 (translational)
  const int4* nb4;   // [npart] up to 4 grid neighbours, packed (pair << 16) | (second << 15) | neighbour (second: this slider is the
                     // pair's second element, its row enters with a minus sign); empty slots = (npair << 16) | self
};

__constant__ DevModel dm;  // single translation unit (usim.cu)

// ---------------------------------------------------------------- vec3
struct v3 {
  float x, y, z;
};
__device__ __forceinline__ v3 mk(float x, float y, float z) { return v3{x, y, z}; }
__device__ __forceinline__ v3 ld3(const float* p) { return v3{p[0], p[1], p[2]}; }
__device__ __forceinline__ v3 xyz(float4 a) { return v3{a.x, a.y, a.z}; }
__device__ __forceinline__ void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return v3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return v3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ v3 operator-(v3 a) { return v3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ v3 operator*(float s, v3 a) { return v3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ v3 operator*(v3 a, float s) { return v3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ v3 cross(v3 a, v3 b) { return v3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float norm(v3 a) { return sqrtf(dot(a, a)); }
// 3x3 row-major
__device__ __forceinline__ v3 mv(const float* R, v3 a) {
  return v3{R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z, R[6] * a.x + R[7] * a.y + R[8] * a.z};
}
__device__ __forceinline__ v3 mtv(const float* R, v3 a) {
  return v3{R[0] * a.x + R[3] * a.y + R[6] * a.z, R[1] * a.x + R[4] * a.y + R[7] * a.z, R[2] * a.x + R[5] * a.y + R[8] * a.z};
}
__device__ __forceinline__ void mm3(const float* A, const float* B, float* C) {
  float r[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) C[i] = r[i];
}
__device__ __forceinline__ void quat2mat(const float* q, float* R) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
// symmetric 3x3 stored (xx,yy,zz,xy,xz,yz)
__device__ __forceinline__ v3 symv(const float* I, v3 a) {
  return v3{I[0] * a.x + I[3] * a.y + I[4] * a.z, I[3] * a.x + I[1] * a.y + I[5] * a.z, I[4] * a.x + I[5] * a.y + I[2] * a.z};
}

// in-place Cholesky of a dense n x n (row-major, lower), n compile-time.  Every loop has a CONSTANT trip count with a
// predicate inside: triangular bounds made the compiler keep inner loops rolled and index the (register) matrix dynamically,
// which put it in local memory.
template <int N>
__device__ __forceinline__ bool chol(float* A) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; j++) {
    float s = A[j * N + j];
#pragma unroll
    for (int k = 0; k < N; k++)
      if (k < j) s -= A[j * N + k] * A[j * N + k];
    if (!(s > 0.f)) { ok = false; s = 1e-20f; }
    s = sqrtf(s);
    A[j * N + j] = s;
    float inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (i > j) {
        float t = A[i * N + j];
#pragma unroll
        for (int k = 0; k < N; k++)
          if (k < j) t -= A[i * N + k] * A[j * N + k];
        A[i * N + j] = t * inv;
      }
    }
  }
  return ok;
}
template <int N>
__device__ __forceinline__ void chol_solve(const float* L, float* x) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    float t = x[i];
#pragma unroll
    for (int k = 0; k < N; k++)
      if (k < i) t -= L[i * N + k] * x[k];
    x[i] = t / L[i * N + i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    float t = x[i];
#pragma unroll
    for (int k = 0; k < N; k++)
      if (k > i) t -= L[k * N + i] * x[k];
    x[i] = t / L[i * N + i];
  }
}

// x <- (L L^T)^-1 x for the leading N x N block of a factor written by chol7_warp2 / chol7_group: lower triangle, diagonal INVERTED (row pitch 7)
template <int N>
__device__ __forceinline__ void chol7_solve(const float* L, float* x) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    float t = x[i];
#pragma unroll
    for (int k = 0; k < i; k++) t -= L[i * 7 + k] * x[k];
    x[i] = t * L[i * 7 + i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    float t = x[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) t -= L[k * 7 + i] * x[k];
    x[i] = t * L[i * 7 + i];
  }
}

// ---------------------------------------------------------------- Philox4x32-10 (bit-exact with the oracle)
__device__ __forceinline__ void philox(unsigned k0, unsigned k1, unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---------------------------------------------------------------- task layer (ultrasound.py, utils/quaternion.py)
#define GQX (-0.69192486f)
#define GQY (0.72186726f)
#define GQZ (-0.00514253f)
#define GQW (-0.01100909f)

__device__ __forceinline__ void quatmul(const float* a, const float* b, float* o) {
  float r0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float r1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float r2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float r3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
}
// utils/quaternion.py:23-35
__device__ __forceinline__ void difference_quat(const float* q1, const float* q2, float* o) {
  float c[4] = {q2[0], -q2[1], -q2[2], -q2[3]};
  quatmul(q1, c, o);
}
// utils/quaternion.py:38-59 (+ q_log :4-20)
__device__ __forceinline__ float distance_quat(const float* q1, const float* q2) {
  float m[4];
  difference_quat(q1, q2, m);
  float v = fminf(fmaxf(m[0], -1.f), 1.f);
  float un = sqrtf(m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
  float d = 0.f;
  if (un != 0.f) d = 2.f * fabsf(acosf(v)); // |acos(v) * u/|u|| = acos(v)
  const float PI_F = 3.14159265358979f;
  if (d > PI_F) d = fabsf(2.f * PI_F - d);
  return d;
}
// robosuite mat2quat: (x,y,z,w) with w >= 0
__device__ __forceinline__ void mat2quat_xyzw(const float* m, float* q) {
  float tr = m[0] + m[4] + m[8], w, x, y, z;
  if (tr > 0) {
    float s = sqrtf(tr + 1.f) * 2;
    w = 0.25f * s; x = (m[7] - m[5]) / s; y = (m[2] - m[6]) / s; z = (m[3] - m[1]) / s;
  } else if (m[0] > m[4] && m[0] > m[8]) {
    float s = sqrtf(1.f + m[0] - m[4] - m[8]) * 2;
    w = (m[7] - m[5]) / s; x = 0.25f * s; y = (m[1] + m[3]) / s; z = (m[2] + m[6]) / s;
  } else if (m[4] > m[8]) {
    float s = sqrtf(1.f + m[4] - m[0] - m[8]) * 2;
    w = (m[2] - m[6]) / s; x = (m[1] + m[3]) / s; y = 0.25f * s; z = (m[5] + m[7]) / s;
  } else {
    float s = sqrtf(1.f + m[8] - m[0] - m[4]) * 2;
    w = (m[3] - m[1]) / s; x = (m[2] + m[6]) / s; y = (m[5] + m[7]) / s; z = 0.25f * s;
  }
  if (w < 0) { w = -w; x = -x; y = -y; z = -z; }
  q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}
__device__ __forceinline__ void goal_mat(float* R) {
  float q[4] = {GQW, GQX, GQY, GQZ};
  float n = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] *= n;
  quat2mat(q, R);
}
// ultrasound.py:230-269
__device__ __forceinline__ float reward_fn(v3 eef, const float* quat_xyzw, v3 traj, float vel_mean, float fz_mean, float dfz,
                                           bool in_contact, float* pe, float* oe) {
  float cur[4] = {quat_xyzw[3], quat_xyzw[0], quat_xyzw[1], quat_xyzw[2]};
  float des[4] = {GQW, GQX, GQY, GQZ};
  float e0 = 90.f * (eef.x - traj.x), e1 = 90.f * (eef.y - traj.y);
  pe[0] = e0 * e0; pe[1] = e1 * e1;
  float r = 5.f * expf(-sqrtf(pe[0] * pe[0] + pe[1] * pe[1]));
  *oe = 0.2f * distance_quat(cur, des);
  r += expf(-*oe);
  float ve = 45.f * (vel_mean - 0.04f);
  r += expf(-ve * ve);
  if (in_contact) {
    float fe = 0.7f * (fz_mean - 5.f), de = 0.01f * dfz;
    r += 3.f * expf(-fe * fe) + 2.f * expf(-de * de);
  }
  return r;
}
__device__ __forceinline__ float scale1(float a, float imin, float imax, float omin, float omax) {
  a = fminf(fmaxf(a, imin), imax);
  return (a - 0.5f * (imax + imin)) * (fabsf(omax - omin) / fabsf(imax - imin)) + 0.5f * (omax + omin);
}
// 0.5 sum_i cur[:,i] x des[:,i]
__device__ __forceinline__ v3 ori_error(const float* des, const float* cur) {
  v3 e = mk(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 3; i++) e = e + 0.5f * cross(mk(cur[i], cur[3 + i], cur[6 + i]), mk(des[i], des[3 + i], des[6 + i]));
  return e;
}
// MuJoCo impedance + reference-acceleration gains
__device__ __forceinline__ void kbi(float sr0, float sr1, float pos, float* K, float* B, float* imp) {
  float dmin = dm.solimp[0], dmax = dm.solimp[1], width = dm.solimp[2], mid = dm.solimp[3], power = dm.solimp[4];
  float x = fabsf(pos) / width, y;
  if (x >= 1.f) y = 1.f;
  else if (x <= mid) y = __powf(x, power) / __powf(mid, power - 1.f);
  else y = 1.f - __powf(1.f - x, power) / __powf(1.f - mid, power - 1.f);
  if (x == 0.f) y = 0.f;
  float d = dmin + y * (dmax - dmin);
  d = fminf(fmaxf(d, 1e-4f), 0.9999f);
  *imp = d;
  if (sr0 > 0.f) {
    float tc = fmaxf(sr0, 2.f * dm.h);
    *K = 1.f / (dmax * dmax * tc * tc * sr1 * sr1);
    *B = 2.f / (dmax * tc);
  } else {
    *K = -sr0 / (dmax * dmax);
    *B = -sr1 / dmax;
  }
}
