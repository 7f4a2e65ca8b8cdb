// armsim_device.cuh -- device-side building blocks of the fused env-step kernels (sm_100a, fp32).
//
// Everything here is per-arm scalar math kept entirely in registers; the robot chain and the task constants arrive as
// __grid_constant__ kernel parameters, i.e. they live in the constant bank and feed FFMA operands directly.
// Reference behaviour being reproduced: envs/rl_reach_env.py:219-319 (+ push / pick / kuka_reach variants) on top of
// Bullet's calculateInverseKinematics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "armsim.h"

#define NJ ARMSIM_NJ

struct ChainParams {
  float Rb[9], tb[3];        // base transform
  float Rf[NJ][9];           // fixed rotation of each joint origin (R = Rz(y)Ry(p)Rx(r)), row-major
  float t[NJ][3];            // joint origin translation in the parent link frame
  float lower[NJ], upper[NJ];
};

struct TaskParams {
  int task, n, max_steps, auto_reset, ik_max_iters, napply, clamp, obs_dim, torque_mode;
  float dv, reach_dis, ik_damping, ik_residual;
  float ws_lo[3], ws_hi[3];
  float goal_lo[3], goal_span[3];
  float tquat[4];            // IK target orientation, xyzw
  float tR[9];               // the same orientation as a row-major rotation matrix
  float init_q[NJ];
  float init_ee[3], init_R[9];   // FK(init_q) computed once on the host in fp64 (every episode starts there)
  uint32_t seed_lo, seed_hi;
  unsigned long long gid_offset;
};

// per-env state, struct-of-arrays: field k of env e lives at ptr[k * n + e]
struct StatePtrs {
  float* q;          // [7][n]
  float* qd;         // [7][n]   torque mode
  float* goal;       // [3][n]
  int* step;         // [n]
  int* episode;      // [n]
  int* ik_iters;     // [n]
  uint8_t* done;     // [n]     latched done flag (auto_reset = 0)
  float* cube;       // [13][n] pos3 quat4 v3 w3
  float* last_dist;  // [n]
  float* grip;       // [n]
  float* ep_return;  // [n]     return of the running episode (armsim_track_episodes)
  unsigned int* explore_count;  // [n]  exploration-noise draws so far (armsim_explore)
};

// ------------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void reset_uniforms(const TaskParams& T, unsigned long long gid, uint32_t episode, uint32_t block,
                                               float (&u)[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), episode, block, T.seed_lo, T.seed_hi, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = __fmul_rn((float)(r[i] >> 8), 5.9604644775390625e-08f);
}


// 4 standard normals for (global env id, draw number, block of 4 action components): Philox4x32-10 + Box-Muller.
// Shared by explore_kernel and the fused policy epilogue, so both draw the same noise.
__device__ __forceinline__ void explore_normals(const TaskParams& T, unsigned long long gid, unsigned int draw, uint32_t blk,
                                                float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), draw, blk, T.seed_lo ^ 0x4E4F4953u, T.seed_hi, r);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)(r[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-08f;     // (0, 1)
    const float u2 = (float)(r[2 * h + 1] >> 8) * 5.9604644775390625e-08f;           // [0, 1)
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z[2 * h] = rad * cs;
    z[2 * h + 1] = rad * sn;
  }
}


// ------------------------------------------------------------------------------------------------ fast scalar math
// Single-MUFU forms (<= 2 ulp) without the IEEE slow paths: sqrtf / fdiv expand into a fast path + a CALL to a
// denormal / rounding fix-up routine, which splits basic blocks (less ILP for ptxas) and bloats the unrolled body.
// The servo tolerates 1e-7 relative error everywhere these are used; the reset sampler, which must be bit-exact
// with the oracle, keeps the _rn intrinsics.
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// atan2(s, c) for s >= 0 -> [0, pi].  Branch-free: a = min/max in [0,1], odd minimax polynomial (degree 17, fitted
// with the a -> 0 slope pinned to 1; max abs error 7.4e-8 evaluated in fp32), then the two octant reflections.
__device__ __forceinline__ float atan2_nonneg(float s, float c) {
  const float ac = fabsf(c);
  const float mx = fmaxf(s, ac), mn = fminf(s, ac);
  const float a = mn * fast_rcp(fmaxf(mx, 1e-30f));
  const float z = a * a;
  float p = fmaf(z, 2.622258617e-03f, -1.513261348e-02f);
  p = fmaf(z, p, 4.112202674e-02f);
  p = fmaf(z, p, -7.366725057e-02f);
  p = fmaf(z, p, 1.057394370e-01f);
  p = fmaf(z, p, -1.418597847e-01f);
  p = fmaf(z, p, 1.999039650e-01f);
  p = fmaf(z, p, -3.333298564e-01f);
  float r = fmaf(p * z, a, a);
  r = (s > ac) ? 1.57079632679489662f - r : r;
  return (c < 0.0f) ? 3.14159265358979324f - r : r;
}

// ------------------------------------------------------------------------------------------------ kinematics
// sin/cos for joint angles: branch-free Cody-Waite reduction by pi/2 (3 constants, exact for |x| < ~1e4; joint angles
// are bounded by a few turns) + the Cephes single-precision minimax polynomials on [-pi/4, pi/4]; <= 1 ulp-ish
// (max abs error ~6e-8), ~22 instructions and no slow path, unlike sincosf whose Payne-Hanek branch bloats the
// unrolled FK code (I-cache) and adds divergence bookkeeping.
__device__ __forceinline__ void sincos_bounded(float x, float& s, float& c) {
  const float k = rintf(x * 0.63661977236758134f);
  float r = fmaf(k, -1.5707962512969971f, x);
  r = fmaf(k, -7.5497894158615964e-08f, r);
  r = fmaf(k, -5.3903029534742384e-15f, r);
  const int i = (int)k;
  const float z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(z, ps, -1.6666654611e-1f);
  ps = fmaf(ps * z, r, r);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(z, pc, 4.166664568298827e-2f);
  pc = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
  const float ss = (i & 1) ? pc : ps;
  const float cc = (i & 1) ? ps : pc;
  s = (i & 2) ? -ss : ss;
  c = ((i + 1) & 2) ? -cc : cc;
}

// FK of the 7-joint chain.  P[j] = origin of joint j (world), Z[j] = its axis (world); p, R = EE link frame.
template <bool WANT_JAC>
__device__ __forceinline__ void chain_fk(const ChainParams& C, const float (&q)[NJ], float (&p)[3], float (&R)[9],
                                         float (&P)[NJ][3], float (&Z)[NJ][3]) {
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = C.Rb[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = C.tb[i];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      p[i] = fmaf(R[3 * i + 2], C.t[j][2], fmaf(R[3 * i + 1], C.t[j][1], fmaf(R[3 * i], C.t[j][0], p[i])));
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        M[3 * i + k] = fmaf(R[3 * i + 2], C.Rf[j][6 + k], fmaf(R[3 * i + 1], C.Rf[j][3 + k], R[3 * i] * C.Rf[j][k]));
    if (WANT_JAC) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { P[j][i] = p[i]; Z[j][i] = M[3 * i + 2]; }
    }
    float s, c;
    sincos_bounded(q[j], s, c);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      R[3 * i] = fmaf(M[3 * i + 1], s, M[3 * i] * c);
      R[3 * i + 1] = fmaf(M[3 * i + 1], c, -M[3 * i] * s);
      R[3 * i + 2] = M[3 * i + 2];
    }
  }
}

// rotation matrix -> unit quaternion (xyzw), Shepperd's method (same branches as btMatrix3x3::getRotation)
__device__ __forceinline__ void mat_to_quat(const float (&m)[9], float (&q)[4]) {
  const float tr = m[0] + m[4] + m[8];
  if (tr > 0.0f) {
    float s = sqrtf(tr + 1.0f);
    q[3] = 0.5f * s;
    s = 0.5f / s;
    q[0] = (m[7] - m[5]) * s; q[1] = (m[2] - m[6]) * s; q[2] = (m[3] - m[1]) * s;
  } else if (m[0] >= m[4] && m[0] >= m[8]) {
    float s = sqrtf(m[0] - m[4] - m[8] + 1.0f);
    q[0] = 0.5f * s;
    s = 0.5f / s;
    q[3] = (m[7] - m[5]) * s; q[1] = (m[3] + m[1]) * s; q[2] = (m[6] + m[2]) * s;
  } else if (m[4] >= m[8]) {
    float s = sqrtf(m[4] - m[8] - m[0] + 1.0f);
    q[1] = 0.5f * s;
    s = 0.5f / s;
    q[3] = (m[2] - m[6]) * s; q[2] = (m[7] + m[5]) * s; q[0] = (m[1] + m[3]) * s;
  } else {
    float s = sqrtf(m[8] - m[0] - m[4] + 1.0f);
    q[2] = 0.5f * s;
    s = 0.5f / s;
    q[3] = (m[3] - m[1]) * s; q[0] = (m[2] + m[6]) * s; q[1] = (m[5] + m[7]) * s;
  }
}

// Orientation error as a rotation vector: angle * axis of (q_target (x) q_cur^-1), angle wrapped to (-pi, pi].
// Bullet (IKTrajectoryHelper::computeIK) forms it as 2*acos(w) * v/sqrt(1-w^2); for a unit quaternion that is
// 2*atan2(|v|, w) * v/|v|, which is the form used here because it stays well-conditioned in fp32 when the error is
// small (acos near 1 loses half the mantissa).
static __device__ __noinline__ void rot_error_quat(const float (&tq)[4], const float (&R)[9], float (&e)[3]) {
  float sq[4];
  mat_to_quat(R, sq);
  const float ix = -sq[0], iy = -sq[1], iz = -sq[2], iw = sq[3];
  const float dx = tq[3] * ix + tq[0] * iw + tq[1] * iz - tq[2] * iy;
  const float dy = tq[3] * iy + tq[1] * iw + tq[2] * ix - tq[0] * iz;
  const float dz = tq[3] * iz + tq[2] * iw + tq[0] * iy - tq[1] * ix;
  const float dw = tq[3] * iw - tq[0] * ix - tq[1] * iy - tq[2] * iz;
  const float vn = sqrtf(dx * dx + dy * dy + dz * dz);
  float k;
  if (vn < 1e-4f) {
    k = dw >= 0.0f ? 2.0f : -2.0f;  // angle ~ 2|v| (or the wrapped equivalent when w < 0)
  } else {
    float ang = 2.0f * atan2f(vn, dw);           // [0, 2pi]
    if (ang > 3.14159265358979f) ang -= 6.28318530717959f;
    k = ang / vn;
  }
  e[0] = k * dx; e[1] = k * dy; e[2] = k * dz;
}


// The same rotation vector from the error matrix E = R_target R^T without going through quaternions:
//   vee(E - E^T)/2 = sin(theta) * axis,  (tr E - 1)/2 = cos(theta),  e = axis * theta = v * theta / |v|.
// Straight-line (no Shepperd branches, which diverge lane by lane when the EE sits at the reference's target
// orientation, trace = -1 with two equal diagonal maxima) and well conditioned for theta in [0, ~170 deg]; beyond
// that sin(theta) -> 0 and the (cold) quaternion form above takes over.
__device__ __forceinline__ void rot_error(const float (&tR)[9], const float (&tq)[4], const float (&R)[9], float (&e)[3]) {
  // E[i][j] = row_i(tR) . row_j(R)
  const float e21 = fmaf(tR[8], R[5], fmaf(tR[7], R[4], tR[6] * R[3]));
  const float e12 = fmaf(tR[5], R[8], fmaf(tR[4], R[7], tR[3] * R[6]));
  const float e02 = fmaf(tR[2], R[8], fmaf(tR[1], R[7], tR[0] * R[6]));
  const float e20 = fmaf(tR[8], R[2], fmaf(tR[7], R[1], tR[6] * R[0]));
  const float e10 = fmaf(tR[5], R[2], fmaf(tR[4], R[1], tR[3] * R[0]));
  const float e01 = fmaf(tR[2], R[5], fmaf(tR[1], R[4], tR[0] * R[3]));
  const float e00 = fmaf(tR[2], R[2], fmaf(tR[1], R[1], tR[0] * R[0]));
  const float e11 = fmaf(tR[5], R[5], fmaf(tR[4], R[4], tR[3] * R[3]));
  const float e22 = fmaf(tR[8], R[8], fmaf(tR[7], R[7], tR[6] * R[6]));
  const float vx = 0.5f * (e21 - e12), vy = 0.5f * (e02 - e20), vz = 0.5f * (e10 - e01);
  const float c = 0.5f * (e00 + e11 + e22 - 1.0f);
  if (c < -0.98f) {            // within ~11 deg of a half turn: cold path
    rot_error_quat(tq, R, e);
    return;
  }
  const float s2 = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
  const float s = fast_sqrt(s2);
  // theta / sin(theta); -> 1 as theta -> 0 (s below 1e-4 rad: the series term s^2/6 is under fp32 resolution)
  const float k = s < 1e-4f ? 1.0f : atan2_nonneg(s, c) * fast_rcp(s);
  e[0] = k * vx; e[1] = k * vy; e[2] = k * vz;
}

// One damped-least-squares update  dq = J^T (J J^T + lambda I)^-1 e  -- algebraically identical to Bullet's
// (J^T J + lambda I)^-1 J^T e (Jacobian::CalcDeltaThetasDLS2) because the reference passes one damping value for all
// joints (rl_reach_env.py:111-113); the 6x6 form is used because J^T J + 1e-5 I has a null-space eigenvalue of 1e-5
// next to O(1) ones, which fp32 cannot resolve, while J J^T + 1e-5 I is well conditioned away from singularities.
__device__ __forceinline__ void dls_update(const float (&p)[3], const float (&P)[NJ][3], const float (&Z)[NJ][3],
                                           const float (&e)[6], float lambda, float (&dq)[NJ]) {
  float Jc[NJ][6];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float r0 = p[0] - P[j][0], r1 = p[1] - P[j][1], r2 = p[2] - P[j][2];
    Jc[j][0] = Z[j][1] * r2 - Z[j][2] * r1;
    Jc[j][1] = Z[j][2] * r0 - Z[j][0] * r2;
    Jc[j][2] = Z[j][0] * r1 - Z[j][1] * r0;
    Jc[j][3] = Z[j][0]; Jc[j][4] = Z[j][1]; Jc[j][5] = Z[j][2];
  }
  // A = J J^T + lambda I (lower triangle)
  float A[6][6];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      float acc = (a == b) ? lambda : 0.0f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc = fmaf(Jc[j][a], Jc[j][b], acc);
      A[a][b] = acc;
    }
  // Cholesky A = L L^T in place, keeping 1/L_ii
  float inv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int k = 0; k <= i; ++k) {
      float acc = A[i][k];
#pragma unroll
      for (int m = 0; m < k; ++m) acc = fmaf(-A[i][m], A[k][m], acc);
      if (k == i) {
        acc = fmaxf(acc, 1e-20f);
        inv[i] = fast_rsqrt(acc);
        A[i][i] = acc * inv[i];
      } else {
        A[i][k] = acc * inv[k];
      }
    }
  }
  float y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float acc = e[i];
#pragma unroll
    for (int m = 0; m < i; ++m) acc = fmaf(-A[i][m], y[m], acc);
    y[i] = acc * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    float acc = y[i];
#pragma unroll
    for (int m = i + 1; m < 6; ++m) acc = fmaf(-A[m][i], y[m], acc);
    y[i] = acc * inv[i];
  }
  float mx = 0.0f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float acc = 0.0f;
#pragma unroll
    for (int a = 0; a < 6; ++a) acc = fmaf(Jc[j][a], y[a], acc);
    dq[j] = acc;
    mx = fmaxf(mx, fabsf(acc));
  }
  const float max_angle = 0.78539816339744831f;  // BussIK MaxAngleDLS = 45 deg
  if (mx > max_angle) {
    const float sc = max_angle * fast_rcp(mx);
#pragma unroll
    for (int j = 0; j < NJ; ++j) dq[j] *= sc;
  }
}

#include "fk_generated.cuh"

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// The kinematic core of Env.step with ONE forward-kinematics site in the instruction stream (code size matters: the
// fully unrolled FK + DLS body must stay inside the 32 KB L1.5 instruction cache):
//
//   pass 0        : FK(q)  -> current EE (getLinkState, rl_reach_env.py:237); target = clip(ee + a*dv) (:239-242)
//   pass 1..iters : DLS update (calculateInverseKinematics, :244-250), FK(q), residual check
//   final pass    : only when the teleport does not take every joint (pick: joints 0..5, rl_pick_env.py:342) or the
//                   optional joint-limit clamp is on: FK of the state actually written back
//
// frozen = true (env finished, waiting for reset) runs pass 0 only.  Returns the DLS iteration count.
template <int ROBOT, bool CLIP>
__device__ __forceinline__ int servo_core(const ChainParams& C, const TaskParams& T, const float (&a)[3], bool frozen,
                                          float (&q)[NJ], float (&p)[3], float (&R)[9]) {
  float P[NJ][3], Z[NJ][3], tgt[3] = {0.f, 0.f, 0.f};
  const float q6_old = q[NJ - 1];
  const bool need_final = (T.napply < NJ) || T.clamp;
  int it = 0, phase = 0;  // 0 = first FK, 1 = iterating, 2 = final FK done
  for (;;) {
    RobotFK<ROBOT>::template run<true>(C, q, p, R, P, Z);
    if (phase == 2 || frozen) break;
    bool converged;
    if (phase == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        tgt[i] = fmaf(a[i], T.dv, p[i]);
        if (CLIP) tgt[i] = clampf(tgt[i], T.ws_lo[i], T.ws_hi[i]);
      }
      phase = 1;
      converged = false;   // Bullet starts from diff = +inf: at least one update whenever max_iters > 0
    } else {
      const float d0 = tgt[0] - p[0], d1 = tgt[1] - p[1], d2 = tgt[2] - p[2];
      converged = fast_sqrt(fmaf(d2, d2, fmaf(d1, d1, d0 * d0))) <= T.ik_residual;
    }
    if (converged || it >= T.ik_max_iters) {
      if (!need_final) break;
      if (T.napply < NJ) q[NJ - 1] = q6_old;
      if (T.clamp) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) q[j] = clampf(q[j], C.lower[j], C.upper[j]);
      }
      phase = 2;
      continue;
    }
    float e[6], er[3], dq[NJ];
    e[0] = tgt[0] - p[0]; e[1] = tgt[1] - p[1]; e[2] = tgt[2] - p[2];
    rot_error(T.tR, T.tquat, R, er);
    e[3] = er[0]; e[4] = er[1]; e[5] = er[2];
    dls_update(p, P, Z, e, T.ik_damping, dq);
#pragma unroll
    for (int j = 0; j < NJ; ++j) q[j] += dq[j];
    ++it;
  }
  return it;
}
