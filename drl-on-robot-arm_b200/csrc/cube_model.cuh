// cube_model.cuh -- device (fp32) cube / pusher / gripper model of the push and pick envs.
//
// Stands in for what p.stepSimulation() does to the free cube in rl_push_env.py:349 / rl_pick_env.py:348,417 (Bullet
// rigid body + contact solver; not reproducible offline).  Model: free box (side 0.04, mass 1,
// box inertia, mu 2.5) on the table plane z = -0.025, pushed by penetration recovery against static sphere proxies
// of the teleported arm; 8 corner/plane contacts + sphere/box contacts, each 1 normal + 2 friction rows, 10 PGS
// sweeps, ERP 0.2, dt 1/240, gravity -10, damping 0.04; pick: latched finger closing within 6 mm and a kinematic hold.
// The fp64 statement of the same model used for parity is oracle/cube_model.h.
#pragma once
#include <cuda_runtime.h>

namespace cube {

constexpr float DT = 1.0f / 240.0f;
constexpr float INV_DT = 240.0f;
constexpr float G = 10.0f;
constexpr float HALF = 0.02f;
constexpr float INV_MASS = 1.0f;
constexpr float INV_INERTIA = 1.0f / (1.0f * (0.04f * 0.04f) / 6.0f);
constexpr float MU = 2.5f;
constexpr float ERP = 0.2f;
constexpr float TABLE_Z = -0.025f;
constexpr float MARGIN = 0.005f;
constexpr int PGS_ITERS = 10;
constexpr float DAMP = 0.99982992284f;
constexpr int MAX_CONTACTS = 11;
constexpr float PUSH_R = 0.045f, PUSH_OFF = 0.02f;
constexpr float PALM_R = 0.05f, PALM_OFF = 0.12f, TIP_R = 0.012f, TIP_OPEN = 0.045f, TIP_CLOSED_R = 0.02f;
constexpr float GRIPPER_LEN = 0.257f, CLOSE_DIST = 0.006f, HOLD_DIST = 0.03f;

struct State {
  float pos[3], quat[4], v[3], w[3];
};

__device__ __forceinline__ void init(State& c, float x, float y, float z, float yaw) {
  c.pos[0] = x; c.pos[1] = y; c.pos[2] = z;
  float s, co;
  sincosf(0.5f * yaw, &s, &co);
  c.quat[0] = 0.f; c.quat[1] = 0.f; c.quat[2] = s; c.quat[3] = co;
#pragma unroll
  for (int i = 0; i < 3; ++i) { c.v[i] = 0.f; c.w[i] = 0.f; }
}

__device__ __forceinline__ void rot(const float (&q)[4], float (&R)[9]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

struct Contact {
  float r[3], n[3], t1[3], t2[3];
  float bias, ln, l1, l2;
};

__device__ __forceinline__ void tangents(Contact& k) {  // btPlaneSpace1
  const float* n = k.n;
  if (fabsf(n[2]) > 0.70710678f) {
    const float a = n[1] * n[1] + n[2] * n[2], s = rsqrtf(a);
    k.t1[0] = 0.f; k.t1[1] = -n[2] * s; k.t1[2] = n[1] * s;
    k.t2[0] = a * s; k.t2[1] = -n[0] * k.t1[2]; k.t2[2] = n[0] * k.t1[1];
  } else {
    const float a = n[0] * n[0] + n[1] * n[1], s = rsqrtf(a);
    k.t1[0] = -n[1] * s; k.t1[1] = n[0] * s; k.t1[2] = 0.f;
    k.t2[0] = -n[2] * k.t1[1]; k.t2[1] = n[2] * k.t1[0]; k.t2[2] = a * s;
  }
}

// signed distance sphere <-> box, closest point on the box (relative to the cube centre, world axes) and the unit
// direction from the sphere towards the cube
__device__ __forceinline__ float sphere_query(const State& cb, const float (&R)[9], const float (&c)[3], float rad,
                                              float (&rrel)[3], float (&n)[3]) {
  const float d[3] = {c[0] - cb.pos[0], c[1] - cb.pos[1], c[2] - cb.pos[2]};
  float l[3], cl[3], nl[3];
  bool inside = true;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    l[i] = R[i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2];
    cl[i] = fminf(fmaxf(l[i], -HALF), HALF);
    inside = inside && (cl[i] == l[i]);
  }
  float dist;
  if (!inside) {
    const float e0 = l[0] - cl[0], e1 = l[1] - cl[1], e2 = l[2] - cl[2];
    dist = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    const float inv = 1.0f / dist;
    nl[0] = -e0 * inv; nl[1] = -e1 * inv; nl[2] = -e2 * inv;
  } else {
    int ax = 0;
    float best = HALF - fabsf(l[0]);
#pragma unroll
    for (int i = 1; i < 3; ++i) {
      const float m = HALF - fabsf(l[i]);
      if (m < best) { best = m; ax = i; }
    }
    dist = -best;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float sgn = l[i] >= 0.f ? 1.0f : -1.0f;
      nl[i] = (i == ax) ? -sgn : 0.f;
      if (i == ax) cl[i] = sgn * HALF;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    rrel[i] = R[3 * i] * cl[0] + R[3 * i + 1] * cl[1] + R[3 * i + 2] * cl[2];
    n[i] = R[3 * i] * nl[0] + R[3 * i + 1] * nl[1] + R[3 * i + 2] * nl[2];
  }
  return dist - rad;
}

__device__ __forceinline__ int arm_proxies(const float (&ee)[3], const float (&Ree)[9], bool pick, float grip,
                                           float (&C)[3][3], float (&rad)[3]) {
  const float zx = Ree[2], zy = Ree[5], zz = Ree[8];
  const float xx = Ree[0], xy = Ree[3], xz = Ree[6];
  if (!pick) {
    C[0][0] = ee[0] + PUSH_OFF * zx; C[0][1] = ee[1] + PUSH_OFF * zy; C[0][2] = ee[2] + PUSH_OFF * zz;
    rad[0] = PUSH_R;
    return 1;
  }
  C[0][0] = ee[0] + PALM_OFF * zx; C[0][1] = ee[1] + PALM_OFF * zy; C[0][2] = ee[2] + PALM_OFF * zz;
  rad[0] = PALM_R;
  const float g[3] = {ee[0] + GRIPPER_LEN * zx, ee[1] + GRIPPER_LEN * zy, ee[2] + GRIPPER_LEN * zz};
  if (grip >= 0.5f) {
    C[1][0] = g[0]; C[1][1] = g[1]; C[1][2] = g[2];
    rad[1] = TIP_CLOSED_R;
    return 2;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float o = s ? -TIP_OPEN : TIP_OPEN;
    C[1 + s][0] = g[0] + o * xx; C[1 + s][1] = g[1] + o * xy; C[1 + s][2] = g[2] + o * xz;
    rad[1 + s] = TIP_R;
  }
  return 3;
}

__device__ __forceinline__ float gripper_distance(const State& cb, const float (&ee)[3], const float (&Ree)[9]) {
  float R[9], C[3][3], rad[3], rr[3], nn[3];
  rot(cb.quat, R);
  const int np = arm_proxies(ee, Ree, true, 0.0f, C, rad);
  float best = 1e30f;
  for (int i = 0; i < np; ++i) best = fminf(best, sphere_query(cb, R, C[i], rad[i], rr, nn));
  return best;
}

__device__ __forceinline__ void row(State& cb, const float (&r)[3], const float (&dir)[3], float target, float lo, float hi,
                                    float& acc) {
  const float rxd[3] = {r[1] * dir[2] - r[2] * dir[1], r[2] * dir[0] - r[0] * dir[2], r[0] * dir[1] - r[1] * dir[0]};
  const float vrel = dir[0] * cb.v[0] + dir[1] * cb.v[1] + dir[2] * cb.v[2] + rxd[0] * cb.w[0] + rxd[1] * cb.w[1] + rxd[2] * cb.w[2];
  const float k = INV_MASS + (rxd[0] * rxd[0] + rxd[1] * rxd[1] + rxd[2] * rxd[2]) * INV_INERTIA;
  float dl = (target - vrel) / k;
  float nl = fminf(fmaxf(acc + dl, lo), hi);
  dl = nl - acc;
  acc = nl;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    cb.v[i] += dl * dir[i] * INV_MASS;
    cb.w[i] += dl * rxd[i] * INV_INERTIA;
  }
}

// one p.stepSimulation() for the cube; grip: 0 open / push, 1 closed, 2 holding
static __device__ __noinline__ void step(State& cb, const float (&ee)[3], const float (&Ree)[9], bool pick, float grip) {
  if (pick && grip >= 1.5f) {
    cb.pos[0] = ee[0] + GRIPPER_LEN * Ree[2];
    cb.pos[1] = ee[1] + GRIPPER_LEN * Ree[5];
    cb.pos[2] = ee[2] + GRIPPER_LEN * Ree[8];
#pragma unroll
    for (int i = 0; i < 3; ++i) { cb.v[i] = 0.f; cb.w[i] = 0.f; }
    return;
  }
  cb.v[2] -= G * DT;
#pragma unroll
  for (int i = 0; i < 3; ++i) { cb.v[i] *= DAMP; cb.w[i] *= DAMP; }

  float R[9];
  rot(cb.quat, R);
  Contact K[MAX_CONTACTS];
  int nk = 0;
  for (int c = 0; c < 8; ++c) {
    const float l0 = (c & 1) ? HALF : -HALF, l1 = (c & 2) ? HALF : -HALF, l2 = (c & 4) ? HALF : -HALF;
    float r[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = R[3 * i] * l0 + R[3 * i + 1] * l1 + R[3 * i + 2] * l2;
    const float gap = cb.pos[2] + r[2] - TABLE_Z;
    if (gap < MARGIN) {
      Contact& k = K[nk++];
      k.r[0] = r[0]; k.r[1] = r[1]; k.r[2] = r[2];
      k.n[0] = 0.f; k.n[1] = 0.f; k.n[2] = 1.f;
      k.bias = gap < 0.f ? -ERP * gap * INV_DT : -gap * INV_DT;
      k.ln = k.l1 = k.l2 = 0.f;
      tangents(k);
    }
  }
  float C[3][3], rad[3];
  const int np = arm_proxies(ee, Ree, pick, grip, C, rad);
  for (int p = 0; p < np; ++p) {
    float rr[3], nn[3];
    const float d = sphere_query(cb, R, C[p], rad[p], rr, nn);
    if (d < 0.f) {
      Contact& k = K[nk++];
#pragma unroll
      for (int i = 0; i < 3; ++i) { k.r[i] = rr[i]; k.n[i] = nn[i]; }
      k.bias = -ERP * d * INV_DT;
      k.ln = k.l1 = k.l2 = 0.f;
      tangents(k);
    }
  }
  for (int it = 0; it < PGS_ITERS; ++it) {
    for (int i = 0; i < nk; ++i) {
      Contact& k = K[i];
      row(cb, k.r, k.n, k.bias, 0.0f, 1e30f, k.ln);
      const float lim = MU * k.ln;
      row(cb, k.r, k.t1, 0.0f, -lim, lim, k.l1);
      row(cb, k.r, k.t2, 0.0f, -lim, lim, k.l2);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) cb.pos[i] += cb.v[i] * DT;
  const float wn = sqrtf(cb.w[0] * cb.w[0] + cb.w[1] * cb.w[1] + cb.w[2] * cb.w[2]);
  const float ang = wn * DT;
  float s, co;
  sincosf(0.5f * ang, &s, &co);
  s = ang > 1e-6f ? s / wn : 0.5f * DT * (1.0f - ang * ang * (1.0f / 24.0f));
  const float dq[4] = {cb.w[0] * s, cb.w[1] * s, cb.w[2] * s, co};
  const float* q = cb.quat;
  float nq[4] = {dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1],
                 dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2],
                 dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0],
                 dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2]};
  const float inv = rsqrtf(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) cb.quat[i] = nq[i] * inv;
}

}  // namespace cube
