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

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

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

// One projected-Gauss-Seidel row for a general direction (arm-proxy contacts).  rxd = r x dir and 1/k do not change
// between sweeps but are recomputed (9 instructions) instead of held: registers are the scarcer resource here.
__device__ __forceinline__ void row(State& cb, const float (&r)[3], const float (&dir)[3], float target, float lo, float hi,
                                    float& acc) {
  const float rxd[3] = {r[1] * dir[2] - r[2] * dir[1], r[2] * dir[0] - r[0] * dir[2], r[0] * dir[1] - r[1] * dir[0]};
  const float vrel = dir[0] * cb.v[0] + dir[1] * cb.v[1] + dir[2] * cb.v[2] + rxd[0] * cb.w[0] + rxd[1] * cb.w[1] + rxd[2] * cb.w[2];
  const float k = fmaf(rxd[0] * rxd[0] + rxd[1] * rxd[1] + rxd[2] * rxd[2], INV_INERTIA, INV_MASS);
  float dl = (target - vrel) * rcp_approx(k);
  const float nl = fminf(fmaxf(acc + dl, lo), hi);
  dl = nl - acc;
  acc = nl;
  const float dli = dl * INV_INERTIA;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    cb.v[i] = fmaf(dl, dir[i], cb.v[i]);      // INV_MASS = 1
    cb.w[i] = fmaf(dli, rxd[i], cb.w[i]);
  }
}

// The three rows of a cube-corner / table contact.  The plane normal is +z, for which btPlaneSpace1 (tangents()) gives
// t1 = (0,-1,0), t2 = (1,0,0); r x dir is then a signed permutation of r and the generic row collapses to a handful
// of FMAs.  ik* = 1 / (1/m + |r x dir|^2 / I) per row, computed once per step.
__device__ __forceinline__ void table_rows(State& cb, float r0, float r1, float r2, float bias, float ikn, float ik1, float ik2,
                                           float& ln, float& l1, float& l2) {
  const float r0i = r0 * INV_INERTIA, r1i = r1 * INV_INERTIA, r2i = r2 * INV_INERTIA;
  {  // normal (0,0,1): r x n = (r1, -r0, 0)
    const float vrel = fmaf(r1, cb.w[0], fmaf(-r0, cb.w[1], cb.v[2]));
    const float nl = fminf(fmaxf(fmaf(bias - vrel, ikn, ln), 0.0f), 1e30f);
    const float dl = nl - ln;
    ln = nl;
    cb.v[2] += dl;
    cb.w[0] = fmaf(dl, r1i, cb.w[0]);
    cb.w[1] = fmaf(-dl, r0i, cb.w[1]);
  }
  const float lim = MU * ln;
  {  // t1 = (0,-1,0): r x t1 = (r2, 0, -r0)
    const float vrel = fmaf(r2, cb.w[0], fmaf(-r0, cb.w[2], -cb.v[1]));
    const float nl = fminf(fmaxf(fmaf(-vrel, ik1, l1), -lim), lim);
    const float dl = nl - l1;
    l1 = nl;
    cb.v[1] -= dl;
    cb.w[0] = fmaf(dl, r2i, cb.w[0]);
    cb.w[2] = fmaf(-dl, r0i, cb.w[2]);
  }
  {  // t2 = (1,0,0): r x t2 = (0, r2, -r1)
    const float vrel = fmaf(r2, cb.w[1], fmaf(-r1, cb.w[2], cb.v[0]));
    const float nl = fminf(fmaxf(fmaf(-vrel, ik2, l2), -lim), lim);
    const float dl = nl - l2;
    l2 = nl;
    cb.v[0] += dl;
    cb.w[1] = fmaf(dl, r2i, cb.w[1]);
    cb.w[2] = fmaf(-dl, r1i, cb.w[2]);
  }
}

// one p.stepSimulation() for the cube; grip: 0 open / push, 1 closed, 2 holding.
//
// Same contact list and the same Gauss-Seidel order as oracle/cube_model.h (corners 0..7 in index order, then the arm
// proxies), but every contact has a STATIC slot -- 8 corner slots + NP proxy slots with an "active" bit -- instead of
// a compacted array: the slots are indexed by unrolled compile-time constants, so the whole contact set lives in
// registers (the compacted, dynamically indexed array of round 1's first version lived in local memory: 1 KB stack
// frame, ~170 instructions per contact per sweep; now ~45 for a corner).  The corner lever arm r is rebuilt from the
// three half-edge vectors with sign flips (6 FADD) rather than held.
template <bool PICK>
static __device__ __noinline__ void step(State& cb, const float (&ee)[3], const float (&Ree)[9], float grip) {
  if (PICK && grip >= 1.5f) {
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
  float H[9];                        // half-edge vectors: column k of R times HALF
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] = R[i] * HALF;

  // ---- corner slots
  unsigned active = 0u;
  float tb[8], tln[8], tl1[8], tl2[8], tkn[8], tk1[8], tk2[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float r[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      r[i] = ((c & 1) ? H[3 * i] : -H[3 * i]) + ((c & 2) ? H[3 * i + 1] : -H[3 * i + 1]) + ((c & 4) ? H[3 * i + 2] : -H[3 * i + 2]);
    const float gap = cb.pos[2] + r[2] - TABLE_Z;
    if (gap < MARGIN) active |= 1u << c;
    tb[c] = gap < 0.f ? -ERP * gap * INV_DT : -gap * INV_DT;
    tln[c] = tl1[c] = tl2[c] = 0.f;
    const float a = r[0] * r[0], b = r[1] * r[1], d = r[2] * r[2];
    tkn[c] = rcp_approx(fmaf(a + b, INV_INERTIA, INV_MASS));
    tk1[c] = rcp_approx(fmaf(d + a, INV_INERTIA, INV_MASS));
    tk2[c] = rcp_approx(fmaf(d + b, INV_INERTIA, INV_MASS));
  }

  // ---- arm-proxy slots
  constexpr int NP = PICK ? 3 : 1;
  float C[3][3], rad[3];
  const int np = arm_proxies(ee, Ree, PICK, grip, C, rad);
  Contact K[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float rr[3], nn[3];
    const float d = sphere_query(cb, R, C[p], rad[p], rr, nn);
    if (p < np && d < 0.f) {
      active |= 1u << (8 + p);
#pragma unroll
      for (int i = 0; i < 3; ++i) { K[p].r[i] = rr[i]; K[p].n[i] = nn[i]; }
      K[p].bias = -ERP * d * INV_DT;
      tangents(K[p]);
    }
    K[p].ln = K[p].l1 = K[p].l2 = 0.f;
  }

  if (active) {
    for (int it = 0; it < PGS_ITERS; ++it) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (active & (1u << c)) {
          const float r0 = ((c & 1) ? H[0] : -H[0]) + ((c & 2) ? H[1] : -H[1]) + ((c & 4) ? H[2] : -H[2]);
          const float r1 = ((c & 1) ? H[3] : -H[3]) + ((c & 2) ? H[4] : -H[4]) + ((c & 4) ? H[5] : -H[5]);
          const float r2 = ((c & 1) ? H[6] : -H[6]) + ((c & 2) ? H[7] : -H[7]) + ((c & 4) ? H[8] : -H[8]);
          table_rows(cb, r0, r1, r2, tb[c], tkn[c], tk1[c], tk2[c], tln[c], tl1[c], tl2[c]);
        }
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (active & (1u << (8 + p))) {
          Contact& k = K[p];
          row(cb, k.r, k.n, k.bias, 0.0f, 1e30f, k.ln);
          const float lim = MU * k.ln;
          row(cb, k.r, k.t1, 0.0f, -lim, lim, k.l1);
          row(cb, k.r, k.t2, 0.0f, -lim, lim, k.l2);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) cb.pos[i] = fmaf(cb.v[i], DT, cb.pos[i]);
  // quaternion exponential map; half angle = |w| dt / 2 is far inside [-pi/4, pi/4] (|w| < 370 rad/s): polynomials only
  const float w2 = cb.w[0] * cb.w[0] + cb.w[1] * cb.w[1] + cb.w[2] * cb.w[2];
  const float h2 = w2 * (0.25f * DT * DT);                   // (half angle)^2
  // sin(h)/|w| = (dt/2) sin(h)/h,  sin(h)/h = 1 - h2/6 + h2^2/120 - h2^3/5040 ;  cos(h) = 1 - h2/2 + h2^2/24 - ...
  const float sinc = fmaf(h2, fmaf(h2, fmaf(h2, -1.9841270e-4f, 8.3333333e-3f), -1.6666667e-1f), 1.0f);
  const float co = fmaf(h2, fmaf(h2, fmaf(h2, fmaf(h2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.0f);
  const float s = 0.5f * DT * sinc;
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
