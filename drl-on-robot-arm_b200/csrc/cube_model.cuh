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

// ---- per-thread contact slots in shared memory ---------------------------------------------------------------------
// The contacts of the oracle's list live in a per-thread column of dynamic shared memory: word k of the calling thread
// is scratch[k * SLOT_STRIDE + threadIdx.x] (conflict-free: a warp reads 32 consecutive words).  Corner / table
// contacts are compacted per lane into slots 0..nc-1 (corner order = the oracle's Gauss-Seidel order), the up to 3
// arm-proxy contacts have static slots with an "active" bit.  A sweep is LDS/STS + a handful of FMAs per row; the
// cube's velocity state stays in registers.
// (Round 1 history: a compacted, dynamically indexed contact array lived in local memory, ~170 instructions per
// contact per sweep; static slots in registers made ptxas rematerialise the corner geometry inside the sweep loop,
// ~70-85; the shared-memory slots are ~50 and cut the kernel from 223 to ~120 registers.)
constexpr int SLOT_STRIDE = 128;                 // = LANE_BLOCK (threads per block of every kernel that steps cubes)
constexpr int CORNER_WORDS = 10;                 // r[3], bias, 1/k for n, t1, t2, lambda n, t1, t2
constexpr int PROXY_WORDS = 16;                  // r[3], n[3], t1[3], t2[3], bias, lambda n, t1, t2
constexpr int SCRATCH_WORDS = 8 * CORNER_WORDS + 3 * PROXY_WORDS;
constexpr int SCRATCH_BYTES = SCRATCH_WORDS * SLOT_STRIDE * 4;     // 64 KB per 128-thread block

extern __shared__ float cube_scratch[];

// One projected-Gauss-Seidel row for a general direction (arm-proxy contacts); v, w = cube twist (registers).
__device__ __forceinline__ void row(float (&v)[3], float (&w)[3], const float (&r)[3], const float (&dir)[3], float target,
                                    float lo, float hi, float& acc) {
  const float rxd[3] = {r[1] * dir[2] - r[2] * dir[1], r[2] * dir[0] - r[0] * dir[2], r[0] * dir[1] - r[1] * dir[0]};
  const float vrel = dir[0] * v[0] + dir[1] * v[1] + dir[2] * v[2] + rxd[0] * w[0] + rxd[1] * w[1] + rxd[2] * w[2];
  const float k = fmaf(rxd[0] * rxd[0] + rxd[1] * rxd[1] + rxd[2] * rxd[2], INV_INERTIA, INV_MASS);
  float dl = (target - vrel) * rcp_approx(k);
  const float nl = fminf(fmaxf(acc + dl, lo), hi);
  dl = nl - acc;
  acc = nl;
  const float dli = dl * INV_INERTIA;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = fmaf(dl, dir[i], v[i]);      // INV_MASS = 1
    w[i] = fmaf(dli, rxd[i], w[i]);
  }
}

// The three rows of a cube-corner / table contact.  The plane normal is +z, for which btPlaneSpace1 (tangents()) gives
// t1 = (0,-1,0), t2 = (1,0,0); r x dir is then a signed permutation of r and the generic row collapses to a handful
// of FMAs.  ik* = 1 / (1/m + |r x dir|^2 / I) per row, computed once per step.
__device__ __forceinline__ void table_rows(float (&v)[3], float (&w)[3], float r0, float r1, float r2, float bias, float ikn,
                                           float ik1, float ik2, float& ln, float& l1, float& l2) {
  const float r0i = r0 * INV_INERTIA, r1i = r1 * INV_INERTIA, r2i = r2 * INV_INERTIA;
  {  // normal (0,0,1): r x n = (r1, -r0, 0)
    const float vrel = fmaf(r1, w[0], fmaf(-r0, w[1], v[2]));
    const float nl = fminf(fmaxf(fmaf(bias - vrel, ikn, ln), 0.0f), 1e30f);
    const float dl = nl - ln;
    ln = nl;
    v[2] += dl;
    w[0] = fmaf(dl, r1i, w[0]);
    w[1] = fmaf(-dl, r0i, w[1]);
  }
  const float lim = MU * ln;
  {  // t1 = (0,-1,0): r x t1 = (r2, 0, -r0)
    const float vrel = fmaf(r2, w[0], fmaf(-r0, w[2], -v[1]));
    const float nl = fminf(fmaxf(fmaf(-vrel, ik1, l1), -lim), lim);
    const float dl = nl - l1;
    l1 = nl;
    v[1] -= dl;
    w[0] = fmaf(dl, r2i, w[0]);
    w[2] = fmaf(-dl, r0i, w[2]);
  }
  {  // t2 = (1,0,0): r x t2 = (0, r2, -r1)
    const float vrel = fmaf(r2, w[1], fmaf(-r1, w[2], v[0]));
    const float nl = fminf(fmaxf(fmaf(-vrel, ik2, l2), -lim), lim);
    const float dl = nl - l2;
    l2 = nl;
    v[0] += dl;
    w[1] = fmaf(dl, r2i, w[1]);
    w[2] = fmaf(-dl, r1i, w[2]);
  }
}

// one p.stepSimulation() for the cube; grip: 0 open / push, 1 closed, 2 holding.  Same contact list and the same
// Gauss-Seidel order as oracle/cube_model.h (corners 0..7 in index order, then the arm proxies).  Needs
// SCRATCH_BYTES of dynamic shared memory in the calling kernel.
template <bool PICK>
static __device__ __noinline__ void step(State& cbm, const float (&ee)[3], const float (&Ree)[9], float grip) {
  if (PICK && grip >= 1.5f) {
    cbm.pos[0] = ee[0] + GRIPPER_LEN * Ree[2];
    cbm.pos[1] = ee[1] + GRIPPER_LEN * Ree[5];
    cbm.pos[2] = ee[2] + GRIPPER_LEN * Ree[8];
#pragma unroll
    for (int i = 0; i < 3; ++i) { cbm.v[i] = 0.f; cbm.w[i] = 0.f; }
    return;
  }
  State cb = cbm;                    // work on a register copy (the caller's object lives behind a reference)
  float v[3], w[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { v[i] = cb.v[i] * DAMP; w[i] = cb.w[i] * DAMP; }
  v[2] = (cb.v[2] - G * DT) * DAMP;

  float* const sm = cube_scratch + threadIdx.x;
  float R[9];
  rot(cb.quat, R);
  unsigned active = 0u;               // bits 8.. : arm-proxy slots in contact
  int nc = 0;                         // corner contacts, compacted in corner order into slots 0 .. nc-1
  {
    float H[9];                      // half-edge vectors: column k of R times HALF
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = R[i] * HALF;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float r[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        r[i] = ((c & 1) ? H[3 * i] : -H[3 * i]) + ((c & 2) ? H[3 * i + 1] : -H[3 * i + 1]) + ((c & 4) ? H[3 * i + 2] : -H[3 * i + 2]);
      const float gap = cb.pos[2] + r[2] - TABLE_Z;
      if (gap < MARGIN) {
        float* s = sm + nc * (CORNER_WORDS * SLOT_STRIDE);
        ++nc;
        const float a = r[0] * r[0], b = r[1] * r[1], d = r[2] * r[2];
        s[0 * SLOT_STRIDE] = r[0]; s[1 * SLOT_STRIDE] = r[1]; s[2 * SLOT_STRIDE] = r[2];
        s[3 * SLOT_STRIDE] = gap < 0.f ? -ERP * gap * INV_DT : -gap * INV_DT;
        s[4 * SLOT_STRIDE] = rcp_approx(fmaf(a + b, INV_INERTIA, INV_MASS));
        s[5 * SLOT_STRIDE] = rcp_approx(fmaf(d + a, INV_INERTIA, INV_MASS));
        s[6 * SLOT_STRIDE] = rcp_approx(fmaf(d + b, INV_INERTIA, INV_MASS));
        s[7 * SLOT_STRIDE] = 0.f; s[8 * SLOT_STRIDE] = 0.f; s[9 * SLOT_STRIDE] = 0.f;
      }
    }
  }
  {
    constexpr int NP = PICK ? 3 : 1;
    float C[3][3], rad[3];
    const int np = arm_proxies(ee, Ree, PICK, grip, C, rad);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float rr[3], nn[3];
      const float d = sphere_query(cb, R, C[p], rad[p], rr, nn);
      if (p < np && d < 0.f) {
        active |= 1u << (8 + p);
        Contact k;
#pragma unroll
        for (int i = 0; i < 3; ++i) { k.r[i] = rr[i]; k.n[i] = nn[i]; }
        tangents(k);
        float* s = sm + (8 * CORNER_WORDS + p * PROXY_WORDS) * SLOT_STRIDE;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          s[(0 + i) * SLOT_STRIDE] = k.r[i]; s[(3 + i) * SLOT_STRIDE] = k.n[i];
          s[(6 + i) * SLOT_STRIDE] = k.t1[i]; s[(9 + i) * SLOT_STRIDE] = k.t2[i];
        }
        s[12 * SLOT_STRIDE] = -ERP * d * INV_DT;
        s[13 * SLOT_STRIDE] = 0.f; s[14 * SLOT_STRIDE] = 0.f; s[15 * SLOT_STRIDE] = 0.f;
      }
    }
  }

  if (active | (unsigned)nc) {
    constexpr int NP = PICK ? 3 : 1;
#pragma unroll 1
    for (int it = 0; it < PGS_ITERS; ++it) {
      // corner slots are compacted per lane, so a warp runs max(nc) bodies (4 for cubes lying flat) instead of one
      // body per corner any of its lanes touches the table with (up to 8 once some cubes have been tipped over)
#pragma unroll 1
      for (int j = 0; j < nc; ++j) {
        float* s = sm + j * (CORNER_WORDS * SLOT_STRIDE);
        float ln = s[7 * SLOT_STRIDE], l1 = s[8 * SLOT_STRIDE], l2 = s[9 * SLOT_STRIDE];
        table_rows(v, w, s[0 * SLOT_STRIDE], s[1 * SLOT_STRIDE], s[2 * SLOT_STRIDE], s[3 * SLOT_STRIDE], s[4 * SLOT_STRIDE],
                   s[5 * SLOT_STRIDE], s[6 * SLOT_STRIDE], ln, l1, l2);
        s[7 * SLOT_STRIDE] = ln; s[8 * SLOT_STRIDE] = l1; s[9 * SLOT_STRIDE] = l2;
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (active & (1u << (8 + p))) {
          float* s = sm + (8 * CORNER_WORDS + p * PROXY_WORDS) * SLOT_STRIDE;
          float r[3], n[3], t1[3], t2[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            r[i] = s[(0 + i) * SLOT_STRIDE]; n[i] = s[(3 + i) * SLOT_STRIDE];
            t1[i] = s[(6 + i) * SLOT_STRIDE]; t2[i] = s[(9 + i) * SLOT_STRIDE];
          }
          float ln = s[13 * SLOT_STRIDE], l1 = s[14 * SLOT_STRIDE], l2 = s[15 * SLOT_STRIDE];
          row(v, w, r, n, s[12 * SLOT_STRIDE], 0.0f, 1e30f, ln);
          const float lim = MU * ln;
          row(v, w, r, t1, 0.0f, -lim, lim, l1);
          row(v, w, r, t2, 0.0f, -lim, lim, l2);
          s[13 * SLOT_STRIDE] = ln; s[14 * SLOT_STRIDE] = l1; s[15 * SLOT_STRIDE] = l2;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) { cbm.v[i] = v[i]; cbm.w[i] = w[i]; cbm.pos[i] = fmaf(v[i], DT, cb.pos[i]); }
  // quaternion exponential map; half angle = |w| dt / 2 is far inside [-pi/4, pi/4] (|w| < 370 rad/s): polynomials only
  const float w2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const float h2 = w2 * (0.25f * DT * DT);                   // (half angle)^2
  // sin(h)/|w| = (dt/2) sin(h)/h,  sin(h)/h = 1 - h2/6 + h2^2/120 - h2^3/5040 ;  cos(h) = 1 - h2/2 + h2^2/24 - ...
  const float sinc = fmaf(h2, fmaf(h2, fmaf(h2, -1.9841270e-4f, 8.3333333e-3f), -1.6666667e-1f), 1.0f);
  const float co = fmaf(h2, fmaf(h2, fmaf(h2, fmaf(h2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.0f);
  const float sh = 0.5f * DT * sinc;
  const float dq[4] = {w[0] * sh, w[1] * sh, w[2] * sh, co};
  const float* q = cb.quat;
  float nq[4] = {dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1],
                 dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2],
                 dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0],
                 dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2]};
  const float inv = rsqrtf(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) cbm.quat[i] = nq[i] * inv;
}

}  // namespace cube
