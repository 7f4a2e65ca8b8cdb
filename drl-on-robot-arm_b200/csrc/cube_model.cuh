// cube_model.cuh -- device (fp32) cube / pusher / gripper model of the push and pick envs.
//
// Stands in for what p.stepSimulation() does to the free cube in rl_push_env.py:349 / rl_pick_env.py:348,417 (Bullet
// multibody world + contact solver; not reproducible offline).  The model is stated in
// oracle/cube_model.h (fp64, the parity reference of this file): free box (side 0.04, mass 1, box inertia, mu 2.5) on
// the table top z = -0.025 (ground z = -0.65 beyond the table edge), moved by penetration recovery against static
// CAPSULES fixed in the EE frame (push: flange / link 6; pick: palm + two fingers); up to 4 corner/plane contacts +
// the capsule contacts, each 1 normal + 2 friction rows; Bullet's solver parameters: ERP 0.2 as a velocity bias,
// dt 1/240, gravity -10, damping 0.04, at most 50 projected-Gauss-Seidel sweeps with pybullet's early exit (largest
// squared velocity residual of a sweep <= 1e-7); pick: fingers close for good within 6 mm and hold by friction only.
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
constexpr float TABLE_Z = -0.025f, GROUND_Z = -0.65f;
constexpr float TABLE_X0 = -0.25f, TABLE_X1 = 1.25f, TABLE_Y0 = -0.5f, TABLE_Y1 = 0.5f;
constexpr float MARGIN = 0.005f;
constexpr int PGS_ITERS = 50;                 // pybullet numSolverIterations
constexpr float PGS_RESIDUAL = 1e-7f;         // pybullet m_leastSquaresResidualThreshold
constexpr float DAMP = 0.99982992284f;
constexpr int MAX_CORNERS = 4, MAX_PROXIES = 3;
constexpr float PUSH_R = 0.04f, PUSH_A0 = -0.10f, PUSH_A1 = 0.005f;
constexpr float PALM_R = 0.045f, PALM_A0 = 0.0f, PALM_A1 = 0.15f;
constexpr float FINGER_R = 0.01f, FINGER_A0 = 0.15f, FINGER_A1 = 0.247f, FINGER_BASE = 0.03f;
constexpr float TIP_OPEN = 0.05f, TIP_CLOSED = 0.025f;
constexpr float GRIPPER_LEN = 0.257f, CLOSE_DIST = 0.006f;

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

__device__ __forceinline__ void tangents(const float (&n)[3], float (&t1)[3], float (&t2)[3]) {  // btPlaneSpace1
  if (fabsf(n[2]) > 0.70710678f) {
    const float a = n[1] * n[1] + n[2] * n[2], s = rsqrtf(a);
    t1[0] = 0.f; t1[1] = -n[2] * s; t1[2] = n[1] * s;
    t2[0] = a * s; t2[1] = -n[0] * t1[2]; t2[2] = n[0] * t1[1];
  } else {
    const float a = n[0] * n[0] + n[1] * n[1], s = rsqrtf(a);
    t1[0] = -n[1] * s; t1[1] = n[0] * s; t1[2] = 0.f;
    t2[0] = -n[2] * t1[1]; t2[1] = n[2] * t1[0]; t2[2] = a * s;
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

// Capsules of the arm in the EE frame: axis point = ee + along * z_ee + side * x_ee.  push: 1 (flange / link 6);
// pick: palm + two fingers whose tips sit TIP_OPEN / TIP_CLOSED off the axis.
struct Capsule { float a[3], b[3], rad; };

__device__ __forceinline__ void axis_point(const float (&ee)[3], const float (&Ree)[9], float along, float side, float (&o)[3]) {
  o[0] = fmaf(side, Ree[0], fmaf(along, Ree[2], ee[0]));
  o[1] = fmaf(side, Ree[3], fmaf(along, Ree[5], ee[1]));
  o[2] = fmaf(side, Ree[6], fmaf(along, Ree[8], ee[2]));
}

template <bool PICK>
__device__ __forceinline__ void arm_capsule(const float (&ee)[3], const float (&Ree)[9], float grip, int p, Capsule& k) {
  if (!PICK) {
    axis_point(ee, Ree, PUSH_A0, 0.f, k.a); axis_point(ee, Ree, PUSH_A1, 0.f, k.b); k.rad = PUSH_R;
  } else if (p == 0) {
    axis_point(ee, Ree, PALM_A0, 0.f, k.a); axis_point(ee, Ree, PALM_A1, 0.f, k.b); k.rad = PALM_R;
  } else {
    const float sg = p == 1 ? 1.0f : -1.0f;
    const float tip = grip >= 0.5f ? TIP_CLOSED : TIP_OPEN;
    axis_point(ee, Ree, FINGER_A0, sg * FINGER_BASE, k.a); axis_point(ee, Ree, FINGER_A1, sg * tip, k.b); k.rad = FINGER_R;
  }
}

// capsule <-> box: the sphere of the capsule's radius at the axis point nearest the cube centre
__device__ __forceinline__ float capsule_query(const State& cb, const float (&R)[9], const Capsule& k, float (&rrel)[3], float (&n)[3]) {
  const float ab[3] = {k.b[0] - k.a[0], k.b[1] - k.a[1], k.b[2] - k.a[2]};
  const float ac[3] = {cb.pos[0] - k.a[0], cb.pos[1] - k.a[1], cb.pos[2] - k.a[2]};
  const float len2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  float t = (ab[0] * ac[0] + ab[1] * ac[1] + ab[2] * ac[2]) / len2;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float c[3] = {fmaf(t, ab[0], k.a[0]), fmaf(t, ab[1], k.a[1]), fmaf(t, ab[2], k.a[2])};
  return sphere_query(cb, R, c, k.rad, rrel, n);
}

// getClosestPoints(kuka, cube, 0.006) stand-in (rl_pick_env.py:412): min signed distance capsule <-> cube
__device__ __forceinline__ float gripper_distance(const State& cb, const float (&ee)[3], const float (&Ree)[9], float grip) {
  float R[9], rr[3], nn[3];
  rot(cb.quat, R);
  float best = 1e30f;
#pragma unroll
  for (int p = 0; p < MAX_PROXIES; ++p) {
    Capsule k;
    arm_capsule<true>(ee, Ree, grip, p, k);
    best = fminf(best, capsule_query(cb, R, k, rr, nn));
  }
  return best;
}

// ---- per-thread contact slots in shared memory ---------------------------------------------------------------------
// The contacts of the oracle's list live in a per-thread column of dynamic shared memory: word k of the calling thread
// is scratch[k * SLOT_STRIDE + threadIdx.x] (conflict-free: a warp reads 32 consecutive words).  Corner / plane
// contacts are compacted per lane into slots 0..nc-1 (corner order = the oracle's Gauss-Seidel order, at most a face),
// the capsule contacts have static slots with an "active" bit.  Everything that does not change during the sweeps --
// lever arms, r x dir, the row's effective mass k and 1/k -- is computed once per step, so a row inside the sweep loop
// is LDS + one 6-term dot + clamp + two 3-term AXPYs; the cube's twist stays in registers.
// (Round 1 history: a compacted, dynamically indexed contact array lived in local memory, ~170 instructions per
// contact per sweep; static slots in registers made ptxas rematerialise the corner geometry inside the sweep loop,
// ~70-85; the shared-memory slots are ~50.)
constexpr int SLOT_STRIDE = 128;                 // = LANE_BLOCK (threads per block of every kernel that steps cubes)
constexpr int CORNER_WORDS = 13;                 // r[3], bias, k n/t1/t2, 1/k n/t1/t2, lambda n/t1/t2
constexpr int ROW_WORDS = 9;                     // dir[3], r x dir [3], k, 1/k, lambda
constexpr int PROXY_WORDS = 3 * ROW_WORDS + 1;   // rows n, t1, t2 + bias
template <bool PICK> constexpr int scratch_words() { return MAX_CORNERS * CORNER_WORDS + (PICK ? 3 : 1) * PROXY_WORDS; }
template <bool PICK> constexpr int scratch_bytes() { return scratch_words<PICK>() * SLOT_STRIDE * 4; }   // push 40 KB, pick 68 KB

extern __shared__ float cube_scratch[];

// One projected-Gauss-Seidel row from its slot; v, w = cube twist (registers); res = running max of the squared
// velocity change along a row (Bullet: deltaImpulse / jacDiagABInv, squared).
__device__ __forceinline__ float slot_row(float* s, float (&v)[3], float (&w)[3], float target, float lo, float hi, float& res) {
  const float d0 = s[0 * SLOT_STRIDE], d1 = s[1 * SLOT_STRIDE], d2 = s[2 * SLOT_STRIDE];
  const float x0 = s[3 * SLOT_STRIDE], x1 = s[4 * SLOT_STRIDE], x2 = s[5 * SLOT_STRIDE];
  const float k = s[6 * SLOT_STRIDE], ik = s[7 * SLOT_STRIDE], acc = s[8 * SLOT_STRIDE];
  const float vrel = fmaf(d0, v[0], fmaf(d1, v[1], d2 * v[2])) + fmaf(x0, w[0], fmaf(x1, w[1], x2 * w[2]));
  const float nl = fminf(fmaxf(fmaf(target - vrel, ik, acc), lo), hi);
  const float dl = nl - acc;
  s[8 * SLOT_STRIDE] = nl;
  const float dli = dl * INV_INERTIA;
  v[0] = fmaf(dl, d0, v[0]); v[1] = fmaf(dl, d1, v[1]); v[2] = fmaf(dl, d2, v[2]);      // INV_MASS = 1
  w[0] = fmaf(dli, x0, w[0]); w[1] = fmaf(dli, x1, w[1]); w[2] = fmaf(dli, x2, w[2]);
  const float dv = dl * k;
  res = fmaxf(res, dv * dv);
  return nl;
}

// The three rows of a cube-corner / plane contact.  The plane normal is +z, for which btPlaneSpace1 (tangents()) gives
// t1 = (0,-1,0), t2 = (1,0,0); r x dir is then a signed permutation of r and the generic row collapses to a handful
// of FMAs.  k* = 1/m + |r x dir|^2 / I per row and ik* = 1 / k*, computed once per step.
__device__ __forceinline__ void table_rows(float (&v)[3], float (&w)[3], float r0, float r1, float r2, float bias, float kn,
                                           float k1, float k2, float ikn, float ik1, float ik2, float& ln, float& l1, float& l2,
                                           float& res) {
  const float r0i = r0 * INV_INERTIA, r1i = r1 * INV_INERTIA, r2i = r2 * INV_INERTIA;
  {  // normal (0,0,1): r x n = (r1, -r0, 0)
    const float vrel = fmaf(r1, w[0], fmaf(-r0, w[1], v[2]));
    const float nl = fmaxf(fmaf(bias - vrel, ikn, ln), 0.0f);
    const float dl = nl - ln;
    ln = nl;
    v[2] += dl;
    w[0] = fmaf(dl, r1i, w[0]);
    w[1] = fmaf(-dl, r0i, w[1]);
    const float dv = dl * kn;
    res = fmaxf(res, dv * dv);
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
    const float dv = dl * k1;
    res = fmaxf(res, dv * dv);
  }
  {  // t2 = (1,0,0): r x t2 = (0, r2, -r1)
    const float vrel = fmaf(r2, w[1], fmaf(-r1, w[2], v[0]));
    const float nl = fminf(fmaxf(fmaf(-vrel, ik2, l2), -lim), lim);
    const float dl = nl - l2;
    l2 = nl;
    v[0] += dl;
    w[1] = fmaf(dl, r2i, w[1]);
    w[2] = fmaf(-dl, r1i, w[2]);
    const float dv = dl * k2;
    res = fmaxf(res, dv * dv);
  }
}

// one p.stepSimulation() for the cube; grip: 0 open / push, >= 0.5 fingers closed.  Same contact list and the same
// Gauss-Seidel order as oracle/cube_model.h (corner contacts in corner order, then the arm capsules).  Needs
// scratch_bytes<PICK>() of dynamic shared memory in the calling kernel.
template <bool PICK>
static __device__ __noinline__ void step(State& cbm, const float (&ee)[3], const float (&Ree)[9], float grip) {
  State cb = cbm;                    // work on a register copy (the caller's object lives behind a reference)
  float v[3], w[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { v[i] = cb.v[i] * DAMP; w[i] = cb.w[i] * DAMP; }
  v[2] = (cb.v[2] - G * DT) * DAMP;

  float* const sm = cube_scratch + threadIdx.x;
  float R[9];
  rot(cb.quat, R);
  unsigned active = 0u;               // bit p : arm capsule p in contact
  int nc = 0;                         // corner contacts, compacted in corner order into slots 0 .. nc-1
  {
    const bool on_table = cb.pos[0] >= TABLE_X0 && cb.pos[0] <= TABLE_X1 && cb.pos[1] >= TABLE_Y0 && cb.pos[1] <= TABLE_Y1;
    const float plane_z = on_table ? TABLE_Z : GROUND_Z;
    float H[9];                      // half-edge vectors: column k of R times HALF
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = R[i] * HALF;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float r[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        r[i] = ((c & 1) ? H[3 * i] : -H[3 * i]) + ((c & 2) ? H[3 * i + 1] : -H[3 * i + 1]) + ((c & 4) ? H[3 * i + 2] : -H[3 * i + 2]);
      const float gap = cb.pos[2] + r[2] - plane_z;
      if (gap < MARGIN && nc < MAX_CORNERS) {
        float* s = sm + nc * (CORNER_WORDS * SLOT_STRIDE);
        ++nc;
        const float a = r[0] * r[0], b = r[1] * r[1], d = r[2] * r[2];
        const float kn = fmaf(a + b, INV_INERTIA, INV_MASS), k1 = fmaf(d + a, INV_INERTIA, INV_MASS), k2 = fmaf(d + b, INV_INERTIA, INV_MASS);
        s[0 * SLOT_STRIDE] = r[0]; s[1 * SLOT_STRIDE] = r[1]; s[2 * SLOT_STRIDE] = r[2];
        s[3 * SLOT_STRIDE] = gap < 0.f ? -ERP * gap * INV_DT : -gap * INV_DT;
        s[4 * SLOT_STRIDE] = kn; s[5 * SLOT_STRIDE] = k1; s[6 * SLOT_STRIDE] = k2;
        s[7 * SLOT_STRIDE] = rcp_approx(kn); s[8 * SLOT_STRIDE] = rcp_approx(k1); s[9 * SLOT_STRIDE] = rcp_approx(k2);
        s[10 * SLOT_STRIDE] = 0.f; s[11 * SLOT_STRIDE] = 0.f; s[12 * SLOT_STRIDE] = 0.f;
      }
    }
  }
  constexpr int NP = PICK ? 3 : 1;
  {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      Capsule cap;
      arm_capsule<PICK>(ee, Ree, grip, p, cap);
      float rr[3], dir[3][3];
      const float d = capsule_query(cb, R, cap, rr, dir[0]);
      if (d < 0.f) {
        active |= 1u << p;
        tangents(dir[0], dir[1], dir[2]);
        float* s = sm + (MAX_CORNERS * CORNER_WORDS + p * PROXY_WORDS) * SLOT_STRIDE;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float* q = s + k * (ROW_WORDS * SLOT_STRIDE);
          const float x0 = rr[1] * dir[k][2] - rr[2] * dir[k][1], x1 = rr[2] * dir[k][0] - rr[0] * dir[k][2],
                      x2 = rr[0] * dir[k][1] - rr[1] * dir[k][0];
          const float kk = fmaf(x0 * x0 + x1 * x1 + x2 * x2, INV_INERTIA, INV_MASS);
          q[0 * SLOT_STRIDE] = dir[k][0]; q[1 * SLOT_STRIDE] = dir[k][1]; q[2 * SLOT_STRIDE] = dir[k][2];
          q[3 * SLOT_STRIDE] = x0; q[4 * SLOT_STRIDE] = x1; q[5 * SLOT_STRIDE] = x2;
          q[6 * SLOT_STRIDE] = kk; q[7 * SLOT_STRIDE] = rcp_approx(kk); q[8 * SLOT_STRIDE] = 0.f;
        }
        s[3 * ROW_WORDS * SLOT_STRIDE] = -ERP * d * INV_DT;
      }
    }
  }

  // Sweeps: every lane stops at ITS OWN convergence (pybullet's residual test) or after 50; a warp runs as long as its
  // slowest cube (resting cubes take ~7 sweeps, cubes squeezed between a capsule and the table all 50).
  bool act = (active | (unsigned)nc) != 0u;
#pragma unroll 1
  for (int it = 0; act && it < PGS_ITERS; ++it) {
    float res = 0.f;
#pragma unroll 1
    for (int j = 0; j < nc; ++j) {
      float* s = sm + j * (CORNER_WORDS * SLOT_STRIDE);
      float ln = s[10 * SLOT_STRIDE], l1 = s[11 * SLOT_STRIDE], l2 = s[12 * SLOT_STRIDE];
      table_rows(v, w, s[0 * SLOT_STRIDE], s[1 * SLOT_STRIDE], s[2 * SLOT_STRIDE], s[3 * SLOT_STRIDE], s[4 * SLOT_STRIDE],
                 s[5 * SLOT_STRIDE], s[6 * SLOT_STRIDE], s[7 * SLOT_STRIDE], s[8 * SLOT_STRIDE], s[9 * SLOT_STRIDE], ln, l1, l2, res);
      s[10 * SLOT_STRIDE] = ln; s[11 * SLOT_STRIDE] = l1; s[12 * SLOT_STRIDE] = l2;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (active & (1u << p)) {
        float* s = sm + (MAX_CORNERS * CORNER_WORDS + p * PROXY_WORDS) * SLOT_STRIDE;
        const float ln = slot_row(s, v, w, s[3 * ROW_WORDS * SLOT_STRIDE], 0.0f, 1e30f, res);
        const float lim = MU * ln;
        slot_row(s + ROW_WORDS * SLOT_STRIDE, v, w, 0.0f, -lim, lim, res);
        slot_row(s + 2 * ROW_WORDS * SLOT_STRIDE, v, w, 0.0f, -lim, lim, res);
      }
    }
    act = res > PGS_RESIDUAL;
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
