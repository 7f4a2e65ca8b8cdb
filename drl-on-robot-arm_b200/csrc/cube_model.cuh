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
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define CUBE_FN __device__ __forceinline__
#define CUBE_STEP_FN static __device__ __noinline__
#else   // host build of the SAME source (fp32) for tests/cube_host_check.cpp: g++ only, no CUDA
#include <math.h>
#define CUBE_FN static inline
#define CUBE_STEP_FN static inline
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#endif

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
constexpr float PGS_RESIDUAL = 1e-7f;         // pybullet m_leastSquaresResidualThreshold (on the squared velocity change)
constexpr float PGS_RESIDUAL_ROOT = 3.16227766e-4f;
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

CUBE_FN float rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return 1.0f / x;
#endif
}

CUBE_FN void init(State& c, float x, float y, float z, float yaw) {
  c.pos[0] = x; c.pos[1] = y; c.pos[2] = z;
  float s, co;
  sincosf(0.5f * yaw, &s, &co);
  c.quat[0] = 0.f; c.quat[1] = 0.f; c.quat[2] = s; c.quat[3] = co;
#pragma unroll
  for (int i = 0; i < 3; ++i) { c.v[i] = 0.f; c.w[i] = 0.f; }
}

CUBE_FN void rot(const float (&q)[4], float (&R)[9]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

CUBE_FN void tangents(const float (&n)[3], float (&t1)[3], float (&t2)[3]) {  // btPlaneSpace1
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
CUBE_FN float sphere_query(const State& cb, const float (&R)[9], const float (&c)[3], float rad,
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

CUBE_FN void axis_point(const float (&ee)[3], const float (&Ree)[9], float along, float side, float (&o)[3]) {
  o[0] = fmaf(side, Ree[0], fmaf(along, Ree[2], ee[0]));
  o[1] = fmaf(side, Ree[3], fmaf(along, Ree[5], ee[1]));
  o[2] = fmaf(side, Ree[6], fmaf(along, Ree[8], ee[2]));
}

template <bool PICK>
CUBE_FN void arm_capsule(const float (&ee)[3], const float (&Ree)[9], float grip, int p, Capsule& k) {
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
CUBE_FN float capsule_query(const State& cb, const float (&R)[9], const Capsule& k, float (&rrel)[3], float (&n)[3]) {
  const float ab[3] = {k.b[0] - k.a[0], k.b[1] - k.a[1], k.b[2] - k.a[2]};
  const float ac[3] = {cb.pos[0] - k.a[0], cb.pos[1] - k.a[1], cb.pos[2] - k.a[2]};
  const float len2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  float t = (ab[0] * ac[0] + ab[1] * ac[1] + ab[2] * ac[2]) / len2;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float c[3] = {fmaf(t, ab[0], k.a[0]), fmaf(t, ab[1], k.a[1]), fmaf(t, ab[2], k.a[2])};
  return sphere_query(cb, R, c, k.rad, rrel, n);
}

// getClosestPoints(kuka, cube, 0.006) stand-in (rl_pick_env.py:412): min signed distance capsule <-> cube
CUBE_FN float gripper_distance(const State& cb, const float (&ee)[3], const float (&Ree)[9], float grip) {
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


// ---- the contact solve ---------------------------------------------------------------------------------------------
// Same contact list and the same projected-Gauss-Seidel ROW ORDER as oracle/cube_model.h (corner contacts compacted in
// corner order into 4 slots, then the arm capsules; each contact = normal, t1, t2), the same clamps and the same early
// exit -- but arranged for the latency of ONE lane, because at the BASELINE sizes (push 4096, pick 2048: a single
// wave) a launch lasts as long as its slowest cube, and a cube squeezed between a capsule and the table runs all 50
// sweeps:
//
//  * fixed row slots, absent rows masked: every lane runs the same straight-line sweep over 4 corner slots + the
//    capsules; a slot without a contact has 1/k = 0, so its row leaves impulse and twist untouched (bit-exactly), and
//    capsule blocks no lane of the warp needs are skipped warp-uniformly.  No dynamic indexing, no loads that depend on
//    the solve: every row constant sits at a static per-thread shared-memory address and is fetched off the chain.
//  * three-row look-ahead.  Row i of Gauss-Seidel needs J_i . u with u = the twist after row i-1, a 10-operation
//    dependent chain per row (dot, bias, clamp, impulse difference, twist update) that nothing can overlap.  Since
//    u_(i-1) = u_(i-3) + U_(i-2) dl_(i-2) + U_(i-1) dl_(i-1)  (U_j = M^-1 J_j^T),
//        J_i . u_(i-1) = J_i . u_(i-3) + B_(i,i-2) dl_(i-2) + B_(i,i-1) dl_(i-1),      B_(i,j) = J_i . U_j
//    the dot product runs on the twist of three rows ago -- off the critical path -- and the two Delassus entries
//    B (times 1/k_i: c2, c1) are computed once per sim step.  Left on the chain: one FMA, the clamp, one subtraction
//    (4 operations per row instead of 10).  Algebraically identical to the oracle's update, rounding differs at 1e-7.
//
// Per-thread column of dynamic shared memory in 16-byte vectors: vector V of the calling thread is the float4 at
// scratch[(V * SLOT_STRIDE + threadIdx.x) * 4] (a warp reads 512 contiguous bytes: conflict-free LDS.128, a quarter of
// the load instructions of a word layout -- with one warp per scheduler the sweep is bound by instruction issue).
//   corner slot c : V = 4c   (r0, r1, r2, h)        V = 4c+1 / +2 / +3 : row n / t1 / t2 = (1/k, c1, c2, k)
//   capsule p     : V = 16 + 9p + 3j + {0,1,2} for row j = (d0,d1,d2,x0) (x1,x2,1/k,c1) (c2,k,h,-)
// The accumulated impulses and the twist history live in registers.
#if defined(__CUDACC__)
constexpr int SLOT_STRIDE = 128;                 // = LANE_BLOCK (threads per block of every kernel that steps cubes)
#else
constexpr int SLOT_STRIDE = 1;
struct float4 { float x, y, z, w; };
#endif
constexpr int CORNER_VECS = 4, PROXY_VECS = 9;
template <bool PICK> constexpr int scratch_vecs() { return MAX_CORNERS * CORNER_VECS + (PICK ? 3 : 1) * PROXY_VECS; }
template <bool PICK> constexpr int scratch_bytes() { return scratch_vecs<PICK>() * SLOT_STRIDE * 16; }   // push 50 KB, pick 86 KB

#if defined(__CUDACC__)
extern __shared__ __align__(16) float cube_scratch[];
CUBE_FN float* scratch_column() { return cube_scratch + threadIdx.x * 4; }
CUBE_FN bool warp_any(bool p) { return __any_sync(__activemask(), p) != 0; }
#else
static float cube_scratch_host[(MAX_CORNERS * CORNER_VECS + 3 * PROXY_VECS) * 4];
CUBE_FN float* scratch_column() { return cube_scratch_host; }
CUBE_FN bool warp_any(bool p) { return p; }
#endif
CUBE_FN float4 ldv(const float* sm, int V) { return *reinterpret_cast<const float4*>(sm + V * (SLOT_STRIDE * 4)); }
CUBE_FN void stv(float* sm, int V, float a, float b, float c, float d) {
  float4 t; t.x = a; t.y = b; t.z = c; t.w = d;
  *reinterpret_cast<float4*>(sm + V * (SLOT_STRIDE * 4)) = t;
}

struct Twist { float v[3], w[3]; };
// a = twist after row i-3, b = after row i-2, c = after row i-1; d1 = impulse change of row i-1, d2 = of row i-2
struct History { Twist a, b, c; float d1, d2; };

CUBE_FN void advance(History& H, const Twist& un, float dl) {
  H.a = H.b; H.b = H.c; H.c = un;
  H.d2 = H.d1; H.d1 = dl;
}

// Delassus entry between two rows given as explicit 6-vectors (dir, r x dir): J_a . M^-1 J_b^T  (INV_MASS = 1)
CUBE_FN float delassus(const float (&a)[6], const float (&b)[6]) {
  const float lin = fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
  const float ang = fmaf(a[5], b[5], fmaf(a[4], b[4], a[3] * b[3]));
  return fmaf(ang, INV_INERTIA, lin);
}

// The new accumulated impulse of a row: clamp(lam + (target - J.u_(i-1)) / k) in the look-ahead form; jv = J . u_(i-3),
// lamh = lam + target / k.  On the dependent chain: the last FMA and the clamp.
CUBE_FN float row_impulse(float jv, float ik, float c1, float c2, float lamh, const History& H) {
  float g = fmaf(-ik, jv, lamh);
  g = fmaf(-c2, H.d2, g);
  return fmaf(-c1, H.d1, g);
}

// Largest |velocity change along a row| of the sweep; pybullet's test "squared residual <= 1e-7" becomes
// |dv| <= sqrt(1e-7) at the end of the sweep (one multiply + one max per row).
CUBE_FN void track_residual(float dl, float k, float& res) { res = fmaxf(res, fabsf(dl * k)); }

// The three rows of corner slot c (vectors V .. V+3).  The plane normal is +z, for which btPlaneSpace1 (tangents())
// gives t1 = (0,-1,0), t2 = (1,0,0): J_n = (0,0,1, r1,-r0,0), J_t1 = (0,-1,0, r2,0,-r0), J_t2 = (1,0,0, 0,r2,-r1).
CUBE_FN void corner_rows(const float* sm, int V, float (&lam)[3], History& H, float& res) {
  const float4 g = ldv(sm, V);
  const float r0 = g.x, r1 = g.y, r2 = g.z;
  {  // normal
    const float4 q = ldv(sm, V + 1);                 // 1/k, c1, c2, k
    const float jv = fmaf(r1, H.a.w[0], fmaf(-r0, H.a.w[1], H.a.v[2]));
    const float nl = fmaxf(row_impulse(jv, q.x, q.y, q.z, lam[0] + g.w, H), 0.0f);
    const float dl = nl - lam[0];
    lam[0] = nl;
    const float dli = dl * INV_INERTIA;
    Twist un = H.c;
    un.v[2] += dl;
    un.w[0] = fmaf(dli, r1, un.w[0]);
    un.w[1] = fmaf(-dli, r0, un.w[1]);
    track_residual(dl, q.w, res);
    advance(H, un, dl);
  }
  const float lim = MU * lam[0];
  {  // t1
    const float4 q = ldv(sm, V + 2);
    const float jv = fmaf(r2, H.a.w[0], fmaf(-r0, H.a.w[2], -H.a.v[1]));
    const float nl = fminf(fmaxf(row_impulse(jv, q.x, q.y, q.z, lam[1], H), -lim), lim);
    const float dl = nl - lam[1];
    lam[1] = nl;
    const float dli = dl * INV_INERTIA;
    Twist un = H.c;
    un.v[1] -= dl;
    un.w[0] = fmaf(dli, r2, un.w[0]);
    un.w[2] = fmaf(-dli, r0, un.w[2]);
    track_residual(dl, q.w, res);
    advance(H, un, dl);
  }
  {  // t2
    const float4 q = ldv(sm, V + 3);
    const float jv = fmaf(r2, H.a.w[1], fmaf(-r1, H.a.w[2], H.a.v[0]));
    const float nl = fminf(fmaxf(row_impulse(jv, q.x, q.y, q.z, lam[2], H), -lim), lim);
    const float dl = nl - lam[2];
    lam[2] = nl;
    const float dli = dl * INV_INERTIA;
    Twist un = H.c;
    un.v[0] += dl;
    un.w[1] = fmaf(dli, r2, un.w[1]);
    un.w[2] = fmaf(-dli, r1, un.w[2]);
    track_residual(dl, q.w, res);
    advance(H, un, dl);
  }
}

// One row of a capsule contact (vectors V .. V+2); NORMAL rows add their bias term h and clamp at zero only
template <bool NORMAL>
CUBE_FN void capsule_row(const float* sm, int V, float& lam, float lim, History& H, float& res) {
  const float4 p0 = ldv(sm, V), p1 = ldv(sm, V + 1), p2 = ldv(sm, V + 2);
  const float d0 = p0.x, d1 = p0.y, d2 = p0.z, x0 = p0.w, x1 = p1.x, x2 = p1.y;
  const float jv = fmaf(d0, H.a.v[0], fmaf(d1, H.a.v[1], d2 * H.a.v[2])) + fmaf(x0, H.a.w[0], fmaf(x1, H.a.w[1], x2 * H.a.w[2]));
  float nl = row_impulse(jv, p1.z, p1.w, p2.x, NORMAL ? lam + p2.z : lam, H);
  nl = NORMAL ? fmaxf(nl, 0.0f) : fminf(fmaxf(nl, -lim), lim);
  const float dl = nl - lam;
  lam = nl;
  const float dli = dl * INV_INERTIA;
  Twist un = H.c;
  un.v[0] = fmaf(dl, d0, un.v[0]); un.v[1] = fmaf(dl, d1, un.v[1]); un.v[2] = fmaf(dl, d2, un.v[2]);      // INV_MASS = 1
  un.w[0] = fmaf(dli, x0, un.w[0]); un.w[1] = fmaf(dli, x1, un.w[1]); un.w[2] = fmaf(dli, x2, un.w[2]);
  track_residual(dl, p2.y, res);
  advance(H, un, dl);
}

// explicit 6-vectors of a corner slot's t1 / t2 / n rows (set-up only)
CUBE_FN void corner_J(const float (&r)[3], float (&jn)[6], float (&j1)[6], float (&j2)[6]) {
  jn[0] = 0.f; jn[1] = 0.f;  jn[2] = 1.f; jn[3] = r[1]; jn[4] = -r[0]; jn[5] = 0.f;
  j1[0] = 0.f; j1[1] = -1.f; j1[2] = 0.f; j1[3] = r[2]; j1[4] = 0.f;   j1[5] = -r[0];
  j2[0] = 1.f; j2[1] = 0.f;  j2[2] = 0.f; j2[3] = 0.f;  j2[4] = r[2];  j2[5] = -r[1];
}

// one p.stepSimulation() for the cube; grip: 0 open / push, >= 0.5 fingers closed.  Needs scratch_bytes<PICK>() of
// dynamic shared memory in the calling kernel.
template <bool PICK>
CUBE_STEP_FN void step(State& cbm, const float (&ee)[3], const float (&Ree)[9], float grip) {
  constexpr int NP = PICK ? 3 : 1;
  State cb = cbm;                    // work on a register copy (the caller's object lives behind a reference)
  History H;
#pragma unroll
  for (int i = 0; i < 3; ++i) { H.c.v[i] = cb.v[i] * DAMP; H.c.w[i] = cb.w[i] * DAMP; }
  H.c.v[2] = (cb.v[2] - G * DT) * DAMP;
  H.a = H.c; H.b = H.c;
  H.d1 = 0.f; H.d2 = 0.f;

  float* const sm = scratch_column();
  float R[9];
  rot(cb.quat, R);
  // ---- corner / plane contacts: compacted in corner order into slots 0 .. nc-1 (r and the bias, through the slots)
  int nc = 0;
  {
    const bool on_table = cb.pos[0] >= TABLE_X0 && cb.pos[0] <= TABLE_X1 && cb.pos[1] >= TABLE_Y0 && cb.pos[1] <= TABLE_Y1;
    const float plane_z = on_table ? TABLE_Z : GROUND_Z;
    float Hh[9];                     // half-edge vectors: column k of R times HALF
#pragma unroll
    for (int i = 0; i < 9; ++i) Hh[i] = R[i] * HALF;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float r[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        r[i] = ((c & 1) ? Hh[3 * i] : -Hh[3 * i]) + ((c & 2) ? Hh[3 * i + 1] : -Hh[3 * i + 1]) + ((c & 4) ? Hh[3 * i + 2] : -Hh[3 * i + 2]);
      const float gap = cb.pos[2] + r[2] - plane_z;
      if (gap < MARGIN && nc < MAX_CORNERS) {
        stv(sm, nc * CORNER_VECS, r[0], r[1], r[2], gap < 0.f ? -ERP * gap * INV_DT : -gap * INV_DT);   // bias; becomes h = bias / k below
        ++nc;
      }
    }
  }
  float rc[MAX_CORNERS][3], bias[MAX_CORNERS];
#pragma unroll
  for (int c = 0; c < MAX_CORNERS; ++c) {
    const bool present = c < nc;
    const float4 g = ldv(sm, c * CORNER_VECS);
    rc[c][0] = present ? g.x : 0.f; rc[c][1] = present ? g.y : 0.f; rc[c][2] = present ? g.z : 0.f;
    bias[c] = present ? g.w : 0.f;
  }
  // ---- arm capsules: rows into their slots; prev1 / prev2 = explicit t1 / t2 rows of the block before (GS order)
  float prev1[6], prev2[6], scratch_n[6];
  corner_J(rc[MAX_CORNERS - 1], scratch_n, prev1, prev2);
  unsigned active = 0u;               // bit p : arm capsule p in contact
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    Capsule cap;
    arm_capsule<PICK>(ee, Ree, grip, p, cap);
    float rr[3], dir[3][3];
    const float d = capsule_query(cb, R, cap, rr, dir[0]);
    const bool on = d < 0.f;
    if (on) active |= 1u << p;
    tangents(dir[0], dir[1], dir[2]);
    float J[3][6], kk[3], ik[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      J[k][0] = on ? dir[k][0] : 0.f; J[k][1] = on ? dir[k][1] : 0.f; J[k][2] = on ? dir[k][2] : 0.f;
      J[k][3] = on ? rr[1] * dir[k][2] - rr[2] * dir[k][1] : 0.f;
      J[k][4] = on ? rr[2] * dir[k][0] - rr[0] * dir[k][2] : 0.f;
      J[k][5] = on ? rr[0] * dir[k][1] - rr[1] * dir[k][0] : 0.f;
      const float k0 = fmaf(J[k][3] * J[k][3] + J[k][4] * J[k][4] + J[k][5] * J[k][5], INV_INERTIA, INV_MASS);
      kk[k] = on ? k0 : 0.f;
      ik[k] = on ? rcp_approx(k0) : 0.f;
    }
    const float c1[3] = {ik[0] * delassus(J[0], prev2), ik[1] * delassus(J[1], J[0]), ik[2] * delassus(J[2], J[1])};
    const float c2[3] = {ik[0] * delassus(J[0], prev1), ik[1] * delassus(J[1], prev2), ik[2] * delassus(J[2], J[0])};
    const float hcap = on ? ik[0] * (-ERP * d * INV_DT) : 0.f;                // h = target / k of the normal row
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int V = MAX_CORNERS * CORNER_VECS + p * PROXY_VECS + 3 * k;
      stv(sm, V, J[k][0], J[k][1], J[k][2], J[k][3]);
      stv(sm, V + 1, J[k][4], J[k][5], ik[k], c1[k]);
      stv(sm, V + 2, c2[k], kk[k], k == 0 ? hcap : 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) { prev1[i] = J[1][i]; prev2[i] = J[2][i]; }
  }
  // ---- corner slots: effective masses and the look-ahead couplings (closed forms for a +z normal)
#pragma unroll
  for (int c = 0; c < MAX_CORNERS; ++c) {
    const bool present = c < nc;
    const float r0 = rc[c][0], r1 = rc[c][1], r2 = rc[c][2];
    const float a = r0 * r0, b = r1 * r1, d = r2 * r2;
    const float kn = fmaf(a + b, INV_INERTIA, INV_MASS), k1 = fmaf(d + a, INV_INERTIA, INV_MASS), k2 = fmaf(d + b, INV_INERTIA, INV_MASS);
    const float ikn = present ? rcp_approx(kn) : 0.f, ik1 = present ? rcp_approx(k1) : 0.f, ik2 = present ? rcp_approx(k2) : 0.f;
    float bn2, bn1, b12;              // B(n, prev t2), B(n, prev t1), B(t1, prev t2)
    if (c == 0) {                     // the block before slot 0 is the last capsule (previous sweep)
      float jn[6], j1[6], j2[6];
      corner_J(rc[0], jn, j1, j2);
      bn2 = delassus(jn, prev2); bn1 = delassus(jn, prev1); b12 = delassus(j1, prev2);
    } else {
      const float p1 = rc[c - 1][1], p2 = rc[c - 1][2];
      bn2 = -INV_INERTIA * r0 * p2; bn1 = INV_INERTIA * r1 * p2; b12 = INV_INERTIA * r0 * p1;
    }
    const int V = c * CORNER_VECS;
    stv(sm, V, present ? r0 : 0.f, present ? r1 : 0.f, present ? r2 : 0.f, ikn * bias[c]);
    //              1/k   c1: coupling with the row before   c2: with the row two before        k
    stv(sm, V + 1, ikn, ikn * bn2,                           ikn * bn1,                         present ? kn : 0.f);
    stv(sm, V + 2, ik1, ik1 * (INV_INERTIA * r1 * r2),       ik1 * b12,                         present ? k1 : 0.f);   // B(t1, n)
    stv(sm, V + 3, ik2, ik2 * (INV_INERTIA * r0 * r1),       ik2 * (-INV_INERTIA * r0 * r2),    present ? k2 : 0.f);   // B(t2, t1), B(t2, n)
  }

  // ---- sweeps: every lane stops at ITS OWN convergence (pybullet's residual test) or after 50; a warp runs as long
  // as its slowest cube (resting cubes take ~7 sweeps, cubes squeezed between a capsule and the table all 50)
  float lam[MAX_CORNERS][3], lamc[NP][3];
#pragma unroll
  for (int c = 0; c < MAX_CORNERS; ++c) { lam[c][0] = 0.f; lam[c][1] = 0.f; lam[c][2] = 0.f; }
#pragma unroll
  for (int p = 0; p < NP; ++p) { lamc[p][0] = 0.f; lamc[p][1] = 0.f; lamc[p][2] = 0.f; }
  bool any_cap[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) any_cap[p] = warp_any((active >> p) & 1u);
  bool act = (active | (unsigned)nc) != 0u;
#pragma unroll 1
  for (int it = 0; act && it < PGS_ITERS; ++it) {
    float res = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_CORNERS; ++c) corner_rows(sm, c * CORNER_VECS, lam[c], H, res);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (any_cap[p]) {
        const int V = MAX_CORNERS * CORNER_VECS + p * PROXY_VECS;
        capsule_row<true>(sm, V, lamc[p][0], 0.f, H, res);
        const float lim = MU * lamc[p][0];
        capsule_row<false>(sm, V + 3, lamc[p][1], lim, H, res);
        capsule_row<false>(sm, V + 6, lamc[p][2], lim, H, res);
      } else {                        // three rows that change nothing: the history collapses onto the current twist
        H.a = H.c; H.b = H.c;
        H.d1 = 0.f; H.d2 = 0.f;
      }
    }
    act = res > PGS_RESIDUAL_ROOT;
  }
  const float (&v)[3] = H.c.v;
  const float (&w)[3] = H.c.w;
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
