// aba_device.cuh -- torque mode: articulated-body forward dynamics of the 7-DoF chain, per arm, in registers (fp32).
//
// No counterpart in the reference (its envs teleport the joints, rl_reach_env.py:252-257); this is the north star's
// "articulated-body forward dynamics built from the repo's URDFs, joint-limit clamp".  The fp64 statement of the
// same algorithm is oracle/aba_model.h (aba_forward_dynamics / aba_torque_step), itself cross-checked against a
// dense M(q)^-1 (tau - h) route built from recursive Newton-Euler.
//
// Featherstone's ABA specialised to revolute-z joints on a serial chain with a fixed base; spatial vectors are
// [angular; linear] in link coordinates; the articulated inertia is carried as blocks [[A, B], [B^T, D]] with A, D
// symmetric (6 values each: xx xy xz yy yz zz) -- because the chain is serial only ONE accumulated inertia is live
// during the inward sweep.  What survives between sweeps is small: sin/cos of the joint angles (the link rotations
// are rebuilt from them: 12 FMAs), the link velocities, and U_i, 1/d_i, u_i.
#pragma once
#include "armsim_device.cuh"

struct DynParams {
  float Ibar[NJ][6];   // rotational inertia about the LINK ORIGIN (xx xy xz yy yz zz) = Ic + m (|c|^2 1 - c c^T)
  float h[NJ][3];      // m * com
  float mass[NJ];
  float effort[NJ], maxvel[NJ], damping[NJ];
  float abase[3];      // -gravity expressed in base coordinates (gravity enters as a base acceleration)
  float dt;
};

namespace aba {

__device__ __forceinline__ void cross(const float (&a)[3], const float (&b)[3], float (&c)[3]) {
  c[0] = fmaf(a[1], b[2], -a[2] * b[1]);
  c[1] = fmaf(a[2], b[0], -a[0] * b[2]);
  c[2] = fmaf(a[0], b[1], -a[1] * b[0]);
}
__device__ __forceinline__ void mv(const float (&M)[9], const float (&v)[3], float (&o)[3]) {     // o = M v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = fmaf(M[3 * i + 2], v[2], fmaf(M[3 * i + 1], v[1], M[3 * i] * v[0]));
}
__device__ __forceinline__ void mtv(const float (&M)[9], const float (&v)[3], float (&o)[3]) {    // o = M^T v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = fmaf(M[6 + i], v[2], fmaf(M[3 + i], v[1], M[i] * v[0]));
}
// symmetric 3x3 stored xx xy xz yy yz zz
__device__ __forceinline__ void symv(const float (&S)[6], const float (&v)[3], float (&o)[3]) {
  o[0] = fmaf(S[2], v[2], fmaf(S[1], v[1], S[0] * v[0]));
  o[1] = fmaf(S[4], v[2], fmaf(S[3], v[1], S[1] * v[0]));
  o[2] = fmaf(S[5], v[2], fmaf(S[4], v[1], S[2] * v[0]));
}
// R S R^T for symmetric S -> symmetric
__device__ __forceinline__ void rot_sym(const float (&R)[9], const float (&S)[6], float (&O)[6]) {
  float T[9];   // T = R S
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T[3 * i + 0] = fmaf(R[3 * i + 2], S[2], fmaf(R[3 * i + 1], S[1], R[3 * i] * S[0]));
    T[3 * i + 1] = fmaf(R[3 * i + 2], S[4], fmaf(R[3 * i + 1], S[3], R[3 * i] * S[1]));
    T[3 * i + 2] = fmaf(R[3 * i + 2], S[5], fmaf(R[3 * i + 1], S[4], R[3 * i] * S[2]));
  }
  auto dot = [&](int i, int j) { return fmaf(T[3 * i + 2], R[3 * j + 2], fmaf(T[3 * i + 1], R[3 * j + 1], T[3 * i] * R[3 * j])); };
  O[0] = dot(0, 0); O[1] = dot(0, 1); O[2] = dot(0, 2); O[3] = dot(1, 1); O[4] = dot(1, 2); O[5] = dot(2, 2);
}
// R B R^T for a general B
__device__ __forceinline__ void rot_gen(const float (&R)[9], const float (&B)[9], float (&O)[9]) {
  float T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[3 * i + j] = fmaf(R[3 * i + 2], B[6 + j], fmaf(R[3 * i + 1], B[3 + j], R[3 * i] * B[j]));
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) O[3 * i + j] = fmaf(T[3 * i + 2], R[3 * j + 2], fmaf(T[3 * i + 1], R[3 * j + 1], T[3 * i] * R[3 * j]));
}

// R_i = Rf_i Rz(q_i): link-i coordinates -> parent coordinates
__device__ __forceinline__ void link_rot(const ChainParams& C, int i, float s, float c, float (&R)[9]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    R[3 * r] = fmaf(C.Rf[i][3 * r + 1], s, C.Rf[i][3 * r] * c);
    R[3 * r + 1] = fmaf(C.Rf[i][3 * r + 1], c, -C.Rf[i][3 * r] * s);
    R[3 * r + 2] = C.Rf[i][3 * r + 2];
  }
}

// qdd = ABA(q, qd, tau)
__device__ __forceinline__ void forward_dynamics(const ChainParams& C, const DynParams& Dn, const float (&q)[NJ],
                                                 const float (&qd)[NJ], const float (&tau)[NJ], float (&qdd)[NJ]) {
  float sn[NJ], cs[NJ], vw[NJ][3], vv[NJ][3];
  // ---- sweep 1: link velocities, base -> tip
  {
    float wp[3] = {0.f, 0.f, 0.f}, vp[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      sincos_bounded(q[i], sn[i], cs[i]);
      float R[9];
      link_rot(C, i, sn[i], cs[i], R);
      const float r[3] = {C.t[i][0], C.t[i][1], C.t[i][2]};
      float wxr[3], tmp[3];
      cross(wp, r, wxr);
#pragma unroll
      for (int k = 0; k < 3; ++k) tmp[k] = vp[k] + wxr[k];
      mtv(R, wp, vw[i]);
      mtv(R, tmp, vv[i]);
      vw[i][2] += qd[i];
#pragma unroll
      for (int k = 0; k < 3; ++k) { wp[k] = vw[i][k]; vp[k] = vv[i][k]; }
    }
  }
  // ---- sweep 2: articulated inertia and bias force, tip -> base (accumulators live in the current link's frame)
  float U[NJ][6], dinv[NJ], u[NJ];
  {
    float A[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, B[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
          D[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pn[3] = {0.f, 0.f, 0.f}, pf[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = NJ - 1; i >= 0; --i) {
      const float m = Dn.mass[i];
      const float h[3] = {Dn.h[i][0], Dn.h[i][1], Dn.h[i][2]};
      // own rigid-body inertia: [[Ibar, hx], [hx^T, m 1]]
#pragma unroll
      for (int k = 0; k < 6; ++k) A[k] += Dn.Ibar[i][k];
      B[1] -= h[2]; B[2] += h[1]; B[3] += h[2]; B[5] -= h[0]; B[6] -= h[1]; B[7] += h[0];
      D[0] += m; D[3] += m; D[5] += m;
      // own bias force p = v x* (I v),  I v = [Ibar w + h x v ; m v - h x w]
      {
        const float (&w)[3] = vw[i];
        const float (&v)[3] = vv[i];
        float Iw[3], hxv[3], hxw[3], Ln[3], Lf[3], t1[3], t2[3], t3[3];
        symv(Dn.Ibar[i], w, Iw);
        cross(h, v, hxv);
        cross(h, w, hxw);
#pragma unroll
        for (int k = 0; k < 3; ++k) { Ln[k] = Iw[k] + hxv[k]; Lf[k] = fmaf(m, v[k], -hxw[k]); }
        cross(w, Ln, t1);
        cross(v, Lf, t2);
        cross(w, Lf, t3);
#pragma unroll
        for (int k = 0; k < 3; ++k) { pn[k] += t1[k] + t2[k]; pf[k] += t3[k]; }
      }
      // U = IA S (S = [z; 0]),  d = S^T U,  u = tau - S^T pA
      U[i][0] = A[2]; U[i][1] = A[4]; U[i][2] = A[5];
      U[i][3] = B[6]; U[i][4] = B[7]; U[i][5] = B[8];
      dinv[i] = fast_rcp(A[5]);
      u[i] = tau[i] - pn[2];
      if (i == 0) break;
      const float di = dinv[i], ud = u[i] * di;
      const float Ua[3] = {U[i][0], U[i][1], U[i][2]}, Ub[3] = {U[i][3], U[i][4], U[i][5]};
      // Ia = IA - U U^T / d
      {
        const float s0 = Ua[0] * di, s1 = Ua[1] * di, s2 = Ua[2] * di;
        A[0] = fmaf(-s0, Ua[0], A[0]); A[1] = fmaf(-s0, Ua[1], A[1]); A[2] = fmaf(-s0, Ua[2], A[2]);
        A[3] = fmaf(-s1, Ua[1], A[3]); A[4] = fmaf(-s1, Ua[2], A[4]); A[5] = fmaf(-s2, Ua[2], A[5]);
        const float sa[3] = {s0, s1, s2};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) B[3 * a + b] = fmaf(-sa[a], Ub[b], B[3 * a + b]);
        const float r0 = Ub[0] * di, r1 = Ub[1] * di, r2 = Ub[2] * di;
        D[0] = fmaf(-r0, Ub[0], D[0]); D[1] = fmaf(-r0, Ub[1], D[1]); D[2] = fmaf(-r0, Ub[2], D[2]);
        D[3] = fmaf(-r1, Ub[1], D[3]); D[4] = fmaf(-r1, Ub[2], D[4]); D[5] = fmaf(-r2, Ub[2], D[5]);
      }
      // pa = pA + Ia c + U u / d,   c = qd [w x z ; v x z] = qd (w_y, -w_x, 0 ; v_y, -v_x, 0)
      {
        const float cw[3] = {qd[i] * vw[i][1], -qd[i] * vw[i][0], 0.f};
        const float cv[3] = {qd[i] * vv[i][1], -qd[i] * vv[i][0], 0.f};
        float t1[3], t2[3], t3[3], t4[3];
        symv(A, cw, t1);
        mv(B, cv, t2);
        mtv(B, cw, t3);
        symv(D, cv, t4);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          pn[k] += t1[k] + t2[k] + Ua[k] * ud;
          pf[k] += t3[k] + t4[k] + Ub[k] * ud;
        }
      }
      // hand everything to the parent: rotate (X' = R X R^T), then shift the origin by r = t_i
      {
        float R[9], Ar[6], Br[9], Dr[6];
        link_rot(C, i, sn[i], cs[i], R);
        rot_sym(R, A, Ar);
        rot_gen(R, B, Br);
        rot_sym(R, D, Dr);
        const float r0 = C.t[i][0], r1 = C.t[i][1], r2 = C.t[i][2];
        const float Dm[9] = {Dr[0], Dr[1], Dr[2], Dr[1], Dr[3], Dr[4], Dr[2], Dr[4], Dr[5]};
        float Bp[9];
#pragma unroll
        for (int j = 0; j < 3; ++j) {       // Bp = Br + [r]x Dr
          Bp[j] = Br[j] + fmaf(r1, Dm[6 + j], -r2 * Dm[3 + j]);
          Bp[3 + j] = Br[3 + j] + fmaf(r2, Dm[j], -r0 * Dm[6 + j]);
          Bp[6 + j] = Br[6 + j] + fmaf(r0, Dm[3 + j], -r1 * Dm[j]);
        }
        // Ap = Ar + [r]x Br^T - Bp [r]x   (symmetric; only the 6 unique entries)
        auto rxBt = [&](int a, int b) {     // ([r]x Br^T)[a][b] = sum_k [r]x[a][k] Br[b][k]
          return a == 0 ? fmaf(r1, Br[3 * b + 2], -r2 * Br[3 * b + 1])
               : a == 1 ? fmaf(r2, Br[3 * b + 0], -r0 * Br[3 * b + 2])
                        : fmaf(r0, Br[3 * b + 1], -r1 * Br[3 * b + 0]);
        };
        auto Bprx = [&](int a, int b) {     // (Bp [r]x)[a][b] = sum_k Bp[a][k] [r]x[k][b]
          return b == 0 ? fmaf(Bp[3 * a + 1], r2, -Bp[3 * a + 2] * r1)
               : b == 1 ? fmaf(Bp[3 * a + 2], r0, -Bp[3 * a + 0] * r2)
                        : fmaf(Bp[3 * a + 0], r1, -Bp[3 * a + 1] * r0);
        };
        A[0] = Ar[0] + rxBt(0, 0) - Bprx(0, 0);
        A[1] = Ar[1] + rxBt(0, 1) - Bprx(0, 1);
        A[2] = Ar[2] + rxBt(0, 2) - Bprx(0, 2);
        A[3] = Ar[3] + rxBt(1, 1) - Bprx(1, 1);
        A[4] = Ar[4] + rxBt(1, 2) - Bprx(1, 2);
        A[5] = Ar[5] + rxBt(2, 2) - Bprx(2, 2);
#pragma unroll
        for (int k = 0; k < 9; ++k) B[k] = Bp[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) D[k] = Dr[k];
        float fpar[3], npar[3], rxf[3];
        mv(R, pf, fpar);
        mv(R, pn, npar);
        const float rr[3] = {r0, r1, r2};
        cross(rr, fpar, rxf);
#pragma unroll
        for (int k = 0; k < 3; ++k) { pf[k] = fpar[k]; pn[k] = npar[k] + rxf[k]; }
      }
    }
  }
  // ---- sweep 3: accelerations, base -> tip
  {
    float aw[3] = {0.f, 0.f, 0.f}, av[3] = {Dn.abase[0], Dn.abase[1], Dn.abase[2]};
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      float R[9];
      link_rot(C, i, sn[i], cs[i], R);
      const float r[3] = {C.t[i][0], C.t[i][1], C.t[i][2]};
      float axr[3], tmp[3], w2[3], v2[3];
      cross(aw, r, axr);
#pragma unroll
      for (int k = 0; k < 3; ++k) tmp[k] = av[k] + axr[k];
      mtv(R, aw, w2);
      mtv(R, tmp, v2);
      w2[0] = fmaf(qd[i], vw[i][1], w2[0]); w2[1] = fmaf(-qd[i], vw[i][0], w2[1]);
      v2[0] = fmaf(qd[i], vv[i][1], v2[0]); v2[1] = fmaf(-qd[i], vv[i][0], v2[1]);
      float Ua = U[i][0] * w2[0];
      Ua = fmaf(U[i][1], w2[1], Ua); Ua = fmaf(U[i][2], w2[2], Ua);
      Ua = fmaf(U[i][3], v2[0], Ua); Ua = fmaf(U[i][4], v2[1], Ua); Ua = fmaf(U[i][5], v2[2], Ua);
      qdd[i] = (u[i] - Ua) * dinv[i];
      w2[2] += qdd[i];
#pragma unroll
      for (int k = 0; k < 3; ++k) { aw[k] = w2[k]; av[k] = v2[k]; }
    }
  }
}

// One torque-mode integration step (oracle/aba_model.h aba_torque_step): effort clip, joint damping, ABA,
// semi-implicit Euler, velocity clip, joint-limit clamp with the velocity zeroed on the active side.
__device__ __forceinline__ void torque_step(const ChainParams& C, const DynParams& Dn, const float (&cmd)[NJ],
                                            float (&q)[NJ], float (&qd)[NJ]) {
  float tau[NJ], qdd[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) tau[j] = fmaf(-Dn.damping[j], qd[j], clampf(cmd[j], -Dn.effort[j], Dn.effort[j]));
  forward_dynamics(C, Dn, q, qd, tau, qdd);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float v = clampf(fmaf(qdd[j], Dn.dt, qd[j]), -Dn.maxvel[j], Dn.maxvel[j]);
    float x = fmaf(v, Dn.dt, q[j]);
    if (x < C.lower[j]) { x = C.lower[j]; v = fmaxf(v, 0.f); }
    if (x > C.upper[j]) { x = C.upper[j]; v = fminf(v, 0.f); }
    q[j] = x;
    qd[j] = v;
  }
}

}  // namespace aba
