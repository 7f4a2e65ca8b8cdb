/* aba_model.h -- TEST INFRASTRUCTURE (oracle side, fp64): rigid-body dynamics of the 7-DoF serial chain.
 *
 * Torque mode (ARMSIM_MODE_TORQUE) has NO counterpart in the reference: every reference env teleports the joints
 * (resetJointState, rl_reach_env.py:252-257) and never applies a torque.  It exists because
 * BASELINE.json's north star asks for "articulated-body forward dynamics built from the repo's URDFs, joint-limit
 * clamp".  PARITY UNPINNED by construction; correctness is established by two independent algorithms agreeing:
 *
 *   aba_forward_dynamics   Featherstone's articulated-body algorithm, O(n), three sweeps over the chain
 *                          (the algorithm the CUDA kernel implements in fp32, csrc/aba_device.cuh)
 *   dense_forward_dynamics qdd = M(q)^-1 (tau - h(q, qd)) with h from recursive Newton-Euler and M built column by
 *                          column from RNEA with unit accelerations, solved by Gaussian elimination
 *
 * Model data: models/diana/DianaS1_robot.urdf:29-216 (masses, COMs, inertia tensors, efforts, velocity limits) and
 * the Kuka iiwa table of SURVEY Appendix A (masses / inertias recalled, unpinned), via include/armsim_robot_models.h.
 * Conventions: spatial vectors are [angular; linear] in LINK coordinates; link i frame = joint i frame rotated by
 * q_i about its z; joint i sits at xyz_i / rpy_i in the parent link frame; base fixed; gravity enters as a base
 * acceleration of -g.
 */
#ifndef ORACLE_ABA_MODEL_H
#define ORACLE_ABA_MODEL_H
#include <math.h>
#include <string.h>

#define ABA_NJ 7

typedef struct AbaChain {
  double Rb[9];               /* base orientation in the world (row-major) */
  double Rf[ABA_NJ][9];       /* fixed rotation of joint i's frame in the parent link frame */
  double t[ABA_NJ][3];        /* joint i origin in the parent link frame */
  double mass[ABA_NJ], com[ABA_NJ][3], Ic[ABA_NJ][6];   /* ixx ixy ixz iyy iyz izz about the COM */
  double gravity[3];          /* world, e.g. (0, 0, -10): rl_reach_env.py:142 */
} AbaChain;

static inline void aba_cross(const double a[3], const double b[3], double c[3]) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x; c[1] = y; c[2] = z;
}
static inline void aba_mv(const double M[9], const double v[3], double o[3]) {       /* o = M v */
  double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2],
         z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void aba_mtv(const double M[9], const double v[3], double o[3]) {      /* o = M^T v */
  double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2],
         z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void aba_mm(const double A[9], const double B[9], double C[9]) {
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof(T));
}
static inline void aba_transpose(const double A[9], double T[9]) {
  double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
  memcpy(T, t, sizeof(t));
}
static inline void aba_skew(const double r[3], double S[9]) {
  S[0] = 0; S[1] = -r[2]; S[2] = r[1];
  S[3] = r[2]; S[4] = 0; S[5] = -r[0];
  S[6] = -r[1]; S[7] = r[0]; S[8] = 0;
}

/* R_i = Rf_i Rz(q_i): link-i coordinates -> parent coordinates */
static inline void aba_link_rot(const AbaChain* c, int i, double q, double R[9]) {
  double cq = cos(q), sq = sin(q);
  double Rz[9] = {cq, -sq, 0, sq, cq, 0, 0, 0, 1};
  aba_mm(c->Rf[i], Rz, R);
}

/* spatial inertia of link i at its frame origin: I [w; v] = [Ibar w + h x v ; m v - h x w], h = m com,
 * Ibar = Ic + m (|c|^2 1 - c c^T) */
static inline void aba_link_inertia(const AbaChain* c, int i, double Ibar[9], double h[3]) {
  const double m = c->mass[i], *cm = c->com[i], *I = c->Ic[i];
  const double cc = cm[0] * cm[0] + cm[1] * cm[1] + cm[2] * cm[2];
  double Icm[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) Ibar[3 * a + b] = Icm[3 * a + b] + m * ((a == b ? cc : 0.0) - cm[a] * cm[b]);
  for (int a = 0; a < 3; ++a) h[a] = m * cm[a];
}

/* ---------------------------------------------------------------------------------------------- RNEA
 * tau = M(q) qdd + C(q, qd) qd + g(q)   (with_gravity = 0 drops g) */
static inline void rnea_inverse_dynamics(const AbaChain* c, const double q[ABA_NJ], const double qd[ABA_NJ],
                                         const double qdd[ABA_NJ], int with_gravity, double tau[ABA_NJ]) {
  double R[ABA_NJ][9], w[ABA_NJ][3], al[ABA_NJ][3], a[ABA_NJ][3], n[ABA_NJ][3], f[ABA_NJ][3];
  double wp[3] = {0, 0, 0}, alp[3] = {0, 0, 0}, ap[3] = {0, 0, 0};
  if (with_gravity) {           /* base acceleration -g, expressed in base coordinates */
    double mg[3] = {-c->gravity[0], -c->gravity[1], -c->gravity[2]};
    aba_mtv(c->Rb, mg, ap);
  }
  for (int i = 0; i < ABA_NJ; ++i) {
    aba_link_rot(c, i, q[i], R[i]);
    /* parent quantities at the child origin, parent coordinates */
    double t1[3], t2[3], accp[3];
    aba_cross(alp, c->t[i], t1);
    aba_cross(wp, c->t[i], t2);
    aba_cross(wp, t2, t2);
    for (int k = 0; k < 3; ++k) accp[k] = ap[k] + t1[k] + t2[k];    /* classical acceleration of the child origin */
    aba_mtv(R[i], wp, w[i]);
    aba_mtv(R[i], alp, al[i]);
    aba_mtv(R[i], accp, a[i]);
    /* joint motion about local z */
    double zq[3] = {0, 0, qd[i]}, wxz[3];
    aba_cross(w[i], zq, wxz);
    for (int k = 0; k < 3; ++k) al[i][k] += wxz[k];
    al[i][2] += qdd[i];
    w[i][2] += qd[i];
    memcpy(wp, w[i], sizeof(wp)); memcpy(alp, al[i], sizeof(alp)); memcpy(ap, a[i], sizeof(ap));
  }
  for (int i = 0; i < ABA_NJ; ++i) {     /* Newton-Euler at the COM */
    const double* cm = c->com[i];
    double t1[3], t2[3], ac[3];
    aba_cross(al[i], cm, t1);
    aba_cross(w[i], cm, t2);
    aba_cross(w[i], t2, t2);
    for (int k = 0; k < 3; ++k) ac[k] = a[i][k] + t1[k] + t2[k];
    for (int k = 0; k < 3; ++k) f[i][k] = c->mass[i] * ac[k];
    const double* I = c->Ic[i];
    double Icm[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]};
    double Ial[3], Iw[3], wIw[3], cxf[3];
    aba_mv(Icm, al[i], Ial);
    aba_mv(Icm, w[i], Iw);
    aba_cross(w[i], Iw, wIw);
    aba_cross(cm, f[i], cxf);
    for (int k = 0; k < 3; ++k) n[i][k] = Ial[k] + wIw[k] + cxf[k];  /* moment about the link origin */
  }
  for (int i = ABA_NJ - 1; i >= 0; --i) {
    tau[i] = n[i][2];
    if (i > 0) {                         /* hand force / moment to the parent (link i-1 coordinates) */
      double fp[3], np[3], rxf[3];
      aba_mv(R[i], f[i], fp);
      aba_mv(R[i], n[i], np);
      aba_cross(c->t[i], fp, rxf);
      for (int k = 0; k < 3; ++k) { f[i - 1][k] += fp[k]; n[i - 1][k] += np[k] + rxf[k]; }
    }
  }
}

static inline void aba_solve_gepp(int n, double* A, double* b, double* x) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(A[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(A[r * n + c]) > best) { best = fabs(A[r * n + c]); piv = r; }
    if (piv != c) {
      for (int k = 0; k < n; ++k) { double t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < n; ++r) {
      double f = A[r * n + c] / A[c * n + c];
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double acc = b[r];
    for (int k = r + 1; k < n; ++k) acc -= A[r * n + k] * x[k];
    x[r] = acc / A[r * n + r];
  }
}

/* M (7x7 row-major, optional) and qdd from the dense route */
static inline void dense_forward_dynamics(const AbaChain* c, const double q[ABA_NJ], const double qd[ABA_NJ],
                                          const double tau[ABA_NJ], double qdd[ABA_NJ], double* Mout) {
  double zero[ABA_NJ] = {0}, h[ABA_NJ], M[ABA_NJ * ABA_NJ], rhs[ABA_NJ];
  rnea_inverse_dynamics(c, q, qd, zero, 1, h);
  for (int j = 0; j < ABA_NJ; ++j) {
    double e[ABA_NJ] = {0}, col[ABA_NJ];
    e[j] = 1.0;
    rnea_inverse_dynamics(c, q, zero, e, 0, col);
    for (int i = 0; i < ABA_NJ; ++i) M[i * ABA_NJ + j] = col[i];
  }
  if (Mout) memcpy(Mout, M, sizeof(M));
  for (int i = 0; i < ABA_NJ; ++i) rhs[i] = tau[i] - h[i];
  aba_solve_gepp(ABA_NJ, M, rhs, qdd);
}

/* ---------------------------------------------------------------------------------------------- ABA
 * Featherstone, "Rigid Body Dynamics Algorithms" (2008), Table 7.1, specialised to revolute-z joints.
 * Articulated inertia as blocks IA = [[A, B], [B^T, D]] acting on [w; v]. */
static inline void aba_forward_dynamics(const AbaChain* c, const double q[ABA_NJ], const double qd[ABA_NJ],
                                        const double tau[ABA_NJ], double qdd[ABA_NJ]) {
  double R[ABA_NJ][9];
  double vw[ABA_NJ][3], vv[ABA_NJ][3];          /* link spatial velocity */
  double cw[ABA_NJ][3], cv[ABA_NJ][3];          /* velocity-product acceleration */
  double A[ABA_NJ][9], B[ABA_NJ][9], D[ABA_NJ][9], pn[ABA_NJ][3], pf[ABA_NJ][3];
  double U[ABA_NJ][6], dinv[ABA_NJ], u[ABA_NJ];
  /* sweep 1: velocities, bias terms, rigid inertias */
  for (int i = 0; i < ABA_NJ; ++i) {
    aba_link_rot(c, i, q[i], R[i]);
    double wp[3] = {0, 0, 0}, vp[3] = {0, 0, 0};
    if (i > 0) { memcpy(wp, vw[i - 1], sizeof(wp)); memcpy(vp, vv[i - 1], sizeof(vp)); }
    double wxr[3], tmp[3];
    aba_cross(wp, c->t[i], wxr);
    for (int k = 0; k < 3; ++k) tmp[k] = vp[k] + wxr[k];
    aba_mtv(R[i], wp, vw[i]);
    aba_mtv(R[i], tmp, vv[i]);
    /* c = v x (S qd), S = [z; 0], using the velocity BEFORE adding the joint's own rate (z x z = 0 anyway) */
    double zq[3] = {0, 0, qd[i]};
    aba_cross(vw[i], zq, cw[i]);
    aba_cross(vv[i], zq, cv[i]);
    vw[i][2] += qd[i];
    double Ibar[9], h[3];
    aba_link_inertia(c, i, Ibar, h);
    double hx[9];
    aba_skew(h, hx);
    memcpy(A[i], Ibar, sizeof(Ibar));
    memcpy(B[i], hx, sizeof(hx));
    for (int k = 0; k < 9; ++k) D[i][k] = 0.0;
    D[i][0] = D[i][4] = D[i][8] = c->mass[i];
    /* p = v x* (I v):  I v = [Ibar w + h x v ; m v - h x w] */
    double Iw[3], hxv[3], hxw[3], Ln[3], Lf[3], t1[3], t2[3];
    aba_mv(Ibar, vw[i], Iw);
    aba_cross(h, vv[i], hxv);
    aba_cross(h, vw[i], hxw);
    for (int k = 0; k < 3; ++k) { Ln[k] = Iw[k] + hxv[k]; Lf[k] = c->mass[i] * vv[i][k] - hxw[k]; }
    aba_cross(vw[i], Ln, t1);
    aba_cross(vv[i], Lf, t2);
    for (int k = 0; k < 3; ++k) pn[i][k] = t1[k] + t2[k];
    aba_cross(vw[i], Lf, pf[i]);
  }
  /* sweep 2: articulated inertias and bias forces, tip to base */
  for (int i = ABA_NJ - 1; i >= 0; --i) {
    /* U = IA S: [A[:,2] ; B^T[:,2]] */
    U[i][0] = A[i][2]; U[i][1] = A[i][5]; U[i][2] = A[i][8];
    U[i][3] = B[i][6]; U[i][4] = B[i][7]; U[i][5] = B[i][8];
    dinv[i] = 1.0 / U[i][2];
    u[i] = tau[i] - pn[i][2];
    if (i == 0) break;
    /* Ia = IA - U U^T / d */
    double Aa[9], Ba[9], Da[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        Aa[3 * a + b] = A[i][3 * a + b] - U[i][a] * U[i][b] * dinv[i];
        Ba[3 * a + b] = B[i][3 * a + b] - U[i][a] * U[i][3 + b] * dinv[i];
        Da[3 * a + b] = D[i][3 * a + b] - U[i][3 + a] * U[i][3 + b] * dinv[i];
      }
    /* pa = pA + Ia c + U u / d */
    double pan[3], paf[3], t1[3], t2[3];
    aba_mv(Aa, cw[i], t1); aba_mv(Ba, cv[i], t2);
    for (int k = 0; k < 3; ++k) pan[k] = pn[i][k] + t1[k] + t2[k] + U[i][k] * u[i] * dinv[i];
    aba_mtv(Ba, cw[i], t1); aba_mv(Da, cv[i], t2);
    for (int k = 0; k < 3; ++k) paf[k] = pf[i][k] + t1[k] + t2[k] + U[i][3 + k] * u[i] * dinv[i];
    /* to the parent: rotate the blocks (X' = R X R^T), then shift the origin by r = t_i */
    double Rt[9], Ar[9], Br[9], Dr[9], rx[9], T[9];
    aba_transpose(R[i], Rt);
    aba_mm(R[i], Aa, T); aba_mm(T, Rt, Ar);
    aba_mm(R[i], Ba, T); aba_mm(T, Rt, Br);
    aba_mm(R[i], Da, T); aba_mm(T, Rt, Dr);
    aba_skew(c->t[i], rx);
    double Bp[9], rxD[9], BrT[9], rxBt[9], Bprx[9];
    aba_mm(rx, Dr, rxD);
    for (int k = 0; k < 9; ++k) Bp[k] = Br[k] + rxD[k];
    aba_transpose(Br, BrT);
    aba_mm(rx, BrT, rxBt);
    aba_mm(Bp, rx, Bprx);
    for (int k = 0; k < 9; ++k) {
      A[i - 1][k] += Ar[k] + rxBt[k] - Bprx[k];
      B[i - 1][k] += Bp[k];
      D[i - 1][k] += Dr[k];
    }
    double fpar[3], npar[3], rxf[3];
    aba_mv(R[i], paf, fpar);
    aba_mv(R[i], pan, npar);
    aba_cross(c->t[i], fpar, rxf);
    for (int k = 0; k < 3; ++k) { pf[i - 1][k] += fpar[k]; pn[i - 1][k] += npar[k] + rxf[k]; }
  }
  /* sweep 3: accelerations, base to tip */
  double aw[3] = {0, 0, 0}, av[3];
  {
    double mg[3] = {-c->gravity[0], -c->gravity[1], -c->gravity[2]};
    aba_mtv(c->Rb, mg, av);
  }
  for (int i = 0; i < ABA_NJ; ++i) {
    double axr[3], tmp[3], w2[3], v2[3];
    aba_cross(aw, c->t[i], axr);
    for (int k = 0; k < 3; ++k) tmp[k] = av[k] + axr[k];
    aba_mtv(R[i], aw, w2);
    aba_mtv(R[i], tmp, v2);
    for (int k = 0; k < 3; ++k) { w2[k] += cw[i][k]; v2[k] += cv[i][k]; }
    double Ua = U[i][0] * w2[0] + U[i][1] * w2[1] + U[i][2] * w2[2] + U[i][3] * v2[0] + U[i][4] * v2[1] + U[i][5] * v2[2];
    qdd[i] = (u[i] - Ua) * dinv[i];
    w2[2] += qdd[i];
    memcpy(aw, w2, sizeof(aw)); memcpy(av, v2, sizeof(av));
  }
}

/* One torque-mode integration step: effort clip, joint damping, ABA, semi-implicit Euler,
 * velocity clip, joint-limit clamp with the velocity zeroed on the active side. */
static inline void aba_torque_step(const AbaChain* c, const double effort[ABA_NJ], const double maxvel[ABA_NJ],
                                   const double damping[ABA_NJ], const double lower[ABA_NJ], const double upper[ABA_NJ],
                                   double dt, const float* torque_cmd, double q[ABA_NJ], double qd[ABA_NJ]) {
  double tau[ABA_NJ], qdd[ABA_NJ];
  for (int j = 0; j < ABA_NJ; ++j) {
    double t = (double)torque_cmd[j];
    t = t < -effort[j] ? -effort[j] : (t > effort[j] ? effort[j] : t);
    tau[j] = t - damping[j] * qd[j];
  }
  aba_forward_dynamics(c, q, qd, tau, qdd);
  for (int j = 0; j < ABA_NJ; ++j) {
    qd[j] += qdd[j] * dt;
    qd[j] = qd[j] < -maxvel[j] ? -maxvel[j] : (qd[j] > maxvel[j] ? maxvel[j] : qd[j]);
    q[j] += qd[j] * dt;
    if (q[j] < lower[j]) { q[j] = lower[j]; if (qd[j] < 0.0) qd[j] = 0.0; }
    if (q[j] > upper[j]) { q[j] = upper[j]; if (qd[j] > 0.0) qd[j] = 0.0; }
  }
}
#endif
