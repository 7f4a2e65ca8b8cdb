/* cube_model.h -- TEST INFRASTRUCTURE (oracle side, fp64) of the cube / pusher / gripper model.
 *
 * What p.stepSimulation() contributes to the push / pick envs (rl_push_env.py:242,349; rl_pick_env.py:242,348,417)
 * lives inside pybullet==3.0.6 (btMultiBodyDynamicsWorld: convex-mesh collision + 50-iteration PGS) and CANNOT be
 * reproduced offline (no Bullet source, no kuka/table meshes: SURVEY 8c, Appendix C) -- PARITY UNPINNED.  This file
 * states the behavioural model both the oracle and the CUDA kernels implement instead; the CUDA side
 * (csrc/cube_model.cuh) must match THIS to fp32 tolerance, and the model is pinned only by the reference's
 * known-answer statistics (untouched-cube push return in [-514.6, -511]).
 *
 * Model, per sim step (Bullet defaults: dt = 1/240 s, gravity (0,0,-10), ERP 0.2, restitution 0, damping 0.04):
 *   cube     : free rigid box, side 0.04 (models/cube_small_push.urdf:17,25), mass 1 (:12), inertia recomputed from
 *              the box as Bullet does by default (m a^2 / 6 = 2.667e-4, isotropic), lateral friction 5.0 (:5)
 *              combined multiplicatively with the default 0.5 of the table / arm links -> mu = 2.5.
 *   table    : plane z = -0.025 (pybullet_data table/table.urdf at (0.5,0,-0.65): top box 0.05 thick centred 0.6 up).
 *   arm      : kinematic, zero velocity (teleported); represented by sphere proxies fixed in the EE link frame
 *              (push: one sphere r 0.045 at EE + 0.02 z_ee, the link-7 flange; pick: palm + finger-tip spheres).
 *              The cube is moved by penetration recovery against those static spheres, as in Bullet.
 *   contacts : 8 box corners vs plane (speculative margin 5 mm) + closest-point box/sphere contacts; each contact
 *              = 1 normal row + 2 friction rows (box friction cone), solved by 10 projected-Gauss-Seidel sweeps;
 *              normal bias = ERP * depth / dt when penetrating, -gap / dt inside the speculative margin.
 *   integrate: semi-implicit Euler; quaternion integrated with the exponential map.
 *   pick     : fingers close (latched) when any proxy is within 6 mm of the cube (rl_pick_env.py:412-416); if the
 *              cube centre is then within 3 cm of the grasp point (EE + 0.257 z_ee, :79,374) the cube is held:
 *              it follows the grasp point kinematically (documented simplification of friction grasping).
 */
#ifndef ORACLE_CUBE_MODEL_H
#define ORACLE_CUBE_MODEL_H
#include <math.h>
#include <string.h>

#define CUBE_DT (1.0 / 240.0)
#define CUBE_G 10.0
#define CUBE_HALF 0.02
#define CUBE_MASS 1.0
#define CUBE_INERTIA (CUBE_MASS * (0.04 * 0.04) / 6.0)
#define CUBE_MU 2.5
#define CUBE_ERP 0.2
#define CUBE_TABLE_Z (-0.025)
#define CUBE_MARGIN 0.005
#define CUBE_PGS_ITERS 10
#define CUBE_DAMP_FACTOR 0.99982992284  /* (1 - 0.04)^(1/240) */
#define CUBE_MAX_CONTACTS 11
#define PUSH_R 0.045
#define PUSH_OFF 0.02
#define PICK_PALM_R 0.05
#define PICK_PALM_OFF 0.12
#define PICK_TIP_R 0.012
#define PICK_TIP_OPEN 0.045
#define PICK_TIP_CLOSED_R 0.02
#define PICK_GRIPPER_LEN 0.257
#define PICK_CLOSE_DIST 0.006
#define PICK_HOLD_DIST 0.03

typedef struct CubeState {
  double pos[3], quat[4], v[3], w[3];
} CubeState;

typedef struct CubeContact {
  double r[3];      /* contact point relative to the cube centre (world axes) */
  double n[3];      /* unit normal pointing INTO the cube (direction the impulse pushes the cube) */
  double t1[3], t2[3];
  double bias;      /* target normal velocity */
  double ln, l1, l2;/* accumulated impulses */
} CubeContact;

static inline void cube_init(CubeState* c, double x, double y, double z, double yaw) {
  memset(c, 0, sizeof(*c));
  c->pos[0] = x; c->pos[1] = y; c->pos[2] = z;
  c->quat[2] = sin(0.5 * yaw);  /* getQuaternionFromEuler([0,0,ang]) rl_push_env.py:201 */
  c->quat[3] = cos(0.5 * yaw);
}

static inline void cube_rot(const double q[4], double R[9]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

static inline void cube_cross(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

static inline void cube_tangents(CubeContact* k) {
  /* btPlaneSpace1 */
  const double* n = k->n;
  if (fabs(n[2]) > 0.7071067811865475244) {
    double a = n[1] * n[1] + n[2] * n[2], s = 1.0 / sqrt(a);
    k->t1[0] = 0; k->t1[1] = -n[2] * s; k->t1[2] = n[1] * s;
    k->t2[0] = a * s; k->t2[1] = -n[0] * k->t1[2]; k->t2[2] = n[0] * k->t1[1];
  } else {
    double a = n[0] * n[0] + n[1] * n[1], s = 1.0 / sqrt(a);
    k->t1[0] = -n[1] * s; k->t1[1] = n[0] * s; k->t1[2] = 0;
    k->t2[0] = -n[2] * k->t1[1]; k->t2[1] = n[2] * k->t1[0]; k->t2[2] = a * s;
  }
}

/* signed distance from a sphere (centre c, radius rad) to the box, with the closest point on the box (world,
 * relative to the cube centre) and the unit direction from the sphere towards the cube */
static inline double cube_sphere_query(const CubeState* cb, const double R[9], const double c[3], double rad,
                                       double rrel[3], double n[3]) {
  double d[3] = {c[0] - cb->pos[0], c[1] - cb->pos[1], c[2] - cb->pos[2]};
  double l[3], cl[3];
  for (int i = 0; i < 3; ++i) l[i] = R[i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2];  /* R^T d */
  int inside = 1;
  for (int i = 0; i < 3; ++i) {
    cl[i] = l[i] < -CUBE_HALF ? -CUBE_HALF : (l[i] > CUBE_HALF ? CUBE_HALF : l[i]);
    if (cl[i] != l[i]) inside = 0;
  }
  double nl[3], dist;
  if (!inside) {
    double e[3] = {l[0] - cl[0], l[1] - cl[1], l[2] - cl[2]};
    dist = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    for (int i = 0; i < 3; ++i) nl[i] = -e[i] / dist;    /* from sphere centre towards the box */
  } else {
    /* centre inside the box: leave through the nearest face */
    int ax = 0;
    double best = CUBE_HALF - fabs(l[0]);
    for (int i = 1; i < 3; ++i) {
      double m = CUBE_HALF - fabs(l[i]);
      if (m < best) { best = m; ax = i; }
    }
    double sgn = l[ax] >= 0 ? 1.0 : -1.0;
    cl[ax] = sgn * CUBE_HALF;
    dist = -best;
    nl[0] = nl[1] = nl[2] = 0.0;
    nl[ax] = -sgn;
  }
  for (int i = 0; i < 3; ++i) {
    rrel[i] = R[3 * i] * cl[0] + R[3 * i + 1] * cl[1] + R[3 * i + 2] * cl[2];
    n[i] = R[3 * i] * nl[0] + R[3 * i + 1] * nl[1] + R[3 * i + 2] * nl[2];
  }
  return dist - rad;
}

/* sphere proxies of the arm for this step: centres (world) and radii; returns the count */
static inline int cube_arm_proxies(const double ee[3], const double Ree[9], int pick, double grip, double C[3][3], double rad[3]) {
  const double zx = Ree[2], zy = Ree[5], zz = Ree[8];   /* EE local +z in world */
  const double xx = Ree[0], xy = Ree[3], xz = Ree[6];   /* EE local +x in world */
  if (!pick) {
    C[0][0] = ee[0] + PUSH_OFF * zx; C[0][1] = ee[1] + PUSH_OFF * zy; C[0][2] = ee[2] + PUSH_OFF * zz;
    rad[0] = PUSH_R;
    return 1;
  }
  C[0][0] = ee[0] + PICK_PALM_OFF * zx; C[0][1] = ee[1] + PICK_PALM_OFF * zy; C[0][2] = ee[2] + PICK_PALM_OFF * zz;
  rad[0] = PICK_PALM_R;
  double g[3] = {ee[0] + PICK_GRIPPER_LEN * zx, ee[1] + PICK_GRIPPER_LEN * zy, ee[2] + PICK_GRIPPER_LEN * zz};
  if (grip >= 0.5) {
    C[1][0] = g[0]; C[1][1] = g[1]; C[1][2] = g[2];
    rad[1] = PICK_TIP_CLOSED_R;
    return 2;
  }
  for (int s = 0; s < 2; ++s) {
    double o = s ? -PICK_TIP_OPEN : PICK_TIP_OPEN;
    C[1 + s][0] = g[0] + o * xx; C[1 + s][1] = g[1] + o * xy; C[1 + s][2] = g[2] + o * xz;
    rad[1 + s] = PICK_TIP_R;
  }
  return 3;
}

/* getClosestPoints(kuka, cube, 0.006) stand-in (rl_pick_env.py:412): min signed distance proxy <-> cube */
static inline double cube_gripper_distance(const CubeState* cb, const double ee[3], const double Ree[9]) {
  double R[9], C[3][3], rad[3], rr[3], nn[3];
  cube_rot(cb->quat, R);
  int np = cube_arm_proxies(ee, Ree, 1, 0.0, C, rad);
  double best = 1e30;
  for (int i = 0; i < np; ++i) {
    double d = cube_sphere_query(cb, R, C[i], rad[i], rr, nn);
    if (d < best) best = d;
  }
  return best;
}

static inline void cube_row(CubeState* cb, const double r[3], const double dir[3], double target, double lo, double hi,
                            double* acc) {
  double rxd[3];
  cube_cross(r, dir, rxd);
  double vrel = dir[0] * cb->v[0] + dir[1] * cb->v[1] + dir[2] * cb->v[2] + rxd[0] * cb->w[0] + rxd[1] * cb->w[1] + rxd[2] * cb->w[2];
  double k = 1.0 / CUBE_MASS + (rxd[0] * rxd[0] + rxd[1] * rxd[1] + rxd[2] * rxd[2]) / CUBE_INERTIA;
  double dl = (target - vrel) / k;
  double nl = *acc + dl;
  nl = nl < lo ? lo : (nl > hi ? hi : nl);
  dl = nl - *acc;
  *acc = nl;
  for (int i = 0; i < 3; ++i) {
    cb->v[i] += dl * dir[i] / CUBE_MASS;
    cb->w[i] += dl * rxd[i] / CUBE_INERTIA;
  }
}

/* one p.stepSimulation() for the cube.  grip: 0 open / push, 1 closed, 2 holding */
static inline void cube_step(CubeState* cb, const double ee[3], const double Ree[9], int pick, double grip) {
  if (pick && grip >= 1.5) {  /* held: follows the grasp point */
    cb->pos[0] = ee[0] + PICK_GRIPPER_LEN * Ree[2];
    cb->pos[1] = ee[1] + PICK_GRIPPER_LEN * Ree[5];
    cb->pos[2] = ee[2] + PICK_GRIPPER_LEN * Ree[8];
    for (int i = 0; i < 3; ++i) cb->v[i] = cb->w[i] = 0.0;
    return;
  }
  /* predictUnconstraintMotion: gravity, then damping */
  cb->v[2] -= CUBE_G * CUBE_DT;
  for (int i = 0; i < 3; ++i) { cb->v[i] *= CUBE_DAMP_FACTOR; cb->w[i] *= CUBE_DAMP_FACTOR; }

  double R[9];
  cube_rot(cb->quat, R);
  CubeContact K[CUBE_MAX_CONTACTS];
  int nk = 0;
  /* 8 corners vs the table plane */
  for (int c = 0; c < 8; ++c) {
    double l[3] = {(c & 1) ? CUBE_HALF : -CUBE_HALF, (c & 2) ? CUBE_HALF : -CUBE_HALF, (c & 4) ? CUBE_HALF : -CUBE_HALF};
    double r[3];
    for (int i = 0; i < 3; ++i) r[i] = R[3 * i] * l[0] + R[3 * i + 1] * l[1] + R[3 * i + 2] * l[2];
    double gap = cb->pos[2] + r[2] - CUBE_TABLE_Z;
    if (gap < CUBE_MARGIN) {
      CubeContact* k = &K[nk++];
      memcpy(k->r, r, sizeof(r));
      k->n[0] = 0; k->n[1] = 0; k->n[2] = 1;
      k->bias = gap < 0 ? -CUBE_ERP * gap / CUBE_DT : -gap / CUBE_DT;
      k->ln = k->l1 = k->l2 = 0;
      cube_tangents(k);
    }
  }
  /* arm proxies */
  double C[3][3], rad[3];
  int np = cube_arm_proxies(ee, Ree, pick, grip, C, rad);
  for (int p = 0; p < np; ++p) {
    double rr[3], nn[3];
    double d = cube_sphere_query(cb, R, C[p], rad[p], rr, nn);
    if (d < 0) {
      CubeContact* k = &K[nk++];
      memcpy(k->r, rr, sizeof(rr));
      memcpy(k->n, nn, sizeof(nn));
      k->bias = -CUBE_ERP * d / CUBE_DT;
      k->ln = k->l1 = k->l2 = 0;
      cube_tangents(k);
    }
  }
  for (int it = 0; it < CUBE_PGS_ITERS; ++it) {
    for (int i = 0; i < nk; ++i) {
      CubeContact* k = &K[i];
      cube_row(cb, k->r, k->n, k->bias, 0.0, 1e30, &k->ln);
      double lim = CUBE_MU * k->ln;
      cube_row(cb, k->r, k->t1, 0.0, -lim, lim, &k->l1);
      cube_row(cb, k->r, k->t2, 0.0, -lim, lim, &k->l2);
    }
  }
  /* integrateTransforms */
  for (int i = 0; i < 3; ++i) cb->pos[i] += cb->v[i] * CUBE_DT;
  double wn = sqrt(cb->w[0] * cb->w[0] + cb->w[1] * cb->w[1] + cb->w[2] * cb->w[2]);
  double ang = wn * CUBE_DT;
  double s = ang > 1e-6 ? sin(0.5 * ang) / wn : 0.5 * CUBE_DT * (1.0 - ang * ang / 24.0);
  double dq[4] = {cb->w[0] * s, cb->w[1] * s, cb->w[2] * s, cos(0.5 * ang)};
  const double* q = cb->quat;
  double nq[4] = {dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1],
                  dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2],
                  dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0],
                  dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2]};
  double inv = 1.0 / sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
  for (int i = 0; i < 4; ++i) cb->quat[i] = nq[i] * inv;
}
#endif
