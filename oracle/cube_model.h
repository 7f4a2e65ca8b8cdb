/* cube_model.h -- TEST INFRASTRUCTURE (oracle side, fp64) of the cube / pusher / gripper model.
 *
 * What p.stepSimulation() contributes to the push / pick envs (rl_push_env.py:242,349; rl_pick_env.py:242,348,417)
 * lives inside pybullet==3.0.6 (btMultiBodyDynamicsWorld: convex-mesh collision + PGS contact solver) and CANNOT be
 * reproduced offline (no Bullet source, no kuka / gripper / table meshes: SURVEY 8c, Appendix C) -- PARITY UNPINNED.
 * This file states the model both the oracle and the CUDA kernels implement instead; the CUDA side
 * (csrc/cube_model.cuh) must match THIS to fp32 tolerance.  System-level pins: the reference's untouched-cube return
 * (-504, visdata/push/updata_TD3) and its push learning curve reproduced by the reference's OWN learner on this model
 * (tests/system/ref_learner_on_oracle.py, tests/golden/ref_learner_{reach,push,pick}_*.json, tests/test_system_pins.py).
 *
 * Model, per sim step, with Bullet's / pybullet's default parameters:
 *   dt 1/240 s, gravity (0,0,-10), ERP 0.2 applied as a velocity bias (btMultiBody contacts have no split impulse),
 *   restitution 0, linear / angular damping 0.04, at most 50 projected-Gauss-Seidel iterations with pybullet's early
 *   exit once the largest squared velocity residual of a sweep drops to 1e-7 (m_leastSquaresResidualThreshold).
 *   cube     : free rigid box, side 0.04 (models/cube_small_push.urdf:17,25), mass 1 (:12), inertia recomputed from
 *              the box as Bullet does by default (m a^2 / 6 = 2.667e-4, isotropic), lateral friction 5.0 (:5)
 *              combined multiplicatively with the default 0.5 of the table / arm links -> mu = 2.5.
 *   table    : top face z = -0.025 over x in [-0.25, 1.25], y in [-0.5, 0.5] (pybullet_data table/table.urdf at
 *              (0.5,0,-0.65): 1.5 x 1 x 0.05 top centred 0.6 up); beyond the edge the ground plane z = -0.65
 *              (plane.urdf at z = -0.65, rl_push_env.py:184).
 *   arm      : kinematic with zero velocity (teleported by resetJointState, held by the default motors), seen by the
 *              cube as CAPSULES fixed in the EE link frame (axis = EE local z, which the IK keeps pointing down):
 *                push : link 7 / flange + the link-6 body above it: radius 0.04, axis from 0.10 behind the EE frame
 *                       to 0.005 in front of it (round end = the flange face 0.045 in front of the frame).
 *                pick : palm (radius 0.045, EE frame to 0.15 in front) and the two WSG50 fingers (radius 0.01, from
 *                       0.15 to 0.247 in front; tip centre 0.05 off the axis when open, 0.025 when closed, along the
 *                       EE local x axis): fingertips reach gripper_length = 0.257 (rl_pick_env.py:79).
 *              Contact of a capsule with the box = contact of the sphere of the capsule's radius centred at the axis
 *              point nearest the cube centre (box / sphere closest point, or minimum-translation face when the centre
 *              is inside the box).  A tall capsule therefore pushes a cube lying on the table horizontally, as the
 *              side of the real flange does, and cannot get under it.  The cube is moved by penetration recovery
 *              against those static shapes, as in Bullet.
 *   contacts : box corners vs the supporting plane (speculative margin 5 mm; the first 4 in corner order -- a face --
 *              which is every corner a rigid box can have that close to a plane) + the capsule contacts; each contact
 *              = 1 normal row + 2 friction rows (btPlaneSpace1 directions, box friction limits mu * normal impulse);
 *              normal bias = ERP * depth / dt when penetrating, -gap / dt inside the speculative margin.
 *   integrate: semi-implicit Euler; quaternion integrated with the exponential map.
 *   pick     : the fingers close for good once any capsule is within 6 mm of the cube (getClosestPoints(kuka, cube,
 *              0.006), rl_pick_env.py:412-416).  Closed fingers squeeze the cube through ordinary contacts (normal +
 *              friction rows against ZERO-velocity fingers): there is no kinematic attachment.  As in Bullet, a cube
 *              held by friction does not follow teleported fingers; it moves only through penetration recovery.
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
#define CUBE_GROUND_Z (-0.65)
#define CUBE_TABLE_X0 (-0.25)
#define CUBE_TABLE_X1 (1.25)
#define CUBE_TABLE_Y0 (-0.5)
#define CUBE_TABLE_Y1 (0.5)
#define CUBE_MARGIN 0.005
#define CUBE_PGS_ITERS 50
#define CUBE_PGS_RESIDUAL 1e-7
#define CUBE_DAMP_FACTOR 0.99982992284  /* (1 - 0.04)^(1/240) */
#define CUBE_MAX_PROXIES 3
#define CUBE_MAX_CORNERS 4   /* a face: no more than 4 corners of a rigid 4 cm box can be within 5 mm of a plane */
#define CUBE_MAX_CONTACTS (CUBE_MAX_CORNERS + CUBE_MAX_PROXIES)
#define PUSH_R 0.04
#define PUSH_A0 (-0.10)
#define PUSH_A1 0.005
#define PICK_PALM_R 0.045
#define PICK_PALM_A0 0.0
#define PICK_PALM_A1 0.15
#define PICK_FINGER_R 0.01
#define PICK_FINGER_A0 0.15
#define PICK_FINGER_A1 0.247
#define PICK_FINGER_BASE 0.03
#define PICK_TIP_OPEN 0.05
#define PICK_TIP_CLOSED 0.025
#define PICK_GRIPPER_LEN 0.257
#define PICK_CLOSE_DIST 0.006

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

/* capsule of the arm: segment a-b (world) swept by a sphere of radius rad */
typedef struct CubeCapsule {
  double a[3], b[3], rad;
} CubeCapsule;

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

/* capsule <-> box: the sphere of the capsule's radius at the axis point nearest the cube centre */
static inline double cube_capsule_query(const CubeState* cb, const double R[9], const CubeCapsule* k, double rrel[3], double n[3]) {
  double ab[3] = {k->b[0] - k->a[0], k->b[1] - k->a[1], k->b[2] - k->a[2]};
  double ac[3] = {cb->pos[0] - k->a[0], cb->pos[1] - k->a[1], cb->pos[2] - k->a[2]};
  double len2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  double t = (ab[0] * ac[0] + ab[1] * ac[1] + ab[2] * ac[2]) / len2;
  t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  double c[3] = {k->a[0] + t * ab[0], k->a[1] + t * ab[1], k->a[2] + t * ab[2]};
  return cube_sphere_query(cb, R, c, k->rad, rrel, n);
}

static inline void cube_axis_point(const double ee[3], const double Ree[9], double along, double side, double out[3]) {
  /* EE local z = column 2 of Ree, local x = column 0 */
  out[0] = ee[0] + along * Ree[2] + side * Ree[0];
  out[1] = ee[1] + along * Ree[5] + side * Ree[3];
  out[2] = ee[2] + along * Ree[8] + side * Ree[6];
}

/* capsules of the arm for this step; grip: 0 open / push, >= 0.5 fingers closed.  Returns the count. */
static inline int cube_arm_capsules(const double ee[3], const double Ree[9], int pick, double grip, CubeCapsule K[CUBE_MAX_PROXIES]) {
  if (!pick) {
    cube_axis_point(ee, Ree, PUSH_A0, 0.0, K[0].a);
    cube_axis_point(ee, Ree, PUSH_A1, 0.0, K[0].b);
    K[0].rad = PUSH_R;
    return 1;
  }
  cube_axis_point(ee, Ree, PICK_PALM_A0, 0.0, K[0].a);
  cube_axis_point(ee, Ree, PICK_PALM_A1, 0.0, K[0].b);
  K[0].rad = PICK_PALM_R;
  const double tip = grip >= 0.5 ? PICK_TIP_CLOSED : PICK_TIP_OPEN;
  for (int s = 0; s < 2; ++s) {
    const double sg = s ? -1.0 : 1.0;
    cube_axis_point(ee, Ree, PICK_FINGER_A0, sg * PICK_FINGER_BASE, K[1 + s].a);
    cube_axis_point(ee, Ree, PICK_FINGER_A1, sg * tip, K[1 + s].b);
    K[1 + s].rad = PICK_FINGER_R;
  }
  return 3;
}

/* getClosestPoints(kuka, cube, 0.006) stand-in (rl_pick_env.py:412): min signed distance capsule <-> cube, fingers
 * in their current state */
static inline double cube_gripper_distance(const CubeState* cb, const double ee[3], const double Ree[9], double grip) {
  double R[9], rr[3], nn[3];
  CubeCapsule K[CUBE_MAX_PROXIES];
  cube_rot(cb->quat, R);
  int np = cube_arm_capsules(ee, Ree, 1, grip, K);
  double best = 1e30;
  for (int i = 0; i < np; ++i) {
    double d = cube_capsule_query(cb, R, &K[i], rr, nn);
    if (d < best) best = d;
  }
  return best;
}

/* one PGS row; returns the squared velocity change along the row (Bullet's least-squares residual term) */
static inline double cube_row(CubeState* cb, const double r[3], const double dir[3], double target, double lo, double hi,
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
  double dv = dl * k;
  return dv * dv;
}

/* one p.stepSimulation() for the cube.  grip: 0 open / push, >= 0.5 fingers closed.  Returns the PGS sweeps used. */
static inline int cube_step(CubeState* cb, const double ee[3], const double Ree[9], int pick, double grip) {
  /* predictUnconstraintMotion: gravity, then damping */
  cb->v[2] -= CUBE_G * CUBE_DT;
  for (int i = 0; i < 3; ++i) { cb->v[i] *= CUBE_DAMP_FACTOR; cb->w[i] *= CUBE_DAMP_FACTOR; }

  double R[9];
  cube_rot(cb->quat, R);
  CubeContact K[CUBE_MAX_CONTACTS];
  int nk = 0;
  /* 8 corners vs the supporting plane: the table top while the cube centre is over it, else the ground */
  const int on_table = cb->pos[0] >= CUBE_TABLE_X0 && cb->pos[0] <= CUBE_TABLE_X1 && cb->pos[1] >= CUBE_TABLE_Y0 && cb->pos[1] <= CUBE_TABLE_Y1;
  const double plane_z = on_table ? CUBE_TABLE_Z : CUBE_GROUND_Z;
  for (int c = 0; c < 8; ++c) {
    double l[3] = {(c & 1) ? CUBE_HALF : -CUBE_HALF, (c & 2) ? CUBE_HALF : -CUBE_HALF, (c & 4) ? CUBE_HALF : -CUBE_HALF};
    double r[3];
    for (int i = 0; i < 3; ++i) r[i] = R[3 * i] * l[0] + R[3 * i + 1] * l[1] + R[3 * i + 2] * l[2];
    double gap = cb->pos[2] + r[2] - plane_z;
    if (gap < CUBE_MARGIN && nk < CUBE_MAX_CORNERS) {
      CubeContact* k = &K[nk++];
      memcpy(k->r, r, sizeof(r));
      k->n[0] = 0; k->n[1] = 0; k->n[2] = 1;
      k->bias = gap < 0 ? -CUBE_ERP * gap / CUBE_DT : -gap / CUBE_DT;
      k->ln = k->l1 = k->l2 = 0;
      cube_tangents(k);
    }
  }
  /* arm capsules */
  CubeCapsule A[CUBE_MAX_PROXIES];
  int np = cube_arm_capsules(ee, Ree, pick, grip, A);
  for (int p = 0; p < np; ++p) {
    double rr[3], nn[3];
    double d = cube_capsule_query(cb, R, &A[p], rr, nn);
    if (d < 0) {
      CubeContact* k = &K[nk++];
      memcpy(k->r, rr, sizeof(rr));
      memcpy(k->n, nn, sizeof(nn));
      k->bias = -CUBE_ERP * d / CUBE_DT;
      k->ln = k->l1 = k->l2 = 0;
      cube_tangents(k);
    }
  }
  int sweeps = 0;
  if (nk > 0) {
    for (int it = 0; it < CUBE_PGS_ITERS; ++it) {
      double res = 0.0;
      for (int i = 0; i < nk; ++i) {
        CubeContact* k = &K[i];
        double r0 = cube_row(cb, k->r, k->n, k->bias, 0.0, 1e30, &k->ln);
        double lim = CUBE_MU * k->ln;
        double r1 = cube_row(cb, k->r, k->t1, 0.0, -lim, lim, &k->l1);
        double r2 = cube_row(cb, k->r, k->t2, 0.0, -lim, lim, &k->l2);
        res = fmax(res, fmax(r0, fmax(r1, r2)));
      }
      ++sweeps;
      if (res <= CUBE_PGS_RESIDUAL) break;   /* m_leastSquaresResidualThreshold */
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
  return sweeps;
}
#endif
