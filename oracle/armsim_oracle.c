/* armsim_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see armsim_oracle.h for the parity statement).
 *
 * Plain-C fp64 restatement of the reference hot path.  Each function cites the reference lines it follows.
 * The arithmetic inside pybullet==3.0.6 (absent from /root/reference) is restated from Bullet's published
 * algorithm: examples/SharedMemory/PhysicsServerCommandProcessor.cpp (processCalculateInverseKinematicsCommand),
 * examples/SharedMemory/IKTrajectoryHelper.cpp (computeIK), examples/ThirdPartyLibs/BussIK/Jacobian.cpp
 * (CalcDeltaThetasDLS2) -- PARITY UNPINNED, see header.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  -ffp-contract=off matters: the only fused
 * multiply-adds are the explicit fmaf() calls in the reset sampler, which the device mirrors with __fmaf_rn so
 * that sampled goals / cube poses are bit-identical.
 */
#include "armsim_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "armsim_defaults.h"
#include "armsim_robot_models.h"
#include "cube_model.h"
#include "aba_model.h"

#define NJ ARMSIM_NJ

/* ------------------------------------------------------------------------------------------------ Philox */
static inline void mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(0xD2511F53u, c0, &hi0, &lo0);
    mulhilo32(0xCD9E8D57u, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_reset_uniforms(uint64_t seed, uint64_t env_gid, uint32_t episode, uint32_t block, float u[4]) {
  uint32_t ctr[4] = {(uint32_t)env_gid, (uint32_t)(env_gid >> 32), episode, block};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  orc_philox4x32_10(ctr, key, r);
  for (int i = 0; i < 4; ++i) u[i] = (float)(r[i] >> 8) * 5.9604644775390625e-08f; /* 2^-24, exact in f32 */
}

/* ------------------------------------------------------------------------------------------------ chain */
typedef struct OrcChain {
  double Rb[9], tb[3];          /* base */
  double Rf[NJ][9], t[NJ][3];   /* fixed joint-origin transform per joint */
  double lower[NJ], upper[NJ];
  double effort[NJ], velocity[NJ], damping[NJ];   /* torque mode */
  AbaChain dyn;
} OrcChain;

static void rpy_to_mat(const double rpy[3], double R[9]) {
  /* URDF: R = Rz(yaw) Ry(pitch) Rx(roll) */
  double cr = cos(rpy[0]), sr = sin(rpy[0]), cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

static void mat_mul(const double A[9], const double B[9], double C[9]) {
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof(T));
}

static void chain_from_model(const ArmsimRobotModel* m, OrcChain* c) {
  rpy_to_mat(m->base_rpy, c->Rb);
  memcpy(c->tb, m->base_xyz, sizeof(c->tb));
  memcpy(c->dyn.Rb, c->Rb, sizeof(c->Rb));
  for (int j = 0; j < NJ; ++j) {
    rpy_to_mat(m->rpy[j], c->Rf[j]);
    memcpy(c->t[j], m->xyz[j], sizeof(c->t[j]));
    c->lower[j] = m->lower[j];
    c->upper[j] = m->upper[j];
    c->effort[j] = m->effort[j];
    c->velocity[j] = m->velocity[j];
    c->damping[j] = m->damping[j];
    memcpy(c->dyn.Rf[j], c->Rf[j], sizeof(c->Rf[j]));
    memcpy(c->dyn.t[j], c->t[j], sizeof(c->t[j]));
    c->dyn.mass[j] = m->mass[j];
    memcpy(c->dyn.com[j], m->com[j], sizeof(c->dyn.com[j]));
    memcpy(c->dyn.Ic[j], m->inertia[j], sizeof(c->dyn.Ic[j]));
  }
  c->dyn.gravity[0] = 0.0; c->dyn.gravity[1] = 0.0; c->dyn.gravity[2] = -10.0;
}

static void chain_from_custom(const ArmsimChain* m, OrcChain* c) {
  rpy_to_mat(m->base_rpy, c->Rb);
  memcpy(c->tb, m->base_xyz, sizeof(c->tb));
  memcpy(c->dyn.Rb, c->Rb, sizeof(c->Rb));
  for (int j = 0; j < NJ; ++j) {
    rpy_to_mat(m->rpy[j], c->Rf[j]);
    memcpy(c->t[j], m->xyz[j], sizeof(c->t[j]));
    c->lower[j] = m->lower[j];
    c->upper[j] = m->upper[j];
    c->effort[j] = m->effort[j];
    c->velocity[j] = m->velocity[j];
    c->damping[j] = m->damping[j];
    memcpy(c->dyn.Rf[j], c->Rf[j], sizeof(c->Rf[j]));
    memcpy(c->dyn.t[j], c->t[j], sizeof(c->t[j]));
    c->dyn.mass[j] = m->mass[j];
    memcpy(c->dyn.com[j], m->com[j], sizeof(c->dyn.com[j]));
    memcpy(c->dyn.Ic[j], m->inertia[j], sizeof(c->dyn.Ic[j]));
  }
  c->dyn.gravity[0] = 0.0; c->dyn.gravity[1] = 0.0; c->dyn.gravity[2] = -10.0;
}

static int chain_builtin(int32_t robot, OrcChain* c) {
  if (robot == ARMSIM_ROBOT_KUKA_IIWA) chain_from_model(&ARMSIM_MODEL_KUKA_IIWA, c);
  else if (robot == ARMSIM_ROBOT_DIANA_S1) chain_from_model(&ARMSIM_MODEL_DIANA_S1, c);
  else return -1;
  return 0;
}

/* FK: what p.getLinkState(kuka, 6)[4] / [5] report (worldLinkFramePosition / Orientation of the URDF link frame of
 * link 7), rl_reach_env.py:237,271.  Link transform = T(xyz, rpy) * Rz(q). */
static void chain_fk(const OrcChain* c, const double q[NJ], double pee[3], double Ree[9], double P[NJ][3], double Z[NJ][3]) {
  double R[9], p[3];
  memcpy(R, c->Rb, sizeof(R));
  memcpy(p, c->tb, sizeof(p));
  for (int j = 0; j < NJ; ++j) {
    for (int i = 0; i < 3; ++i) p[i] += R[3 * i] * c->t[j][0] + R[3 * i + 1] * c->t[j][1] + R[3 * i + 2] * c->t[j][2];
    mat_mul(R, c->Rf[j], R);
    if (P) memcpy(P[j], p, sizeof(p));
    if (Z) { Z[j][0] = R[2]; Z[j][1] = R[5]; Z[j][2] = R[8]; }
    double cq = cos(q[j]), sq = sin(q[j]);
    double Rz[9] = {cq, -sq, 0, sq, cq, 0, 0, 0, 1};
    mat_mul(R, Rz, R);
  }
  if (pee) memcpy(pee, p, sizeof(p));
  if (Ree) memcpy(Ree, R, sizeof(R));
}

static void chain_jacobian(const double pee[3], double P[NJ][3], double Z[NJ][3], double J[6][NJ]) {
  for (int j = 0; j < NJ; ++j) {
    double r[3] = {pee[0] - P[j][0], pee[1] - P[j][1], pee[2] - P[j][2]};
    J[0][j] = Z[j][1] * r[2] - Z[j][2] * r[1];
    J[1][j] = Z[j][2] * r[0] - Z[j][0] * r[2];
    J[2][j] = Z[j][0] * r[1] - Z[j][1] * r[0];
    J[3][j] = Z[j][0]; J[4][j] = Z[j][1]; J[5][j] = Z[j][2];
  }
}

/* btMatrix3x3::getRotation */
static void mat_to_quat(const double m[9], double q[4]) {
  double trace = m[0] + m[4] + m[8];
  if (trace > 0.0) {
    double s = sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (m[7] - m[5]) * s;
    q[1] = (m[2] - m[6]) * s;
    q[2] = (m[3] - m[1]) * s;
  } else {
    int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * s;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * s;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * s;
  }
}

/* p.getQuaternionFromEuler: btQuaternion::setEulerZYX(yaw, pitch, roll) */
void orc_quat_from_euler(const double rpy[3], double q[4]) {
  double hr = rpy[0] * 0.5, hp = rpy[1] * 0.5, hy = rpy[2] * 0.5;
  double cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy = sin(hy);
  q[0] = sr * cp * cy - cr * sp * sy;
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
  q[3] = cr * cp * cy + sr * sp * sy;
}

/* MatrixRmn::Solve: Gaussian elimination with partial pivoting on the augmented matrix, then back substitution. */
static void solve_gepp(int n, double* A /* n x n row-major, destroyed */, double* b /* destroyed */, double* x) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(A[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(A[r * n + c]) > best) { best = fabs(A[r * n + c]); piv = r; }
    if (piv != c) {
      for (int k = 0; k < n; ++k) { double tmp = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = tmp; }
      double tmp = b[c]; b[c] = b[piv]; b[piv] = tmp;
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

/* p.calculateInverseKinematics(body, 6, targetPosition, targetOrientation, jointDamping) as called at
 * rl_reach_env.py:244-250 (no null-space limits, no currentPositions, default maxNumIterations = 20 and
 * residualThreshold = 1e-4) -- SURVEY Appendix B:
 *   diff = +inf
 *   for it in 0..19 while diff > 1e-4:
 *       p, R, J = FK + geometric Jacobian of the EE LINK frame at q
 *       e  = [p* - p ; angle * axis of (q* (x) q^-1), angle wrapped to (-pi, pi]]   (IKTrajectoryHelper::computeIK;
 *            Bullet stores that angle in a `float`, reproduced here)
 *       dq = (J^T J + diag(jointDamping))^-1 J^T e                                   (Jacobian::CalcDeltaThetasDLS2)
 *       if max|dq| > 45 deg: dq *= 45deg / max|dq|
 *       q += dq ; diff = |p* - FK(q).p|
 *   joint limits are NOT enforced.  */
static int chain_ik(const OrcChain* c, const double q_in[NJ], const double tp[3], const double tq[4], double damping,
                    int max_iters, double residual, double q_out[NJ], double* final_diff) {
  double q[NJ];
  memcpy(q, q_in, sizeof(q));
  double diff = 1e30;
  int it = 0;
  const double max_angle = 45.0 * M_PI / 180.0; /* BussIK MaxAngleDLS */
  for (; it < max_iters && diff > residual; ++it) {
    double pee[3], Ree[9], P[NJ][3], Z[NJ][3], J[6][NJ];
    chain_fk(c, q, pee, Ree, P, Z);
    chain_jacobian(pee, P, Z, J);
    double e[6];
    for (int i = 0; i < 3; ++i) e[i] = tp[i] - pee[i];
    /* deltaQ = endQ * startQ.inverse() */
    double sq[4];
    mat_to_quat(Ree, sq);
    double ix = -sq[0], iy = -sq[1], iz = -sq[2], iw = sq[3];
    double dx = tq[3] * ix + tq[0] * iw + tq[1] * iz - tq[2] * iy;
    double dy = tq[3] * iy + tq[1] * iw + tq[2] * ix - tq[0] * iz;
    double dz = tq[3] * iz + tq[2] * iw + tq[0] * iy - tq[1] * ix;
    double dw = tq[3] * iw - tq[0] * ix - tq[1] * iy - tq[2] * iz;
    double wc = dw < -1.0 ? -1.0 : (dw > 1.0 ? 1.0 : dw);
    float angle = (float)(2.0 * acos(wc)); /* `float angle = deltaQ.getAngle();` */
    double s2 = 1.0 - dw * dw, ax[3];
    if (s2 < 10.0 * 2.220446049250313e-16) { ax[0] = 1.0; ax[1] = 0.0; ax[2] = 0.0; }
    else { double s = 1.0 / sqrt(s2); ax[0] = dx * s; ax[1] = dy * s; ax[2] = dz * s; }
    if (angle > (float)M_PI) angle -= (float)(2.0 * M_PI);
    else if (angle < -(float)M_PI) angle += (float)(2.0 * M_PI);
    double an = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    for (int i = 0; i < 3; ++i) e[3 + i] = (double)angle * ax[i] / an;

    double U[NJ * NJ], rhs[NJ], dq[NJ];
    for (int a = 0; a < NJ; ++a) {
      for (int b = 0; b < NJ; ++b) {
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) acc += J[k][a] * J[k][b];
        U[a * NJ + b] = acc;
      }
      U[a * NJ + a] += damping;
      double acc = 0.0;
      for (int k = 0; k < 6; ++k) acc += J[k][a] * e[k];
      rhs[a] = acc;
    }
    solve_gepp(NJ, U, rhs, dq);
    double mx = 0.0;
    for (int j = 0; j < NJ; ++j) mx = fmax(mx, fabs(dq[j]));
    if (mx > max_angle)
      for (int j = 0; j < NJ; ++j) dq[j] *= max_angle / mx;
    for (int j = 0; j < NJ; ++j) q[j] += dq[j];
    double pn[3];
    chain_fk(c, q, pn, NULL, NULL, NULL);
    diff = sqrt((tp[0] - pn[0]) * (tp[0] - pn[0]) + (tp[1] - pn[1]) * (tp[1] - pn[1]) + (tp[2] - pn[2]) * (tp[2] - pn[2]));
  }
  memcpy(q_out, q, sizeof(q));
  if (final_diff) *final_diff = diff;
  return it;
}

int orc_fk(int32_t robot, const double q[7], double pos[3], double rot[9], double origins[21], double axes[21]) {
  OrcChain c;
  if (chain_builtin(robot, &c)) return -1;
  double P[NJ][3], Z[NJ][3];
  chain_fk(&c, q, pos, rot, P, Z);
  if (origins) memcpy(origins, P, sizeof(P));
  if (axes) memcpy(axes, Z, sizeof(Z));
  return 0;
}

int orc_jacobian(int32_t robot, const double q[7], double Jout[42]) {
  OrcChain c;
  if (chain_builtin(robot, &c)) return -1;
  double pee[3], P[NJ][3], Z[NJ][3], J[6][NJ];
  chain_fk(&c, q, pee, NULL, P, Z);
  chain_jacobian(pee, P, Z, J);
  memcpy(Jout, J, sizeof(J));
  return 0;
}

int orc_ik(int32_t robot, const double q_in[7], const double target_pos[3], const double target_quat_xyzw[4], double damping,
           int max_iters, double residual, double q_out[7], double* final_diff) {
  OrcChain c;
  if (chain_builtin(robot, &c)) return -1;
  return chain_ik(&c, q_in, target_pos, target_quat_xyzw, damping, max_iters, residual, q_out, final_diff);
}

/* torque-mode dynamics of a built-in chain (aba_model.h); gravity (0,0,-10) */
int orc_aba(int32_t robot, const double q[7], const double qd[7], const double tau[7], double qdd[7]) {
  OrcChain c;
  if (chain_builtin(robot, &c)) return -1;
  aba_forward_dynamics(&c.dyn, q, qd, tau, qdd);
  return 0;
}

int orc_rnea(int32_t robot, const double q[7], const double qd[7], const double qdd[7], int with_gravity, double tau[7]) {
  OrcChain c;
  if (chain_builtin(robot, &c)) return -1;
  rnea_inverse_dynamics(&c.dyn, q, qd, qdd, with_gravity, tau);
  return 0;
}

int orc_dense_fd(int32_t robot, const double q[7], const double qd[7], const double tau[7], double qdd[7], double M[49]) {
  OrcChain c;
  if (chain_builtin(robot, &c)) return -1;
  dense_forward_dynamics(&c.dyn, q, qd, tau, qdd, M);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ sim */
/* diagnostics: contact-solver sweeps used so far by this process (not thread-safe: read it from single-thread runs) */
static _Thread_local uint64_t g_pgs_sweeps = 0, g_pgs_steps = 0;   /* per thread: the calling thread reads its own */
void orc_pgs_stats(uint64_t out[2], int reset) {
  out[0] = g_pgs_sweeps; out[1] = g_pgs_steps;
  if (reset) g_pgs_sweeps = g_pgs_steps = 0;
}

struct OrcSim {
  ArmsimConfig cfg;
  OrcChain chain;
  double tquat[4];
  int n;
  double (*q)[NJ];
  double (*qd)[NJ];     /* torque mode */
  float (*goal)[3];     /* reach goal / push,pick target -- held as f32 like `object_state.astype(np.float32)` */
  int32_t* step;
  int32_t* episode;
  int32_t* ik_iters;
  uint8_t* done;
  CubeState* cube;      /* push / pick */
  double* last_dist;
  double* grip;
  double* grip_dist;    /* diagnostic: the last getClosestPoints stand-in distance evaluated for the env (pick) */
};

static int32_t base_obs_dim(const OrcSim* s) {
  switch (s->cfg.task) {
    case ARMSIM_TASK_REACH: return 6;
    case ARMSIM_TASK_KUKA_REACH: return 3;
    default: return 9;
  }
}

/* diagnostic read-back for the parity tests: distance the last pick step compared with the 6 mm closing threshold */
void orc_grip_distance(const OrcSim* s, double* out) { memcpy(out, s->grip_dist, (size_t)s->n * sizeof(double)); }

int32_t orc_obs_dim(const OrcSim* s) { return base_obs_dim(s) + (s->cfg.mode == ARMSIM_MODE_TORQUE ? 2 * NJ : 0); }
int32_t orc_action_dim(const OrcSim* s) { return s->cfg.mode == ARMSIM_MODE_TORQUE ? ARMSIM_TORQUE_DIM : ARMSIM_ACT_DIM; }

OrcSim* orc_create(const ArmsimConfig* cfg) {
  if (!cfg || cfg->struct_size != (int32_t)sizeof(ArmsimConfig) || cfg->n_envs <= 0) return NULL;
  OrcSim* s = (OrcSim*)calloc(1, sizeof(OrcSim));
  s->cfg = *cfg;
  if (cfg->robot == ARMSIM_ROBOT_CUSTOM) {
    if (!cfg->custom_chain) { free(s); return NULL; }
    chain_from_custom(cfg->custom_chain, &s->chain);
  } else if (chain_builtin(cfg->robot, &s->chain)) { free(s); return NULL; }
  orc_quat_from_euler(cfg->target_rpy, s->tquat);
  memcpy(s->chain.dyn.gravity, cfg->gravity, sizeof(cfg->gravity));
  int n = s->n = cfg->n_envs;
  s->q = calloc(n, sizeof(*s->q));
  s->qd = calloc(n, sizeof(*s->qd));
  s->goal = calloc(n, sizeof(*s->goal));
  s->step = calloc(n, sizeof(int32_t));
  s->episode = calloc(n, sizeof(int32_t));
  s->ik_iters = calloc(n, sizeof(int32_t));
  s->done = calloc(n, 1);
  s->cube = calloc(n, sizeof(CubeState));
  s->last_dist = calloc(n, sizeof(double));
  s->grip = calloc(n, sizeof(double));
  s->grip_dist = calloc(n, sizeof(double));
  orc_reset(s, NULL, NULL);
  return s;
}

void orc_destroy(OrcSim* s) {
  if (!s) return;
  free(s->q); free(s->qd); free(s->goal); free(s->step); free(s->episode); free(s->ik_iters); free(s->done);
  free(s->cube); free(s->last_dist); free(s->grip); free(s->grip_dist);
  free(s);
}

static inline float lerp_f32(double lo, double hi, float u) { return fmaf((float)(hi - lo), u, (float)lo); }

static void write_obs(const OrcSim* s, int e, const double ee[3], float* obs) {
  const int od = orc_obs_dim(s);
  float* o = obs + (size_t)e * od;
  for (int i = 0; i < 3; ++i) o[i] = (float)ee[i];
  if (s->cfg.task == ARMSIM_TASK_REACH) {
    for (int i = 0; i < 3; ++i) o[3 + i] = s->goal[e][i];
  } else if (s->cfg.task == ARMSIM_TASK_PUSH || s->cfg.task == ARMSIM_TASK_PICK) {
    for (int i = 0; i < 3; ++i) o[3 + i] = (float)s->cube[e].pos[i];
    for (int i = 0; i < 3; ++i) o[6 + i] = s->goal[e][i];
  }
  if (s->cfg.mode == ARMSIM_MODE_TORQUE) {
    const int b = base_obs_dim(s);
    for (int j = 0; j < NJ; ++j) { o[b + j] = (float)s->q[e][j]; o[b + NJ + j] = (float)s->qd[e][j]; }
  }
}

/* Env.reset(): rl_reach_env.py:132-217, rl_push_env.py:145-256, rl_pick_env.py:141-256, kuka_reach_env.py:133-212.
 * The reference draws from Python's global `random` (Mersenne Twister); that stream cannot be reproduced on a GPU
 * (SURVEY 5), so both the device and this oracle draw from Philox4x32-10 keyed by (seed, global env id, episode) and
 * apply the reference's formulas to those uniforms: random.uniform(a, b) = a + (b - a) * u. */
static void reset_env(OrcSim* s, int e, float* obs) {
  const ArmsimConfig* c = &s->cfg;
  const uint64_t gid = c->env_id_offset + (uint64_t)e;
  const uint32_t ep = (uint32_t)s->episode[e];
  for (int j = 0; j < NJ; ++j) { s->q[e][j] = c->init_q[j]; s->qd[e][j] = 0.0; }  /* resetJointState(i, init_joint_positions[i]) :193-198 */
  s->step[e] = 0;                                           /* :135 */
  s->done[e] = 0;
  s->ik_iters[e] = 0;
  s->grip[e] = 0.0;
  float u[4];
  if (c->task == ARMSIM_TASK_REACH || c->task == ARMSIM_TASK_KUKA_REACH) {
    orc_reset_uniforms(c->seed, gid, ep, 0, u);
    for (int i = 0; i < 3; ++i) s->goal[e][i] = lerp_f32(c->goal_lo[i], c->goal_hi[i], u[i]);  /* :180-182 */
  } else {
    /* rejection loop rl_push_env.py:195-214: up to 1000 draws until 0.22 <= |cube - target| <= 0.25 */
    float cx = 0, cy = 0, cz = 0, cyaw = 0, tx = 0, ty = 0, tz = 0;
    for (uint32_t attempt = 0; attempt < 1000; ++attempt) {
      float v[4];
      orc_reset_uniforms(c->seed, gid, ep, 2 * attempt, u);
      orc_reset_uniforms(c->seed, gid, ep, 2 * attempt + 1, v);
      cx = lerp_f32(c->goal_lo[0], c->goal_hi[0], u[0]);
      cy = lerp_f32(c->goal_lo[1], c->goal_hi[1], u[1]);
      cz = 0.01f;                                                         /* :199 */
      cyaw = fmaf(3.1415925438f, u[2], 1.57f);                            /* :200 ang = 3.14*0.5 + 3.1415925438*random() */
      tx = lerp_f32(c->goal_lo[0], c->goal_hi[0], u[3]);
      ty = lerp_f32(c->goal_lo[1], c->goal_hi[1], v[0]);
      tz = (c->task == ARMSIM_TASK_PICK) ? lerp_f32(c->goal_lo[2], c->goal_hi[2], v[1]) : 0.01f;  /* rl_pick_env.py:202 */
      float ddx = cx - tx, ddy = cy - ty, ddz = cz - tz;
      float d2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
      float d = sqrtf(d2);
      if (d >= 0.22f && d <= 0.25f) break;                                /* :213 */
    }
    s->goal[e][0] = tx; s->goal[e][1] = ty; s->goal[e][2] = tz;
    cube_init(&s->cube[e], (double)cx, (double)cy, (double)cz, (double)cyaw);
  }
  s->episode[e] += 1;
  double ee[3];
  chain_fk(&s->chain, s->q[e], ee, NULL, NULL, NULL);        /* robot_pos_obs = getLinkState(...)[4]  :202 */
  if (c->task == ARMSIM_TASK_PUSH || c->task == ARMSIM_TASK_PICK) {
    /* p.stepSimulation() :242 then obs / last distances :243-245 */
    double Ree[9];
    chain_fk(&s->chain, s->q[e], ee, Ree, NULL, NULL);
    (void)cube_step(&s->cube[e], ee, Ree, c->task == ARMSIM_TASK_PICK, s->grip[e]);
    double d[3] = {s->cube[e].pos[0] - (double)s->goal[e][0], s->cube[e].pos[1] - (double)s->goal[e][1],
                   s->cube[e].pos[2] - (double)s->goal[e][2]};
    s->last_dist[e] = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  }
  if (obs) write_obs(s, e, ee, obs);
}

void orc_reset(OrcSim* s, const uint8_t* mask, float* obs) {
  for (int e = 0; e < s->n; ++e)
    if (!mask || mask[e]) reset_env(s, e, obs);
}

static inline double clip_val(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Env.step(): rl_reach_env.py:219-319 (reach), kuka_reach_env.py:214-305, rl_push_env.py:310-440,
 * rl_pick_env.py:310-445. */
static void step_env(OrcSim* s, int e, const float* action, float* obs, double* reward, uint8_t* done, uint8_t* success) {
  const ArmsimConfig* c = &s->cfg;
  if (s->done[e] && !c->auto_reset) {  /* a finished env waits for reset (the reference caller never steps it) */
    double ee[3];
    chain_fk(&s->chain, s->q[e], ee, NULL, NULL, NULL);
    if (obs) write_obs(s, e, ee, obs);
    reward[e] = 0.0; done[e] = 1; success[e] = 0;
    return;
  }
  if (c->mode == ARMSIM_MODE_TORQUE) {
    /* north-star addition: the servo + teleport is replaced by one dynamics step */
    aba_torque_step(&s->chain.dyn, s->chain.effort, s->chain.velocity, s->chain.damping, s->chain.lower, s->chain.upper,
                    c->sim_dt, action + (size_t)e * ARMSIM_TORQUE_DIM, s->q[e], s->qd[e]);
    s->ik_iters[e] = 0;
    goto after_servo;
  }
  const float* a = action + (size_t)e * ARMSIM_ACT_DIM;
  double cur[3], Rcur[9];
  chain_fk(&s->chain, s->q[e], cur, Rcur, NULL, NULL);                  /* current_pos = getLinkState(...)[4]  :237 */
  if (c->task == ARMSIM_TASK_PICK)
    for (int i = 0; i < 3; ++i) cur[i] = (double)(float)cur[i];         /* rl_pick_env.py:327 .astype(np.float32) */
  double tgt[3];
  for (int i = 0; i < 3; ++i) {
    double d = (double)a[i] * c->dv;                                    /* dx = action[0] * dv  :232-234 */
    tgt[i] = cur[i] + d;
    if (c->task != ARMSIM_TASK_KUKA_REACH) tgt[i] = clip_val(tgt[i], c->ws_lo[i], c->ws_hi[i]);  /* :239-242 */
  }
  double qn[NJ];
  s->ik_iters[e] = chain_ik(&s->chain, s->q[e], tgt, s->tquat, c->ik_damping, c->ik_max_iters, c->ik_residual, qn, NULL);
  /* resetJointState for joints 0..6 (:252-257); pick only joints 0..5 (rl_pick_env.py:342-347) */
  const int napply = (c->task == ARMSIM_TASK_PICK) ? 6 : NJ;
  for (int j = 0; j < napply; ++j) s->q[e][j] = qn[j];
  if (c->clamp_joint_limits)
    for (int j = 0; j < NJ; ++j) s->q[e][j] = clip_val(s->q[e][j], s->chain.lower[j], s->chain.upper[j]);
after_servo:;
  /* p.stepSimulation() :258 -- the arm is held by Bullet's default velocity motors: static */
  double ee[3], Ree[9];
  chain_fk(&s->chain, s->q[e], ee, Ree, NULL, NULL);
  s->step[e] += 1;                                                      /* :264 */

  double r = 0.0;
  int term = 0, succ = 0, wrote_obs = 0;
  if (c->task == ARMSIM_TASK_REACH) {
    /* _reward rl_reach_env.py:267-319 */
    double d0 = ee[0] - (double)s->goal[e][0], d1 = ee[1] - (double)s->goal[e][1], d2 = ee[2] - (double)s->goal[e][2];
    double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);                    /* :281 */
    if (s->step[e] > c->max_steps) { r = -dist * 10.0; term = 1; }      /* :299-301 */
    else if (dist < c->reach_dis) { r = 0.0; term = 1; succ = 1; }      /* :303-306 */
    else { r = -dist * 10.0; term = 0; }                                /* :307-309 */
  } else if (c->task == ARMSIM_TASK_KUKA_REACH) {
    /* kuka_reach_env.py:252-305 */
    double d0 = ee[0] - (double)s->goal[e][0], d1 = ee[1] - (double)s->goal[e][1], d2 = ee[2] - (double)s->goal[e][2];
    double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    int oob = ee[0] < c->ws_lo[0] || ee[0] > c->ws_hi[0] || ee[1] < c->ws_lo[1] || ee[1] > c->ws_hi[1] ||
              ee[2] < c->ws_lo[2] || ee[2] > c->ws_hi[2];               /* :276-278 */
    if (oob) { r = -1.0; term = 1; }                                    /* :280-282 */
    else if (s->step[e] > c->max_steps) { r = -1.0; term = 1; }         /* :285-287 */
    else if (dist < c->reach_dis) { r = 10.0; term = 1; succ = 1; }     /* :289-291 */
    else { r = 0.0; term = 0; }
  } else {
    /* push / pick: cube dynamics inside stepSimulation, then _reward rl_push_env.py:368-440 / rl_pick_env.py:367-445 */
    const int pick = c->task == ARMSIM_TASK_PICK;
    g_pgs_sweeps += (uint64_t)cube_step(&s->cube[e], ee, Ree, pick, s->grip[e]);
    g_pgs_steps += 1;
    const CubeState* cb = &s->cube[e];
    double d[3] = {cb->pos[0] - (double)s->goal[e][0], cb->pos[1] - (double)s->goal[e][1], cb->pos[2] - (double)s->goal[e][2]};
    double dist_cur = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);    /* :383 distance_current (fp64 obs) */
    double test = dist_cur - s->last_dist[e];                           /* :385 */
    if (fabs(test) < 1e-5) test = 0.01;                                 /* :386-387 */
    s->last_dist[e] = dist_cur;                                         /* :389-390 */
    /* distance_target: norm of f32 object_state - f32 target_state (:393), evaluated in f32 */
    float f0 = (float)cb->pos[0] - s->goal[e][0], f1 = (float)cb->pos[1] - s->goal[e][1], f2 = (float)cb->pos[2] - s->goal[e][2];
    double dist_t = (double)sqrtf(f0 * f0 + f1 * f1 + f2 * f2);
    if (s->step[e] > c->max_steps) { r = -dist_t * 50.0; term = 1; }    /* :417-419 */
    else if (dist_t < 0.05) { r = 100.0; term = 1; }                    /* :421-423 */
    else { r = -test * 100.0; term = 0; }                               /* :424-428 */
    succ = dist_cur < c->reach_dis;                                     /* _is_success :442-445 on the fp64 obs */
    if (pick) {
      /* rl_pick_env.py:403-417: obs / distances above were taken BEFORE this block (`obs = self._get_obs()` :381);
       * any arm link within 6 mm of the cube -> the four finger joints snap to 0 (never reopened before reset),
       * then a 2nd sim step whose effect the NEXT env step observes */
      if (obs && !(term && c->auto_reset)) write_obs(s, e, ee, obs);
      wrote_obs = 1;
      s->grip_dist[e] = cube_gripper_distance(&s->cube[e], ee, Ree, s->grip[e]);
      if (s->grip[e] < 0.5 && s->grip_dist[e] < PICK_CLOSE_DIST) s->grip[e] = 1.0;
      g_pgs_sweeps += (uint64_t)cube_step(&s->cube[e], ee, Ree, pick, s->grip[e]);
      g_pgs_steps += 1;
    }
  }
  s->done[e] = (uint8_t)term;
  reward[e] = r;
  done[e] = (uint8_t)term;
  success[e] = (uint8_t)succ;
  if (term && c->auto_reset) reset_env(s, e, obs);
  else if (obs && !wrote_obs) write_obs(s, e, ee, obs);
}

void orc_step_range(OrcSim* s, int32_t lo, int32_t hi, const float* action, float* obs, double* reward, uint8_t* done,
                    uint8_t* success) {
  if (lo < 0) lo = 0;
  if (hi > s->n) hi = s->n;
  for (int e = lo; e < hi; ++e) step_env(s, e, action, obs, reward, done, success);
}

void orc_step(OrcSim* s, const float* action, float* obs, double* reward, uint8_t* done, uint8_t* success) {
  orc_step_range(s, 0, s->n, action, obs, reward, done, success);
}

/* ------------------------------------------------------------------------------------------------ state io */
static int field_width(int32_t f) {
  switch (f) {
    case ARMSIM_F_Q: case ARMSIM_F_QD: return 7;
    case ARMSIM_F_GOAL: case ARMSIM_F_CUBE_POS: case ARMSIM_F_CUBE_LINVEL: case ARMSIM_F_CUBE_ANGVEL: return 3;
    case ARMSIM_F_CUBE_QUAT: return 4;
    case ARMSIM_F_STEP: case ARMSIM_F_EPISODE: case ARMSIM_F_LAST_DIST: case ARMSIM_F_GRIP: case ARMSIM_F_IK_ITERS: return 1;
    default: return -1;
  }
}

static double* field_f64(OrcSim* s, int32_t f, int e) {
  switch (f) {
    case ARMSIM_F_Q: return s->q[e];
    case ARMSIM_F_QD: return s->qd[e];
    case ARMSIM_F_CUBE_POS: return s->cube[e].pos;
    case ARMSIM_F_CUBE_QUAT: return s->cube[e].quat;
    case ARMSIM_F_CUBE_LINVEL: return s->cube[e].v;
    case ARMSIM_F_CUBE_ANGVEL: return s->cube[e].w;
    case ARMSIM_F_LAST_DIST: return &s->last_dist[e];
    case ARMSIM_F_GRIP: return &s->grip[e];
    default: return NULL;
  }
}

int orc_set_state(OrcSim* s, int32_t field, const void* src, size_t bytes) {
  int w = field_width(field);
  if (w < 0 || bytes != (size_t)s->n * w * 4) return ARMSIM_E_STATE;
  for (int e = 0; e < s->n; ++e) {
    if (field == ARMSIM_F_STEP) { s->step[e] = ((const int32_t*)src)[e]; s->done[e] = 0; }
    else if (field == ARMSIM_F_EPISODE) s->episode[e] = ((const int32_t*)src)[e];
    else if (field == ARMSIM_F_GOAL) memcpy(s->goal[e], (const float*)src + 3 * e, 12);
    else {
      double* d = field_f64(s, field, e);
      if (!d) return ARMSIM_E_STATE;
      for (int k = 0; k < w; ++k) d[k] = (double)((const float*)src)[e * w + k];
      if (field == ARMSIM_F_Q) s->done[e] = 0;
    }
  }
  return ARMSIM_OK;
}

int orc_get_state(OrcSim* s, int32_t field, void* dst, size_t bytes) {
  int w = field_width(field);
  if (w < 0 || bytes != (size_t)s->n * w * 4) return ARMSIM_E_STATE;
  for (int e = 0; e < s->n; ++e) {
    if (field == ARMSIM_F_STEP) ((int32_t*)dst)[e] = s->step[e];
    else if (field == ARMSIM_F_EPISODE) ((int32_t*)dst)[e] = s->episode[e];
    else if (field == ARMSIM_F_IK_ITERS) ((int32_t*)dst)[e] = s->ik_iters[e];
    else if (field == ARMSIM_F_GOAL) memcpy((float*)dst + 3 * e, s->goal[e], 12);
    else {
      double* d = field_f64(s, field, e);
      if (!d) return ARMSIM_E_STATE;
      for (int k = 0; k < w; ++k) ((float*)dst)[e * w + k] = (float)d[k];
    }
  }
  return ARMSIM_OK;
}

int orc_get_state_f64(OrcSim* s, int32_t field, double* dst, size_t count) {
  int w = field_width(field);
  if (w < 0 || count != (size_t)s->n * w) return ARMSIM_E_STATE;
  for (int e = 0; e < s->n; ++e) {
    if (field == ARMSIM_F_GOAL) { for (int k = 0; k < 3; ++k) dst[3 * e + k] = (double)s->goal[e][k]; continue; }
    double* d = field_f64(s, field, e);
    if (!d) return ARMSIM_E_STATE;
    for (int k = 0; k < w; ++k) dst[e * w + k] = d[k];
  }
  return ARMSIM_OK;
}

int orc_default_config(int32_t task, ArmsimConfig* cfg) { return armsim_fill_default_config(task, cfg); }

/* ------------------------------------------------------------------------------------------------ threaded stepping
 * CPU-baseline leg of bench.py: one Env.step of the whole batch on `nthreads` persistent worker threads, each owning a
 * contiguous slice of the envs (they are independent units).  Workers spin on a generation counter (sense-reversing
 * barrier with a short pause / yield back-off) so that a 4096-env step, ~0.3 ms of work on 32 threads, is not
 * dominated by futex wake-ups the way a Python thread pool or pthread_barrier_wait would be. */
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>

typedef struct OrcPool OrcPool;
typedef struct OrcWorker { OrcPool* pool; int idx; pthread_t th; } OrcWorker;
struct OrcPool {
  OrcSim* sim;
  int nthreads;
  OrcWorker* w;
  atomic_uint gen;        /* bumped by the caller to start a step */
  atomic_int pending;     /* workers still running the current step */
  atomic_int quit;
  const float* action; float* obs; double* reward; uint8_t* done; uint8_t* success;
};

static inline void orc_cpu_relax(unsigned spins) {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
  if ((spins & 0x3FFu) == 0x3FFu) sched_yield();
}

static void orc_pool_slice(OrcPool* p, int idx) {
  const int n = p->sim->n, t = p->nthreads;
  const int lo = (int)((long long)n * idx / t), hi = (int)((long long)n * (idx + 1) / t);
  orc_step_range(p->sim, lo, hi, p->action, p->obs, p->reward, p->done, p->success);
}

static void* orc_pool_main(void* arg) {
  OrcWorker* w = (OrcWorker*)arg;
  OrcPool* p = w->pool;
  unsigned seen = 0;
  for (;;) {
    unsigned spins = 0;
    while (atomic_load_explicit(&p->gen, memory_order_acquire) == seen) {
      if (atomic_load_explicit(&p->quit, memory_order_relaxed)) return NULL;
      orc_cpu_relax(++spins);
    }
    seen = atomic_load_explicit(&p->gen, memory_order_acquire);
    orc_pool_slice(p, w->idx);
    atomic_fetch_sub_explicit(&p->pending, 1, memory_order_release);
  }
}

OrcPool* orc_pool_create(OrcSim* s, int nthreads) {
  if (!s || nthreads < 1) return NULL;
  OrcPool* p = (OrcPool*)calloc(1, sizeof(OrcPool));
  p->sim = s; p->nthreads = nthreads;
  p->w = (OrcWorker*)calloc((size_t)nthreads, sizeof(OrcWorker));
  atomic_init(&p->gen, 0u); atomic_init(&p->pending, 0); atomic_init(&p->quit, 0);
  for (int i = 1; i < nthreads; ++i) {          /* slice 0 runs on the calling thread */
    p->w[i].pool = p; p->w[i].idx = i;
    pthread_create(&p->w[i].th, NULL, orc_pool_main, &p->w[i]);
  }
  return p;
}

void orc_pool_destroy(OrcPool* p) {
  if (!p) return;
  atomic_store(&p->quit, 1);
  for (int i = 1; i < p->nthreads; ++i) pthread_join(p->w[i].th, NULL);
  free(p->w); free(p);
}

/* one Env.step of every env of the pool's sim, work split over the pool's threads; returns when all slices are done */
void orc_step_mt(OrcPool* p, const float* action, float* obs, double* reward, uint8_t* done, uint8_t* success) {
  p->action = action; p->obs = obs; p->reward = reward; p->done = done; p->success = success;
  atomic_store_explicit(&p->pending, p->nthreads - 1, memory_order_relaxed);
  atomic_fetch_add_explicit(&p->gen, 1u, memory_order_release);
  orc_pool_slice(p, 0);
  unsigned spins = 0;
  while (atomic_load_explicit(&p->pending, memory_order_acquire) > 0) orc_cpu_relax(++spins);
}

/* `steps` consecutive Env.steps inside ONE call (action set k % n_sets of a [n_sets, n, act_dim] ring): what a C host
 * loop around the reference step would cost, with no per-step Python / ctypes dispatch at all */
void orc_run_mt(OrcPool* p, const float* actions, int n_sets, int steps, float* obs, double* reward, uint8_t* done, uint8_t* success) {
  const size_t stride = (size_t)p->sim->n * (size_t)orc_action_dim(p->sim);
  for (int k = 0; k < steps; ++k) orc_step_mt(p, actions + (size_t)(k % n_sets) * stride, obs, reward, done, success);
}
