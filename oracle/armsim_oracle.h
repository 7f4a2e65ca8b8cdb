/* armsim_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64) of the reference hot path, used ONLY as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * drl-on-robot-arm_b200/ may link, import or call this library.
 *
 * Parity status: FK / workspace clip / reward / done / success / episode length are PINNED by the
 * reference's own artefacts (main.py:106 golden EE vector, envs/bmirobot_joints_info_pybullet.txt fixture, the
 * Python control flow).  The arithmetic that lives in the third-party dependency pybullet==3.0.6 (ReadMe.md:17;
 * source absent from /root/reference, wheel not installable) -- calculateInverseKinematics and stepSimulation --
 * is restated from Bullet's published algorithm and is PARITY UNPINNED: no reference test or fixture fixes its
 * numerics.  Known open points of the restatement, each with its consequence:
 *   - Jacobian reference point.  chain_jacobian() takes the linear Jacobian at the origin of the EE link FRAME (the
 *     point whose position error drives the iteration).  Bullet's processCalculateInverseKinematicsCommand obtains its
 *     body Jacobian from btInverseDynamics, which (as far as can be recalled without the source) refers it to the
 *     link's INERTIAL frame: 2 cm up the local z axis for link 7.  If so, Bullet's linear rows
 *     differ by omega x (0.02 m lever): another iterate path and sometimes another iteration count, but the same
 *     stopping rule on the link-frame position, i.e. the same end point to within the 1e-4 m residual.  Nothing
 *     downstream of the step (obs, reward, done) sees more than that.  Measured (tools/ik_reference_point_study.py,
 *     1500 servo moves of the reach workspace, both Jacobians from the same start): end points differ by 2.9e-6 m at
 *     the median, 4.7e-5 m at the 99th percentile, 1.0e-4 m at most; same iteration count in 99.8 % of the moves
 *     (profiles/r02_ik_reference_point_study.json).
 *   - Orientation error is formed as 2 atan2(|v|, w) v/|v| instead of 2 acos(w) v/sqrt(1-w^2): the same value for a
 *     unit quaternion, better conditioned near zero.
 *   - An iteration that does not converge within 20 updates (targets clipped to an unreachable workspace corner, or
 *     pick's untouched joint 7) is CHAOTIC with damping 1e-5: changing the action by 2e-6 relative moves this oracle's
 *     own end point by up to 7 cm (tests/test_parity_gpu.py::test_cube_tasks_teacher_forced measures it).  For those
 *     env-steps no implementation, Bullet included, is reproducible beyond the statistics.
 *   - stepSimulation for the free cube: oracle/cube_model.h states the model; pinned only at system level (the
 *     reference's untouched-cube return and its push learning curve).
 */
#ifndef ARMSIM_ORACLE_H
#define ARMSIM_ORACLE_H
#include <stdint.h>
#include "armsim.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcSim OrcSim;

/* Philox4x32-10 (Salmon et al. 2011), the counter-based generator the device reset uses. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the 4 uniforms in [0,1) of reset draw `block` of (seed, global env id, episode) */
void orc_reset_uniforms(uint64_t seed, uint64_t env_gid, uint32_t episode, uint32_t block, float u[4]);

/* forward kinematics of a built-in chain (robot = ARMSIM_ROBOT_*): EE (link-7 frame) position, rotation (row-major),
 * the 7 joint-frame origins and world joint axes.  Any output may be NULL. */
int orc_fk(int32_t robot, const double q[7], double pos[3], double rot[9], double origins[21], double axes[21]);
/* 6x7 geometric Jacobian of the EE link frame, rows 0-2 linear, 3-5 angular, row-major */
int orc_jacobian(int32_t robot, const double q[7], double J[42]);
/* Bullet calculateInverseKinematics restated.  Returns the iteration count. */
int orc_ik(int32_t robot, const double q_in[7], const double target_pos[3], const double target_quat_xyzw[4],
           double damping, int max_iters, double residual, double q_out[7], double* final_diff);
void orc_quat_from_euler(const double rpy[3], double quat_xyzw[4]);
/* torque-mode dynamics (oracle/aba_model.h): articulated-body forward dynamics, recursive Newton-Euler inverse
 * dynamics, and the dense route qdd = M^-1 (tau - h) built from RNEA (M optional, 7x7 row-major) */
int orc_aba(int32_t robot, const double q[7], const double qd[7], const double tau[7], double qdd[7]);
int orc_rnea(int32_t robot, const double q[7], const double qd[7], const double qdd[7], int with_gravity, double tau[7]);
int orc_dense_fd(int32_t robot, const double q[7], const double qd[7], const double tau[7], double qdd[7], double M[49]);

OrcSim* orc_create(const ArmsimConfig* cfg);
void orc_destroy(OrcSim* s);
void orc_reset(OrcSim* s, const uint8_t* mask, float* obs);
/* one Env.step for envs [lo, hi) (disjoint ranges may run on different threads) */
void orc_step_range(OrcSim* s, int32_t lo, int32_t hi, const float* action, float* obs, double* reward, uint8_t* done,
                    uint8_t* success);
void orc_step(OrcSim* s, const float* action, float* obs, double* reward, uint8_t* done, uint8_t* success);
/* state access with the ARMSIM_F_* field ids; f32/i32 host arrays like armsim_set_state, plus an fp64 read-back */
int orc_set_state(OrcSim* s, int32_t field, const void* src, size_t bytes);
int orc_get_state(OrcSim* s, int32_t field, void* dst, size_t bytes);
int orc_get_state_f64(OrcSim* s, int32_t field, double* dst, size_t count);
int32_t orc_obs_dim(const OrcSim* s);
/* diagnostics for the parity tests */
void orc_grip_distance(const OrcSim* s, double* out);          /* pick: last distance compared with the 6 mm threshold */
void orc_pgs_stats(uint64_t out[2], int reset);                 /* {contact-solver sweeps, cube sim steps} so far */
int32_t orc_action_dim(const OrcSim* s);
int orc_default_config(int32_t task, ArmsimConfig* cfg);

/* persistent worker threads for the CPU-baseline leg of bench.py: each Env.step is split over `nthreads` slices */
typedef struct OrcPool OrcPool;
OrcPool* orc_pool_create(OrcSim* s, int nthreads);
void orc_pool_destroy(OrcPool* p);
void orc_step_mt(OrcPool* p, const float* action, float* obs, double* reward, uint8_t* done, uint8_t* success);
void orc_run_mt(OrcPool* p, const float* actions, int n_sets, int steps, float* obs, double* reward, uint8_t* done, uint8_t* success);

#ifdef __cplusplus
}
#endif
#endif
