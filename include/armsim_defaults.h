/* armsim_defaults.h -- the reference's constants per task, as one inline function shared by libarmsim and the
 * test oracle (so neither depends on the other).  Every value cites where the reference sets it. */
#ifndef ARMSIM_DEFAULTS_H
#define ARMSIM_DEFAULTS_H
#include <string.h>
#include "armsim.h"

static inline int armsim_fill_default_config(int32_t task, ArmsimConfig* c) {
  if (!c || task < 0 || task > ARMSIM_TASK_KUKA_REACH) return ARMSIM_E_INVALID;
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(ArmsimConfig);
  c->task = task;
  c->robot = ARMSIM_ROBOT_KUKA_IIWA;          /* rl_reach_env.py:174 kuka_iiwa/model.urdf */
  c->mode = ARMSIM_MODE_IK_TELEPORT;
  c->mapping = ARMSIM_MAP_AUTO;
  c->n_envs = 1;
  c->device = 0;
  c->auto_reset = 0;
  c->seed = 0;                                /* config.py:47 random_seed = 0 */
  c->env_id_offset = 0;
  /* workspace box, rl_reach_env.py:65-70 (x_low_obs .. z_high_obs) and :221-223 (limit_x/y/z) */
  c->ws_lo[0] = 0.2;  c->ws_hi[0] = 0.7;
  c->ws_lo[1] = -0.3; c->ws_hi[1] = 0.3;
  c->ws_lo[2] = 0.0;  c->ws_hi[2] = 0.55;
  for (int i = 0; i < 3; ++i) { c->goal_lo[i] = c->ws_lo[i]; c->goal_hi[i] = c->ws_hi[i]; }  /* :180-182 */
  c->max_steps = 500;                         /* config.py:51 */
  switch (task) {
    case ARMSIM_TASK_REACH:
      c->dv = 0.02;                          /* config.py:41 reach_ctr */
      c->reach_dis = 0.01;                   /* config.py:42 reach_dis */
      break;
    case ARMSIM_TASK_PUSH:
      c->dv = 0.08;                          /* rl_push_env.py:322 */
      c->reach_dis = 0.05;                   /* rl_push_env.py:86 distance_threshold, :421 */
      c->ws_hi[2] = 0.1;                     /* rl_push_env.py:314 limit_z */
      c->goal_lo[2] = c->goal_hi[2] = 0.01;  /* rl_push_env.py:199,206 zpos = 0.01 */
      break;
    case ARMSIM_TASK_PICK:
      c->dv = 0.08;                          /* rl_pick_env.py:321 */
      c->reach_dis = 0.05;
      c->ws_hi[2] = 0.55 + 0.257;           /* rl_pick_env.py:313 0.55 + gripper_length (:79) */
      /* cube z = 0.01 (:199), target z ~ U(0, 0.55) (:202): goal box keeps the full z range for the target */
      break;
    case ARMSIM_TASK_KUKA_REACH:
      c->dv = 0.005;                         /* kuka_reach_env.py:215 */
      c->reach_dis = 0.1;                    /* kuka_reach_env.py:289 */
      c->max_steps = 1000;                    /* kuka_reach_env.py:59 */
      c->goal_lo[2] = c->goal_hi[2] = 0.01;  /* kuka_reach_env.py:184 object z = 0.01 */
      break;
  }
  /* IK target orientation p.getQuaternionFromEuler([0, -pi, pi/2]), rl_reach_env.py:121-122 */
  c->target_rpy[0] = 0.0; c->target_rpy[1] = -3.14159265358979323846; c->target_rpy[2] = 1.57079632679489661923;
  /* init_joint_positions, rl_reach_env.py:116-119 */
  static const double q0[ARMSIM_NJ] = {0.006418, 0.413184, -0.011401, -1.589317, 0.005379, 1.137684, -0.006539};
  for (int i = 0; i < ARMSIM_NJ; ++i) c->init_q[i] = q0[i];
  c->ik_damping = 1e-5;                      /* rl_reach_env.py:111-113 joint_damping */
  c->ik_max_iters = 20;                       /* Bullet default maxNumIterations */
  c->ik_residual = 1e-4;                     /* Bullet default residualThreshold */
  c->clamp_joint_limits = 0;
  c->sim_dt = 1.0 / 240.0;                   /* Bullet default fixedTimeStep used by p.stepSimulation() */
  c->gravity[0] = 0.0; c->gravity[1] = 0.0; c->gravity[2] = -10.0;   /* rl_reach_env.py:142 */
  c->custom_chain = NULL;
  return ARMSIM_OK;
}
#endif
