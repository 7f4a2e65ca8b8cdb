/* armsim.h -- C-ABI of the B200-native batched robot-arm step engine (libarmsim.so).
 *
 * Drop-in boundary for ONE hot path of Shimly-2/DRL-on-robot-arm: the gym-style Env.reset()/Env.step()
 * of envs/rl_reach_env.py, envs/rl_push_env.py, envs/rl_pick_env.py and envs/kuka_reach_env.py, which in the
 * reference drive ONE arm per process through pybullet on the CPU.  The reference has no FFI of its own (its
 * boundary is Python duck-typing of gym.Env, main.py:83 `getattr(envs, opt.env)(...)`, envs/__init__.py:1-3);
 * these entry points are what a ctypes/pybind binding inside those Env classes binds to.
 *
 * Conventions
 *   - plain C types, pointers and sizes only; no exceptions cross the ABI; 0 = OK, negative = ARMSIM_E_*;
 *     armsim_last_error() returns a thread-local message for the last failing call.
 *   - `*_dev` pointers are CUDA device pointers on the handle's device, `*_host` pointers are host memory;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device entry points are
 *     asynchronous on `stream` and never synchronise; the *_host entry points synchronise before returning.
 *   - the caller owns every I/O buffer; the library owns the per-env state (struct-of-arrays in HBM).
 *   - one handle per GPU; a handle is not thread-safe.
 *   - there is NO CPU fallback: armsim_create fails with ARMSIM_E_CUDA when no CUDA device is usable.
 */
#ifndef ARMSIM_H
#define ARMSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARMSIM_ABI_VERSION 5
#define ARMSIM_NJ 7            /* arm joints (Kuka iiwa, DianaS1) */
#define ARMSIM_ACT_DIM 3       /* Cartesian EE servo action, reference envs' action_space */
#define ARMSIM_TORQUE_DIM 7    /* torque mode action */

/* error codes */
#define ARMSIM_OK 0
#define ARMSIM_E_INVALID (-1)  /* bad argument / config */
#define ARMSIM_E_CUDA (-2)     /* CUDA runtime error (message has the cudaError string) */
#define ARMSIM_E_NOMEM (-3)
#define ARMSIM_E_STATE (-4)    /* unknown state field / wrong size */

/* tasks: which reference Env.step the fused kernel reproduces */
enum {
  ARMSIM_TASK_REACH = 0,       /* envs/rl_reach_env.py:219-319   obs f32[6] = [ee, goal]            */
  ARMSIM_TASK_PUSH = 1,        /* envs/rl_push_env.py:310-445    obs f32[9] = [ee, cube, target]    */
  ARMSIM_TASK_PICK = 2,        /* envs/rl_pick_env.py:310-450    obs f32[9] = [link-6, cube, target]*/
  ARMSIM_TASK_KUKA_REACH = 3   /* envs/kuka_reach_env.py:214-305 obs f32[3] = ee                    */
};
enum { ARMSIM_ROBOT_KUKA_IIWA = 0, ARMSIM_ROBOT_DIANA_S1 = 1, ARMSIM_ROBOT_CUSTOM = 2 };
enum {
  ARMSIM_MODE_IK_TELEPORT = 0, /* what the reference does: EE target -> DLS IK -> teleport joints */
  ARMSIM_MODE_TORQUE = 1       /* north-star addition: joint torques [n,7] -> effort clip + joint damping -> ABA forward
                                  dynamics -> semi-implicit Euler (sim_dt) -> velocity clip -> joint-limit clamp -> the
                                  task's reward.  obs = task obs followed by q[7], qd[7] (obs_dim + 14)              */
};
/* Thread mapping of the step kernels.  There is ONE: a CUDA lane per arm (struct-of-arrays state, warp-local
 * staging of the row-major I/O).  A four-lanes-per-arm mapping for single-wave batches (the joints' sincos split over
 * a quad, everything else redundant) was built, proven bit-identical and measured in round 2: 13 % fewer instructions
 * per warp but 4.08 us against 3.55 us per 4096-arm launch (shuffle latency on the dependent chain, four times the
 * instruction-fetch and load traffic) -- profiles/r02_quad_mapping_experiment.patch, r02_ncu_summary.json keys
 * r02_reach_n4096 / r02_reach_n4096_quad.  The `mapping` field of ArmsimConfig is kept for ABI stability and must be
 * ARMSIM_MAP_AUTO or ARMSIM_MAP_LANE. */
enum {
  ARMSIM_MAP_AUTO = 0,         /* = ARMSIM_MAP_LANE */
  ARMSIM_MAP_LANE = 1          /* one CUDA lane per arm */
};

/* A custom 7-DoF serial chain (ARMSIM_ROBOT_CUSTOM).  Same content as include/armsim_robot_models.h. */
typedef struct ArmsimChain {
  double base_xyz[3], base_rpy[3];
  double xyz[ARMSIM_NJ][3], rpy[ARMSIM_NJ][3];         /* joint origin in parent link frame; joint axis = local +z */
  double lower[ARMSIM_NJ], upper[ARMSIM_NJ], effort[ARMSIM_NJ], velocity[ARMSIM_NJ], damping[ARMSIM_NJ];
  double mass[ARMSIM_NJ], com[ARMSIM_NJ][3], inertia[ARMSIM_NJ][6];
} ArmsimChain;

typedef struct ArmsimConfig {
  int32_t struct_size;         /* = sizeof(ArmsimConfig), ABI guard */
  int32_t task, robot, mode, mapping;
  int32_t n_envs;              /* envs owned by THIS handle (this rank's shard) */
  int32_t device;              /* CUDA ordinal */
  int32_t auto_reset;          /* 1: envs that finish are re-initialised inside the same launch (obs = first obs of
                                  the next episode; reward/done/success describe the finished step) */
  uint64_t seed;               /* Philox key */
  uint64_t env_id_offset;      /* global id of env 0 (sharding: results do not depend on the rank layout) */
  double dv;                   /* EE metres per unit action: opt.reach_ctr 0.02 (config.py:41) / 0.08 push,pick / 0.005 kuka_reach */
  double reach_dis;            /* success distance: opt.reach_dis 0.01 (config.py:42); 0.05 push/pick; 0.1 kuka_reach */
  int32_t max_steps;           /* opt.max_steps_one_episode 500 (config.py:51): done when step_counter > max_steps */
  double ws_lo[3], ws_hi[3];   /* EE target clip box (rl_reach_env.py:221-223); kuka_reach: OOB box, no clip */
  double goal_lo[3], goal_hi[3];/* reset sampling box for goal / cube / target (rl_reach_env.py:180-182) */
  double target_rpy[3];        /* IK target orientation euler(0,-pi,pi/2) (rl_reach_env.py:121-122) */
  double init_q[ARMSIM_NJ];    /* init_joint_positions (rl_reach_env.py:116-119) */
  double ik_damping;           /* jointDamping 1e-5 (rl_reach_env.py:111-113) */
  int32_t ik_max_iters;        /* Bullet default 20 */
  double ik_residual;          /* Bullet default 1e-4 */
  int32_t clamp_joint_limits;  /* 0 = reference behaviour (no clamp in IK mode); torque mode always clamps */
  int32_t reserved[7];
  double sim_dt;               /* torque mode: integration step, Bullet default 1/240 s (p.stepSimulation) */
  double gravity[3];           /* torque mode: world gravity, (0,0,-10) (rl_reach_env.py:142 p.setGravity) */
  const ArmsimChain* custom_chain; /* only for ARMSIM_ROBOT_CUSTOM */
} ArmsimConfig;

typedef struct ArmSim ArmSim;  /* opaque handle */

/* state fields for armsim_get_state / armsim_set_state (host arrays, row-major [n_envs, width]) */
enum {
  ARMSIM_F_Q = 0,              /* f32 [n,7]  joint angles                         */
  ARMSIM_F_QD = 1,             /* f32 [n,7]  joint velocities (torque mode)        */
  ARMSIM_F_GOAL = 2,           /* f32 [n,3]  reach goal / push,pick target         */
  ARMSIM_F_STEP = 3,           /* i32 [n]    step_counter                          */
  ARMSIM_F_EPISODE = 4,        /* i32 [n]    episodes started (Philox counter)     */
  ARMSIM_F_CUBE_POS = 5,       /* f32 [n,3]                                        */
  ARMSIM_F_CUBE_QUAT = 6,      /* f32 [n,4]  xyzw                                  */
  ARMSIM_F_CUBE_LINVEL = 7,    /* f32 [n,3]                                        */
  ARMSIM_F_CUBE_ANGVEL = 8,    /* f32 [n,3]                                        */
  ARMSIM_F_LAST_DIST = 9,      /* f32 [n]    previous cube-target distance (push reward) */
  ARMSIM_F_GRIP = 10,          /* f32 [n]    pick: 1 = fingers closed              */
  ARMSIM_F_IK_ITERS = 11,      /* i32 [n]    read-only: DLS iterations used by the last step */
  ARMSIM_F_EP_RETURN = 12,     /* f32 [n]    return of the running episode (armsim_track_episodes) */
  ARMSIM_F_EXPLORE_COUNT = 13, /* i32 [n]    exploration-noise draws so far (Philox counter of armsim_explore) */
  ARMSIM_F_COUNT = 14
};

/* Fill *cfg with the reference constants for `task` (robot kuka_iiwa, IK-teleport mode, n_envs = 1). */
int armsim_default_config(int32_t task, ArmsimConfig* cfg);

/* Create n_envs simulators on cfg->device and reset them all (episode 0).  Replaces Env.__init__
 * (rl_reach_env.py:44-125: p.connect + constants + first reset). */
int armsim_create(const ArmsimConfig* cfg, ArmSim** out);

/* Env.close() (rl_reach_env.py:321-322). */
void armsim_destroy(ArmSim* sim);

/* Env.reset() for the envs whose mask byte is non-zero (mask_dev == NULL: all).  Replaces rl_reach_env.py:132-217 /
 * rl_push_env.py:145-256 / rl_pick_env.py:141-256.  obs_dev f32 [n, obs_dim] may be NULL. */
int armsim_reset(ArmSim* sim, const uint8_t* mask_dev, float* obs_dev, void* stream);

/* Env.step(action) for the whole batch in ONE kernel launch.  Replaces rl_reach_env.py:219-319 (and the push /
 * pick / kuka_reach equivalents): EE servo + IK + teleport + cube/gripper contact + reward + done, fused.
 *   action_dev  f32 [n, 3]  (IK mode)  or f32 [n, 7] joint torques (torque mode)
 *   obs_dev     f32 [n, obs_dim]
 *   reward_dev  f32 [n]
 *   done_dev    u8  [n]
 *   success_dev u8  [n]     reach: is_success; push/pick: info['is_success']; kuka_reach: distance < 0.1
 * Envs already done (auto_reset = 0) are left untouched and report reward 0, done 1. */
int armsim_step(ArmSim* sim, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                uint8_t* success_dev, void* stream);

/* armsim_step plus one more output: final_obs_dev f32 [n, obs_dim] (nullable) receives the observation of THIS step
 * before any in-kernel auto-reset -- what the reference's step() returns on a terminal step (rl_reach_env.py:319)
 * and what a replay buffer must store as next_state; equal to obs_dev for envs that did not terminate. */
int armsim_step_ex(ArmSim* sim, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                   uint8_t* success_dev, float* final_obs_dev, void* stream);

/* armsim_step_ex and armsim_track_episodes (below) of the same step in ONE launch: the rollout's per-step bookkeeping
 * (main.py:202-207, :222-229) rides in the step kernel's epilogue.  IK-teleport mode only. */
int armsim_step_tracked(ArmSim* sim, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                        uint8_t* success_dev, float* final_obs_dev, void* stream);

/* Same step through HOST buffers (what a host-side Env.step sees): the actions cross to the device, the fused
 * launch runs, obs/reward/done/success cross back, and the call returns when they are in the caller's buffers.
 * Arbitrary host pointers are staged through the handle's pinned block; see armsim_host_buffers for the copy-free
 * form.  Used for the end-to-end measurement. */
int armsim_step_host(ArmSim* sim, const float* action_host, float* obs_host, float* reward_host, uint8_t* done_host,
                     uint8_t* success_host);
int armsim_reset_host(ArmSim* sim, const uint8_t* mask_host, float* obs_host);

/* The same call split in two, gym.vector's step_async / step_wait: _async stages the actions and puts the launch in
 * flight, _wait blocks until the results are in the caller's buffers.  A host loop that owns two (or more) handles
 * can overlap one group's launch + PCIe round trip with its own work on the other group's results.  At most one step
 * per handle may be in flight (ARMSIM_E_STATE otherwise).  No counterpart in the reference (single env, synchronous). */
int armsim_step_host_async(ArmSim* sim, const float* action_host);
int armsim_step_host_wait(ArmSim* sim, float* obs_host, float* reward_host, uint8_t* done_host, uint8_t* success_host);

/* Resident step server for the host path (off by default).  idle_us > 0: armsim_step_host / _async stop launching a
 * kernel per step; ONE kernel stays on the GPU, block 0 polls a command word in the handle's pinned block, every block
 * serves the step and rings its doorbell, and the kernel leaves by itself after idle_us microseconds without a command
 * (the next host step starts it again; no step is ever run twice or skipped).  Saves the launch -> first-instruction
 * latency that dominates a launch-per-step host call.  While the server is alive any OTHER entry point that
 * touches this handle's device state first makes it leave (one stream synchronise), and a device-wide synchronisation
 * issued elsewhere in the process (cudaDeviceSynchronize, cudaFree, cudaMalloc) waits up to idle_us for it.
 * idle_us = 0 turns it off.  Zero-copy host path only (n_envs <= 65536), IK-teleport mode. */
int armsim_host_server(ArmSim* sim, int32_t idle_us);

/* The handle's own pinned (page-locked, device-mapped) I/O block: action f32 [n, act_dim], obs f32 [n, obs_dim],
 * reward f32 [n], done u8 [n], success u8 [n].  Passing exactly these pointers to armsim_step_host makes the call
 * copy-free on the host: for n_envs <= 65536 the kernel reads the actions from and writes the results to this block
 * over PCIe itself and rings a doorbell in it (no DMA launches, no stream synchronise).  Any pointer may be NULL. */
int armsim_host_buffers(ArmSim* sim, float** action, float** obs, float** reward, uint8_t** done, uint8_t** success);

/* Rollout bookkeeping of the reference's train loops, on the device (no counterpart kernel in the reference: these are
 * the numpy lines around env.step in main.py).
 *
 * armsim_explore:  action_out = actor_out + noise_std * N(0,1), clipped to [-clip, clip] when clip > 0
 *   (main.py:200 `action + np.random.normal(0, action_bound * opt.gamma)`; main.py:116-117 adds the clip).  Normal draws
 *   come from Philox4x32-10 keyed by the handle's seed with counter (GLOBAL env id, per-env draw count), so a CUDA-graph
 *   replay draws fresh noise every time and the stream does not depend on how envs are sharded over ranks.
 *   actor_out_dev, action_out_dev: f32 [n, action_dim]; may alias.
 * armsim_track_episodes:  episode_return += reward; every env with done != 0 adds {1, success != 0, episode_return} to the
 *   handle's statistics and restarts its return (main.py:202-207 `episode_return += reward`, :203,:222-229 success-rate
 *   bookkeeping).  The return sum is kept in 2^-16 fixed point so it does not depend on the order of the atomics.
 * armsim_episode_stats / armsim_set_episode_stats:  {episodes finished, successes, sum of finished returns}; synchronous. */
int armsim_explore(ArmSim* sim, const float* actor_out_dev, float noise_std, float clip, float* action_out_dev, void* stream);

/* The acting policy of the rollout as ONE launch: action = tanh(fc3(relu(fc2(relu(fc1(obs)))))) * action_bound, then the
 * exploration step of armsim_explore (same Philox stream, same draw counter) when noise_std >= 0; noise_std < 0 returns
 * the bare policy output and leaves the draw counters alone.  Replaces `agent.take_action(state)` + noise of
 * main.py:196-200 for the MLP actors (PolicyNet.forward, algo/TD3/net_mlp.py:29-40; identical in DDPG/DADDPG/DATD3/DARC).
 * The weights are read where PyTorch keeps them: nn.Linear layout, row-major [out, in], fp32, device pointers:
 * w1 [hidden, obs_dim], w2 [hidden, hidden], w3 [action_dim, hidden], biases [out].  hidden must be 256 (the reference's
 * opt.hidden_dim); obs_dim / action_dim are the handle's.  fp32 FFMA throughout (the rollout acts with the policy that is
 * being trained, to fp32 rounding).  obs_dev f32 [n, obs_dim]; action_out_dev f32 [n, action_dim]. */
int armsim_policy_act(ArmSim* sim, const float* obs_dev, const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* w3, const float* b3, int32_t hidden, float action_bound, float noise_std, float clip,
                      float* action_out_dev, void* stream);
int armsim_track_episodes(ArmSim* sim, const float* reward_dev, const uint8_t* done_dev, const uint8_t* success_dev,
                          void* stream);
int armsim_episode_stats(ArmSim* sim, double out[3]);
int armsim_set_episode_stats(ArmSim* sim, const double in[3]);

/* Inject / read back per-env state (parity tests inject q, goal, cube pose; SURVEY 5 "seeding facts").
 * `bytes` must equal n_envs * width * 4.  Synchronous. */
int armsim_set_state(ArmSim* sim, int32_t field, const void* host_src, size_t bytes);
int armsim_get_state(ArmSim* sim, int32_t field, void* host_dst, size_t bytes);

/* Introspection */
int32_t armsim_obs_dim(const ArmSim* sim);
int32_t armsim_action_dim(const ArmSim* sim);
int32_t armsim_n_envs(const ArmSim* sim);
int32_t armsim_mapping(const ArmSim* sim);        /* resolved ARMSIM_MAP_* */
int64_t armsim_launch_count(const ArmSim* sim);   /* kernels launched by this handle so far */
int32_t armsim_abi_version(void);
const char* armsim_last_error(void);

/* Forward kinematics of the handle's chain for a host batch q[n,7] -> pos[n,3], rot[n,9] (row-major), computed
 * on the device by the same routine the step kernel uses.  Synchronous; for tests and for Env shims that need
 * getLinkState-style queries. */
int armsim_fk_host(ArmSim* sim, const float* q_host, int32_t n, float* pos_host, float* rot_host);

/* ------------------------------------------------------------------------------------------------------------------
 * Trajectory replay with HER "future" relabelling, resident in HBM.  Replaces utils/rl_utils.py:91-105 (Trajectory)
 * and :108-199 (ReplayBuffer_Trajectory_reach / _push: add_trajectory, size, sample) for the lockstep batch: every
 * env step appends ONE row for all n envs; an episode becomes sampleable when its terminal step is stored (the
 * reference adds a trajectory at episode end, main.py:129).  All cursors live on the device: store / sample are
 * asynchronous on `stream` and CUDA-graph capturable.
 */
typedef struct ArmReplay ArmReplay;
typedef struct ArmReplayConfig {
  int32_t struct_size;   /* = sizeof(ArmReplayConfig) */
  int32_t n_envs;        /* envs of this shard (rows are [n_envs, ...]) */
  int32_t obs_dim;       /* >= 6: obs[0:3] = achieved position (EE), obs[3:6] = goal slot (rl_utils.py:134,140) */
  int32_t act_dim;
  int32_t window;        /* ring length in lockstep steps; episodes longer than this are dropped */
  int32_t table_cap;     /* trajectory-table slots (the reference's deque capacity, counted in trajectories) */
  int32_t kind;          /* 0 = reach relabelling (rl_utils.py:133-141), 1 = push (:180-188) */
  int32_t device;
  uint64_t seed;         /* Philox key of the sampler */
} ArmReplayConfig;

int armsim_replay_create(const ArmReplayConfig* cfg, ArmReplay** out);
void armsim_replay_destroy(ArmReplay* rep);
/* Start (or restart) every env's episode at the current row with obs0_dev f32 [n, obs_dim] as states[0]
 * (Trajectory(init_state), rl_utils.py:93-94).  Call after Env.reset() and before the first store. */
int armsim_replay_begin(ArmReplay* rep, const float* obs0_dev, void* stream);
/* Append the row of one lockstep Env.step (Trajectory.store_step, rl_utils.py:100-105): action [n,A], reward [n],
 * done u8 [n], final_obs [n,O] (this step's observation before auto-reset, armsim_step_ex) and obs_out [n,O] (the
 * observation the next step starts from).  Envs with done != 0 commit their trajectory (add_trajectory, :112). */
int armsim_replay_store(ArmReplay* rep, const float* action_dev, const float* reward_dev, const uint8_t* done_dev,
                        const float* final_obs_dev, const float* obs_out_dev, void* stream);
/* sample(batch_size, use_her, dis_threshold, her_ratio) (rl_utils.py:119-152 / :165-199): states [B,O], actions
 * [B,A], next_states [B,O], rewards [B], dones f32 0/1 [B].  picks_dev i32 [B,3] (nullable) receives the drawn
 * (table slot, step, goal step or -1; all -1 when nothing is sampleable).  Trajectories are drawn uniformly over the
 * intact ones (those whose rows the ring still holds). */
int armsim_replay_sample(ArmReplay* rep, int32_t batch, int32_t use_her, float dis_threshold, float her_ratio,
                         float* states_dev, float* actions_dev, float* next_states_dev, float* rewards_dev,
                         float* dones_dev, int32_t* picks_dev, void* stream);
/* The same transition builder for EXPLICIT picks (slot, step, goal_step or -1): the deterministic entry the parity
 * tests use against the reference's relabelling. */
int armsim_replay_gather(ArmReplay* rep, int32_t batch, const int32_t* slot_dev, const int32_t* step_dev,
                         const int32_t* goal_step_dev, float dis_threshold, float* states_dev, float* actions_dev,
                         float* next_states_dev, float* rewards_dev, float* dones_dev, void* stream);
/* Synchronous read-backs: info = {rows stored, trajectories committed (size(), rl_utils.py:115), sample calls,
 * sample calls that found no intact trajectory}; the first `count` trajectory-table entries (env, absolute start
 * row, length).  Sampling from an empty replay (the reference raises, rl_utils.py:126) yields a zero-filled batch
 * with done = 1 and bumps info[3]. */
int armsim_replay_info(ArmReplay* rep, int64_t info[4]);
int armsim_replay_table(ArmReplay* rep, int32_t* env_host, int64_t* start_host, int32_t* len_host, int32_t count);
/* Checkpoint / resume: the ring, the trajectory table and every cursor as one opaque blob (only valid for a replay
 * created with the identical ArmReplayConfig).  Synchronous. */
int64_t armsim_replay_state_bytes(ArmReplay* rep);
int armsim_replay_get_state(ArmReplay* rep, void* host_dst, int64_t bytes);
int armsim_replay_set_state(ArmReplay* rep, const void* host_src, int64_t bytes);
const char* armsim_replay_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ARMSIM_H */
