// armsim_kernels.cuh -- the fused Env.step / Env.reset kernels (sm_100a).
//
// Mapping "lane": one CUDA lane per arm, state as struct-of-arrays so every state load/store of a warp is one
// coalesced 128-byte line; the row-major [n,3] actions and [n,obs_dim] observations the caller (PyTorch) wants are
// staged through shared memory so their global accesses are coalesced too.  One launch = one Env.step for the whole
// batch: FK -> workspace clip -> DLS IK (<= 20 iterations) -> teleport -> cube / gripper contact step(s) -> reward /
// done / success -> optional in-kernel auto-reset (Philox) -> obs.
#pragma once
#include "armsim_device.cuh"
#include "aba_device.cuh"
#include "cube_model.cuh"

#ifndef ARMSIM_LANE_BLOCK
#define ARMSIM_LANE_BLOCK 128
#endif
constexpr int LANE_BLOCK = ARMSIM_LANE_BLOCK;
// Builds of the step kernels, picked by grid size in launch_step (each size gets the build measured fastest there):
//   BUILD_LATENCY  grids of at most one block per SM (reach N <= 18944): every warp sits alone on a scheduler slot; takes
//                  the registers it wants (reach 142, push 227, pick 254, no spills: shortest dependent chain) and
//                  prefetches its env lines before griddepcontrol.wait (prefetch_env_lines below)
//   BUILD_WAVE     reach-type tasks, up to one full wave of 4 blocks per SM: capped at 128 registers, no prefetch
//                  (reach N = 32768: 5.1 us; the 142-register build measures 6.0 us there, the 128-register one 3.22
//                  instead of 3.12 us at N = 4096).  Cube tasks use BUILD_LATENCY up to 2 blocks per SM.
//   BUILD_DENSE    multi-wave grids: capped at 80 registers (reach: 6 resident blocks per SM, a few spills, 50 % more
//                  warps to hide latency with: +8 % env-steps/s at N >= 1 M, -7 % at N = 4096) / 168 (cube tasks: 3)
#ifndef ARMSIM_SPARSE_MIN_BLOCKS
#define ARMSIM_SPARSE_MIN_BLOCKS 1
#endif
constexpr int BUILD_LATENCY = 0, BUILD_WAVE = 1, BUILD_DENSE = 2;
constexpr int DENSE_MIN_BLOCKS = 6;
constexpr int WAVE_MIN_BLOCKS_REACH = 4;
constexpr int DENSE_MIN_BLOCKS_CUBE = 3;        // push / pick: 227 / 254 registers by default, <= 168 when dense
constexpr int DENSE_GRID_THRESHOLD = 148 * 4;   // more blocks than one wave of the 128-register reach build
constexpr int DENSE_GRID_THRESHOLD_CUBE = 148 * 2;
constexpr int PREFETCH_MAX_GRID = 148;          // one block per SM: the next launch's blocks are resident while this one runs

template <int TASK>
struct TaskTraits {
  static constexpr int OBS = (TASK == ARMSIM_TASK_REACH) ? 6 : (TASK == ARMSIM_TASK_KUKA_REACH ? 3 : 9);
  static constexpr bool HAS_CUBE = (TASK == ARMSIM_TASK_PUSH || TASK == ARMSIM_TASK_PICK);
};

__device__ __forceinline__ void load_cube(const StatePtrs& S, int n, int e, cube::State& c) {
#pragma unroll
  for (int i = 0; i < 3; ++i) c.pos[i] = S.cube[(0 + i) * n + e];
#pragma unroll
  for (int i = 0; i < 4; ++i) c.quat[i] = S.cube[(3 + i) * n + e];
#pragma unroll
  for (int i = 0; i < 3; ++i) c.v[i] = S.cube[(7 + i) * n + e];
#pragma unroll
  for (int i = 0; i < 3; ++i) c.w[i] = S.cube[(10 + i) * n + e];
}

__device__ __forceinline__ void store_cube(const StatePtrs& S, int n, int e, const cube::State& c) {
#pragma unroll
  for (int i = 0; i < 3; ++i) S.cube[(0 + i) * n + e] = c.pos[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) S.cube[(3 + i) * n + e] = c.quat[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) S.cube[(7 + i) * n + e] = c.v[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) S.cube[(10 + i) * n + e] = c.w[i];
}

template <int TASK>
__device__ __forceinline__ void make_obs(const float (&ee)[3], const float (&goal)[3], const cube::State& cb,
                                         float (&o)[TaskTraits<TASK>::OBS]) {
  o[0] = ee[0]; o[1] = ee[1]; o[2] = ee[2];
  if constexpr (TASK == ARMSIM_TASK_REACH) {
    o[3] = goal[0]; o[4] = goal[1]; o[5] = goal[2];
  } else if constexpr (TaskTraits<TASK>::HAS_CUBE) {
    o[3] = cb.pos[0]; o[4] = cb.pos[1]; o[5] = cb.pos[2];
    o[6] = goal[0]; o[7] = goal[1]; o[8] = goal[2];
  }
}

// Env.reset() for one env (rl_reach_env.py:132-217, rl_push_env.py:145-256, rl_pick_env.py:141-256).  Draws come from
// Philox4x32-10 keyed by (seed, global env id, episode); formulas follow the reference: random.uniform(a,b) = a+(b-a)u.
template <int TASK>
__device__ __forceinline__ void reset_env(const ChainParams& C, const TaskParams& T, const StatePtrs& S, int e,
                                          float (&o)[TaskTraits<TASK>::OBS]) {
  const int n = T.n;
  const unsigned long long gid = T.gid_offset + (unsigned long long)e;
  const uint32_t ep = (uint32_t)S.episode[e];
  float q[NJ], goal[3], u[4];
  cube::State cb;
#pragma unroll
  for (int j = 0; j < NJ; ++j) { q[j] = T.init_q[j]; S.q[j * n + e] = q[j]; }
  if (T.torque_mode) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) S.qd[j * n + e] = 0.f;
  }
  if constexpr (!TaskTraits<TASK>::HAS_CUBE) {
    reset_uniforms(T, gid, ep, 0u, u);
#pragma unroll
    for (int i = 0; i < 3; ++i) goal[i] = __fmaf_rn(T.goal_span[i], u[i], T.goal_lo[i]);
  } else {
    float cx = 0.f, cy = 0.f, cz = 0.f, cyaw = 0.f;
    for (uint32_t attempt = 0; attempt < 1000u; ++attempt) {   // rejection loop rl_push_env.py:195-214
      float v[4];
      reset_uniforms(T, gid, ep, 2u * attempt, u);
      reset_uniforms(T, gid, ep, 2u * attempt + 1u, v);
      cx = __fmaf_rn(T.goal_span[0], u[0], T.goal_lo[0]);
      cy = __fmaf_rn(T.goal_span[1], u[1], T.goal_lo[1]);
      cz = 0.01f;
      cyaw = __fmaf_rn(3.1415925438f, u[2], 1.57f);
      goal[0] = __fmaf_rn(T.goal_span[0], u[3], T.goal_lo[0]);
      goal[1] = __fmaf_rn(T.goal_span[1], v[0], T.goal_lo[1]);
      goal[2] = (TASK == ARMSIM_TASK_PICK) ? __fmaf_rn(T.goal_span[2], v[1], T.goal_lo[2]) : 0.01f;
      const float dx = __fsub_rn(cx, goal[0]), dy = __fsub_rn(cy, goal[1]), dz = __fsub_rn(cz, goal[2]);
      const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      const float d = __fsqrt_rn(d2);
      if (d >= 0.22f && d <= 0.25f) break;
    }
    cube::init(cb, cx, cy, cz, cyaw);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) S.goal[i * n + e] = goal[i];
  S.step[e] = 0;
  S.done[e] = 0;
  S.ik_iters[e] = 0;
  S.episode[e] = (int)(ep + 1u);
  // every episode starts at init_joint_positions: its EE pose is a constant (FK done once on the host, fp64)
  float p[3], R[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = T.init_ee[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = T.init_R[i];
  if constexpr (TaskTraits<TASK>::HAS_CUBE) {
    S.grip[e] = 0.f;
    cube::step<TASK == ARMSIM_TASK_PICK>(cb, p, R, 0.f);   // p.stepSimulation() rl_push_env.py:242
    const float d0 = cb.pos[0] - goal[0], d1 = cb.pos[1] - goal[1], d2 = cb.pos[2] - goal[2];
    S.last_dist[e] = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    store_cube(S, n, e, cb);
  }
  make_obs<TASK>(p, goal, cb, o);
}

// Per-env state of one step, held in registers between the load and the store phase.
template <int TASK>
struct EnvRegs {
  float q[NJ], qd[NJ], goal[3];
  int stepc;
  uint8_t latched;     // done flag latched by a previous step (auto_reset = 0)
  cube::State cb;
  float last_dist, grip;
};

// Issue EVERY global load of the step back to back (one HBM round trip, not one per field).
template <int TASK>
__device__ __forceinline__ void load_env(const TaskParams& T, const StatePtrs& S, int e, EnvRegs<TASK>& E) {
  const int n = T.n;
#pragma unroll
  for (int j = 0; j < NJ; ++j) E.q[j] = S.q[j * n + e];
#pragma unroll
  for (int i = 0; i < 3; ++i) E.goal[i] = S.goal[i * n + e];
  E.stepc = S.step[e];
  E.latched = T.auto_reset ? (uint8_t)0 : S.done[e];
  if constexpr (TaskTraits<TASK>::HAS_CUBE) {
    load_cube(S, n, e, E.cb);
    E.last_dist = S.last_dist[e];
    E.grip = S.grip[e];
  }
}

// Everything of Env.step() after the arm has moved: cube / gripper contact step(s), reward, done, success, state
// write-back, auto-reset, observation.  p, R = EE link frame of the state in E.q.  Returns true when the env was
// re-initialised in place (auto-reset), i.e. the stored q is init_q again.
template <int TASK>
__device__ __forceinline__ bool task_epilogue(const ChainParams& C, const TaskParams& T, const StatePtrs& S, int e, bool live,
                                              EnvRegs<TASK>& E, const float (&p)[3], const float (&R)[9], int its,
                                              float (&o)[TaskTraits<TASK>::OBS], float (&of)[TaskTraits<TASK>::OBS],
                                              float& r, uint8_t& d, uint8_t& su) {
  const int n = T.n;
  float (&q)[NJ] = E.q;
  float (&goal)[3] = E.goal;
  cube::State& cb = E.cb;
  int stepc = E.stepc;
  if (live) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) S.q[j * n + e] = q[j];           // resetJointState :252-257
    S.ik_iters[e] = its;
  }
  stepc += 1;                                                     // :264

  bool term = false, succ = false;
  if constexpr (TASK == ARMSIM_TASK_REACH) {
    const float d0 = p[0] - goal[0], d1 = p[1] - goal[1], d2 = p[2] - goal[2];
    const float dist = fast_sqrt(fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));   // :281
    if (stepc > T.max_steps) { r = -dist * 10.f; term = true; }  // :299-301
    else if (dist < T.reach_dis) { r = 0.f; term = true; succ = true; }  // :303-306
    else { r = -dist * 10.f; }                                    // :307-309
  } else if constexpr (TASK == ARMSIM_TASK_KUKA_REACH) {
    const float d0 = p[0] - goal[0], d1 = p[1] - goal[1], d2 = p[2] - goal[2];
    const float dist = fast_sqrt(fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));
    const bool oob = p[0] < T.ws_lo[0] || p[0] > T.ws_hi[0] || p[1] < T.ws_lo[1] || p[1] > T.ws_hi[1] ||
                     p[2] < T.ws_lo[2] || p[2] > T.ws_hi[2];     // kuka_reach_env.py:276-278
    if (oob) { r = -1.f; term = true; }
    else if (stepc > T.max_steps) { r = -1.f; term = true; }
    else if (dist < T.reach_dis) { r = 10.f; term = true; succ = true; }
    else { r = 0.f; }
  } else {
    constexpr bool PICK = TASK == ARMSIM_TASK_PICK;
    float grip = E.grip;
    cube::step<PICK>(cb, p, R, grip);                             // p.stepSimulation() rl_push_env.py:349
    const float d0 = cb.pos[0] - goal[0], d1 = cb.pos[1] - goal[1], d2 = cb.pos[2] - goal[2];
    const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);       // rl_push_env.py:383 / :393
    float test = dist - E.last_dist;                              // :385
    if (fabsf(test) < 1e-5f) test = 0.01f;                        // :386-387
    if (live) S.last_dist[e] = dist;
    if (stepc > T.max_steps) { r = -dist * 50.f; term = true; }   // :417-419
    else if (dist < 0.05f) { r = 100.f; term = true; }            // :421-423
    else { r = -test * 100.f; }                                   // :424-428
    succ = dist < T.reach_dis;                                    // _is_success :442-445
    make_obs<TASK>(p, goal, cb, of);                              // obs = self._get_obs() (rl_pick_env.py:381) ...
    if constexpr (PICK) {
      // ... taken BEFORE the finger snap + second p.stepSimulation() (rl_pick_env.py:412-417), whose effect the next
      // env step observes
      if (grip < 0.5f && cube::gripper_distance(cb, p, R, grip) < cube::CLOSE_DIST) grip = 1.f;
      cube::step<PICK>(cb, p, R, grip);
      if (live) S.grip[e] = grip;
    }
    if (live) store_cube(S, n, e, cb);
  }
  d = term ? 1 : 0;
  su = succ ? 1 : 0;
  if constexpr (!TaskTraits<TASK>::HAS_CUBE) make_obs<TASK>(p, goal, cb, of);   // the observation of THIS step (gymnasium's final_observation)
#pragma unroll
  for (int k = 0; k < TaskTraits<TASK>::OBS; ++k) o[k] = of[k];
  if (!live) return false;
  S.step[e] = stepc;
  if (term && T.auto_reset) {
    reset_env<TASK>(C, T, S, e, o);                // obs = first observation of the next episode
    return true;
  }
  if (term) S.done[e] = 1;
  return false;
}

// Env.step() for one env on pre-loaded registers (IK-teleport mode, what the reference does).  Returns through
// o / r / d / su; `live` = false lanes (padding of the last warp) compute on a clone of the last env and store nothing.
template <int TASK, int ROBOT>
__device__ __forceinline__ void step_env(const ChainParams& C, const TaskParams& T, const StatePtrs& S, int e, bool live,
                                         EnvRegs<TASK>& E, const float (&a)[3], float (&o)[TaskTraits<TASK>::OBS],
                                         float (&of)[TaskTraits<TASK>::OBS], float& r, uint8_t& d, uint8_t& su) {
  float p[3], R[9];
  const bool frozen = E.latched != 0;               // finished env waiting for reset: report its frozen state
  const int its = servo_core<ROBOT, TASK != ARMSIM_TASK_KUKA_REACH>(C, T, a, frozen, E.q, p, R);
  if (frozen) {
    make_obs<TASK>(p, E.goal, E.cb, o);
#pragma unroll
    for (int k = 0; k < TaskTraits<TASK>::OBS; ++k) of[k] = o[k];
    r = 0.f; d = 1; su = 0;
    return;
  }
  task_epilogue<TASK>(C, T, S, e, live, E, p, R, its, o, of, r, d, su);
}

// Completion doorbells of the host-buffer path (armsim_step_host): when `flags` is non-null the kernel's outputs go
// straight to mapped pinned host memory and EVERY block publishes its own sequence number there once its outputs are
// out, so the host learns of completion by polling gridDim.x consecutive words instead of paying a stream
// synchronise.  No device-wide atomic and one system-scope fence per block: the block barrier orders every thread's
// output stores before thread 0's fence (PTX memory model: the fence is cumulative over what the barrier made
// visible), and the fence orders them before the doorbell store.  The sequence number is a per-block counter in
// device memory, so the launch parameters never change and the whole call replays from an instantiated CUDA graph.
struct HostNotify {
  unsigned int* cta_seq;   // device: [gridDim.x] launches seen by each block
  unsigned int* flags;     // mapped host memory: [gridDim.x]
  unsigned long long* track_stats = nullptr;   // non-null: fold armsim_track_episodes into this launch (armsim_step_tracked)
};

// main.py:202-207, :222-229 for one warp of envs: accumulate the running return, fold finished episodes into
// {episodes, successes, return sum}.  Warp-aggregated; the return sum is an integer (2^-16 fixed point) so the total is
// independent of the order in which warps arrive.  Every lane of the warp must call it (live = false for padding).
__device__ __forceinline__ void track_warp(const StatePtrs& S, int e, bool live, float reward, bool done, bool success,
                                           unsigned long long* __restrict__ stats) {
  bool fin = false, suc = false;
  long long fx = 0;
  if (live) {
    float ret = S.ep_return[e] + reward;
    fin = done;
    if (fin) {
      suc = success;
      fx = llrintf(ret * 65536.0f);
      ret = 0.0f;
    }
    S.ep_return[e] = ret;
  }
  const unsigned mf = __ballot_sync(0xffffffffu, fin);
  if (mf == 0u) return;
  const unsigned ms = __ballot_sync(0xffffffffu, suc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fx += __shfl_xor_sync(0xffffffffu, fx, o);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(stats + 0, (unsigned long long)__popc(mf));
    if (ms) atomicAdd(stats + 1, (unsigned long long)__popc(ms));
    atomicAdd(stats + 2, (unsigned long long)fx);
  }
}

// Programmatic dependent launch (see launch_k in armsim_capi.cu): nothing may be read from global memory before
// pdl_wait(); pdl_release() lets the next PDL-launched kernel of the stream start its own prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Before griddepcontrol.wait a kernel may not READ anything its predecessor writes -- but it may ask L2 to fetch the
// lines it is about to need: a prefetch returns no data to the SM, and L2 is the coherence point, so whatever the
// predecessor still writes lands in the very lines that were fetched.  Each warp prefetches the 128-byte lines of its
// own 32 envs (q[7], goal[3], step, 3 action lines; cube tasks: cube[13], last_dist, grip) while the previous launch
// of the stream is still running; when the state is HBM-cold (a caller cycling through more env batches than L2
// holds: bench.py's pool, multi-wave N) the round trip is over by the time the loads issue.  L2-hot state: no-op.
// Only grids of at most one block per SM take it (gridDim.x <= PREFETCH_MAX_GRID): there the next launch's blocks are
// resident and waiting while this one runs; larger grids start block by block as slots free up, so the prefetch would
// lead its loads by nothing and only add requests (measured: reach N = 32768 5.1 -> 5.8 us with it).

template <int TASK>
__device__ __forceinline__ void prefetch_env_lines(int n, const float* q, const float* goal, const int* step, const float* cube,
                                                       const float* last_dist, const float* grip, const float* action, int wbase,
                                                       int lane) {
  const void* p = nullptr;
  if (lane < 7) p = q + (size_t)lane * n + wbase;
  else if (lane < 10) p = goal + (size_t)(lane - 7) * n + wbase;
  else if (lane == 10) p = step + wbase;
  else if (lane < 14) { if (action != nullptr && (wbase * 3 + (lane - 11) * 32) < n * 3) p = action + (size_t)wbase * 3 + (lane - 11) * 32; }
  else if (TaskTraits<TASK>::HAS_CUBE) {
    if (lane < 27) p = cube + (size_t)(lane - 14) * n + wbase;
    else if (lane == 27) p = last_dist + wbase;
    else if (lane == 28) p = grip + wbase;
  }
  if (p != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ void notify_host(const HostNotify& H) {
  if (H.flags == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int seq = H.cta_seq[blockIdx.x] + 1u;
    H.cta_seq[blockIdx.x] = seq;
    __threadfence_system();
    *(volatile unsigned int*)(H.flags + blockIdx.x) = seq;
  }
}

// Env.step() of the 32 envs one warp owns.  Each warp is self-contained: it stages its [32,3] action rows and
// [32,OBS] observation rows through its own slice of shared memory (row-major caller layout <-> one-value-per-lane),
// ordered by __syncwarp only -- no block-wide barrier on the device path, so warps never wait for each other's HBM
// latency.  COHERENT_ACTIONS: the actions are re-read from mapped host memory by a RESIDENT kernel (step_server_kernel)
// and must not come out of a non-coherent cache.
template <int TASK, int ROBOT, bool COHERENT_ACTIONS>
__device__ __forceinline__ void step_warp(const ChainParams& C, const TaskParams& T, const StatePtrs& S,
                                          const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
                                          uint8_t* __restrict__ done, uint8_t* __restrict__ success, float* __restrict__ final_obs,
                                          const HostNotify& H, float* st, int wbase, int lane) {
  constexpr int OD = TaskTraits<TASK>::OBS;
  const int cnt = min(32, T.n - wbase);
  const bool live = lane < cnt;
  const int e = wbase + min(lane, cnt - 1);

  float araw[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = k * 32 + lane;
    if (COHERENT_ACTIONS) araw[k] = i < cnt * 3 ? *(const volatile float*)(action + (size_t)wbase * 3 + i) : 0.f;
    else araw[k] = i < cnt * 3 ? __ldg(action + (size_t)wbase * 3 + i) : 0.f;
  }
  EnvRegs<TASK> E;
  load_env<TASK>(T, S, e, E);
#pragma unroll
  for (int k = 0; k < 3; ++k) st[k * 32 + lane] = araw[k];
  __syncwarp();
  const int al = min(lane, cnt - 1) * 3;
  const float a[3] = {st[al], st[al + 1], st[al + 2]};
  __syncwarp();

  float o[OD], of[OD], r;
  uint8_t d, su;
  step_env<TASK, ROBOT>(C, T, S, e, live, E, a, o, of, r, d, su);
  if (live) {
    reward[e] = r;
    done[e] = d;
    success[e] = su;
  }
  if (H.track_stats != nullptr) track_warp(S, e, live, r, d != 0, su != 0, H.track_stats);
#pragma unroll
  for (int k = 0; k < OD; ++k) st[lane * OD + k] = o[k];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < OD; ++k) {
    const int i = k * 32 + lane;
    if (i < cnt * OD) obs[(size_t)wbase * OD + i] = st[i];
  }
  if (final_obs != nullptr) {        // second pass through the same staging slice
    __syncwarp();
#pragma unroll
    for (int k = 0; k < OD; ++k) st[lane * OD + k] = of[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < OD; ++k) {
      const int i = k * 32 + lane;
      if (i < cnt * OD) final_obs[(size_t)wbase * OD + i] = st[i];
    }
  }
}

// One launch = Env.step() of the whole batch.
template <int TASK, int ROBOT, int BUILD = BUILD_LATENCY>
__global__ void __launch_bounds__(LANE_BLOCK, BUILD == BUILD_DENSE ? (TaskTraits<TASK>::HAS_CUBE ? DENSE_MIN_BLOCKS_CUBE : DENSE_MIN_BLOCKS)
                                              : (BUILD == BUILD_WAVE ? WAVE_MIN_BLOCKS_REACH : ARMSIM_SPARSE_MIN_BLOCKS))
step_lane_kernel(const __grid_constant__ ChainParams C, const __grid_constant__ TaskParams T, const StatePtrs S,
                 const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
                 uint8_t* __restrict__ done, uint8_t* __restrict__ success, float* __restrict__ final_obs,
                 const HostNotify H) {
  constexpr int OD = TaskTraits<TASK>::OBS;
  constexpr int STAGE = 32 * (OD > 3 ? OD : 3);
  __shared__ float s_io[LANE_BLOCK / 32][STAGE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wbase = blockIdx.x * LANE_BLOCK + warp * 32;
  if (BUILD == BUILD_LATENCY && gridDim.x <= PREFETCH_MAX_GRID && wbase < T.n)
    prefetch_env_lines<TASK>(T.n, S.q, S.goal, S.step, S.cube, S.last_dist, S.grip, H.flags == nullptr ? action : nullptr, wbase,
                             lane);   // (host path: actions sit in mapped host memory)
  pdl_wait();
  pdl_release();
  if (wbase < T.n) step_warp<TASK, ROBOT, false>(C, T, S, action, obs, reward, done, success, final_obs, H, s_io[warp], wbase, lane);
  notify_host(H);
}

// The resident form of the host step (armsim_host_server): the same Env.step(), but the kernel stays on the GPU and
// serves one step per command instead of being launched per step -- a launch-per-step host call spends ~12 of its ~18 us
// between cudaGraphLaunch and the kernel's first instruction.  Protocol (tools/micro/e2e_breakdown.cu measured the
// alternatives: 32 blocks polling host memory themselves take 98 us per round trip, one poller + a device relay 7):
//   host:    writes the actions into the pinned block, then the step's sequence number into the mapped word `cmd`
//   block 0: thread 0 polls `cmd` over PCIe and republishes it in the device word `relay`; the other blocks poll that
//   blocks:  a block runs step number cta_seq[b] + 1 when the relayed number has reached it, rings its doorbell as in the
//            launched path (notify_host: the host waits for gridDim.x doorbells), and goes back to polling
//   exit:    SERVER_STOP in `cmd`, or no command for idle_ns (block 0 then relays the stop) -- the kernel can never
//            outlive its idle time-out, and the host relaunches it on the next step (cta_seq persists, so no step runs twice)
struct ServerCtl {
  const volatile unsigned int* cmd;     // mapped host memory
  volatile unsigned int* relay;         // device memory
  unsigned long long idle_ns;
};
constexpr unsigned int SERVER_STOP = 0xffffffffu;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int TASK, int ROBOT>
__global__ void __launch_bounds__(LANE_BLOCK, 1)
step_server_kernel(const __grid_constant__ ChainParams C, const __grid_constant__ TaskParams T, const StatePtrs S,
                   const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
                   uint8_t* __restrict__ done, uint8_t* __restrict__ success, const HostNotify H, const ServerCtl ctl) {
  constexpr int OD = TaskTraits<TASK>::OBS;
  constexpr int STAGE = 32 * (OD > 3 ? OD : 3);
  __shared__ float s_io[LANE_BLOCK / 32][STAGE];
  __shared__ unsigned int s_cmd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wbase = blockIdx.x * LANE_BLOCK + warp * 32;
  for (;;) {
    if (threadIdx.x == 0) {
      const unsigned int have = H.cta_seq[blockIdx.x];                    // steps this block has completed
      const unsigned long long t0 = global_ns();
      unsigned int v;
      if (blockIdx.x == 0) {
        for (unsigned int polls = 0;; ++polls) {
          v = *ctl.cmd;
          if (v == SERVER_STOP || (int)(v - have) > 0) break;
          if ((polls & 15u) == 15u && (global_ns() - t0 > ctl.idle_ns || polls > 50000000u)) { v = SERVER_STOP; break; }   // (second bound: belt and braces)
        }
        *ctl.relay = v;
      } else {
        for (unsigned int polls = 0;; ++polls) {
          v = *ctl.relay;
          if (v == SERVER_STOP || (int)(v - have) > 0) break;
          if ((polls & 63u) == 63u && (global_ns() - t0 > 2ull * ctl.idle_ns + 1000000ull || polls > 200000000u)) { v = SERVER_STOP; break; }   // (block 0 relays its own time-out first)
        }
      }
      s_cmd = v;
    }
    __syncthreads();
    if (s_cmd == SERVER_STOP) return;
    if (wbase < T.n) step_warp<TASK, ROBOT, true>(C, T, S, action, obs, reward, done, success, nullptr, H, s_io[warp], wbase, lane);
    notify_host(H);      // barrier (every thread has read s_cmd by now), then thread 0: cta_seq[b] += 1, system fence, doorbell
  }
}

// Torque mode: one launch = one dynamics step of the whole batch.  action [n,7] joint torques; obs [n, OBS+14] = the
// task observation followed by q[7], qd[7].  Same warp-local staging as the IK kernel.
template <int TASK, int ROBOT>
__global__ void __launch_bounds__(LANE_BLOCK)
step_torque_kernel(const __grid_constant__ ChainParams C, const __grid_constant__ TaskParams T,
                   const __grid_constant__ DynParams Dn, const StatePtrs S, const float* __restrict__ action,
                   float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done,
                   uint8_t* __restrict__ success, float* __restrict__ final_obs, const HostNotify H) {
  constexpr int OB = TaskTraits<TASK>::OBS;
  constexpr int OD = OB + 2 * NJ;
  __shared__ float s_io[LANE_BLOCK / 32][32 * OD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wbase = blockIdx.x * LANE_BLOCK + warp * 32;
  if (gridDim.x <= PREFETCH_MAX_GRID && wbase < T.n) {   // same pre-wait L2 prefetch as the IK kernel: q[7], qd[7], goal[3], step, the 7 torque lines
    const void* pf = nullptr;
    if (lane < 7) pf = S.q + (size_t)lane * T.n + wbase;
    else if (lane < 14) pf = S.qd + (size_t)(lane - 7) * T.n + wbase;
    else if (lane < 17) pf = S.goal + (size_t)(lane - 14) * T.n + wbase;
    else if (lane == 17) pf = S.step + wbase;
    else if (lane < 25 && H.flags == nullptr && (wbase * NJ + (lane - 18) * 32) < T.n * NJ) pf = action + (size_t)wbase * NJ + (lane - 18) * 32;
    if (pf != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
  }
  pdl_wait();
  pdl_release();
  if (wbase < T.n) {
    const int n = T.n;
    const int cnt = min(32, n - wbase);
    const bool live = lane < cnt;
    const int e = wbase + min(lane, cnt - 1);
    float* st = s_io[warp];

    float araw[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int i = k * 32 + lane;
      araw[k] = i < cnt * NJ ? __ldg(action + (size_t)wbase * NJ + i) : 0.f;
    }
    EnvRegs<TASK> E;
    load_env<TASK>(T, S, e, E);
#pragma unroll
    for (int j = 0; j < NJ; ++j) E.qd[j] = S.qd[j * n + e];
#pragma unroll
    for (int k = 0; k < NJ; ++k) st[k * 32 + lane] = araw[k];
    __syncwarp();
    float cmd[NJ];
    const int al = min(lane, cnt - 1) * NJ;
#pragma unroll
    for (int k = 0; k < NJ; ++k) cmd[k] = st[al + k];
    __syncwarp();

    float o[OB], of[OB], r = 0.f, p[3], R[9], P[NJ][3], Z[NJ][3];
    uint8_t d = 1, su = 0;
    bool was_reset = false;
    const bool frozen = E.latched != 0;
    if (!frozen) aba::torque_step(C, Dn, cmd, E.q, E.qd);
    RobotFK<ROBOT>::template run<false>(C, E.q, p, R, P, Z);
    if (frozen) {
      make_obs<TASK>(p, E.goal, E.cb, o);
#pragma unroll
      for (int k = 0; k < OB; ++k) of[k] = o[k];
    } else {
      if (live) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) S.qd[j * n + e] = E.qd[j];
      }
      was_reset = task_epilogue<TASK>(C, T, S, e, live, E, p, R, 0, o, of, r, d, su);
    }
    if (live) {
      reward[e] = r;
      done[e] = d;
      success[e] = su;
    }
#pragma unroll
    for (int k = 0; k < OB; ++k) st[lane * OD + k] = o[k];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      st[lane * OD + OB + j] = was_reset ? T.init_q[j] : E.q[j];
      st[lane * OD + OB + NJ + j] = was_reset ? 0.f : E.qd[j];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < OD; ++k) {
      const int i = k * 32 + lane;
      if (i < cnt * OD) obs[(size_t)wbase * OD + i] = st[i];
    }
    if (final_obs != nullptr) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < OB; ++k) st[lane * OD + k] = of[k];
#pragma unroll
      for (int j = 0; j < NJ; ++j) { st[lane * OD + OB + j] = E.q[j]; st[lane * OD + OB + NJ + j] = E.qd[j]; }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < OD; ++k) {
        const int i = k * 32 + lane;
        if (i < cnt * OD) final_obs[(size_t)wbase * OD + i] = st[i];
      }
    }
  }
  notify_host(H);
}

// ---------------------------------------------------------------------------------------------- rollout bookkeeping
// main.py:200 (and :116-117 with the clip) for the whole batch: action = actor output + noise_std * N(0,1).
// One lane per env; Philox4x32-10 counter = (global env id, draws so far of this env, block of 4 normals), key = seed
// with a domain tag, Box-Muller on the 4 words.
__global__ void __launch_bounds__(LANE_BLOCK)
explore_kernel(const __grid_constant__ TaskParams T, const StatePtrs S, int act_dim, const float* __restrict__ actor_out,
               float noise_std, float clip, float* __restrict__ action_out) {
  const int e = blockIdx.x * LANE_BLOCK + threadIdx.x;
  if (e >= T.n) return;
  const unsigned long long gid = T.gid_offset + (unsigned long long)e;
  const unsigned int draw = S.explore_count[e];
  S.explore_count[e] = draw + 1u;
  for (int k0 = 0; k0 < act_dim; k0 += 4) {
    float z[4];
    explore_normals(T, gid, draw, (uint32_t)(k0 >> 2), z);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k0 + k < act_dim) {
        float a = fmaf(noise_std, z[k], actor_out[(size_t)e * act_dim + k0 + k]);
        if (clip > 0.0f) a = fminf(fmaxf(a, -clip), clip);
        action_out[(size_t)e * act_dim + k0 + k] = a;
      }
    }
  }
}

// armsim_track_episodes as its own launch (track_warp above; armsim_step_tracked folds it into the step)
__global__ void __launch_bounds__(LANE_BLOCK)
track_episodes_kernel(int n, const StatePtrs S, const float* __restrict__ reward, const uint8_t* __restrict__ done,
                      const uint8_t* __restrict__ success, unsigned long long* __restrict__ stats) {
  const int e = blockIdx.x * LANE_BLOCK + threadIdx.x;
  const bool live = e < n;
  track_warp(S, e, live, live ? reward[e] : 0.f, live && done[e] != 0, live && success[e] != 0, stats);
}

template <int TASK>
__global__ void __launch_bounds__(LANE_BLOCK)
reset_lane_kernel(const __grid_constant__ ChainParams C, const __grid_constant__ TaskParams T, const StatePtrs S,
                  const uint8_t* __restrict__ mask, float* __restrict__ obs) {
  constexpr int OD = TaskTraits<TASK>::OBS;
  const int e = blockIdx.x * LANE_BLOCK + threadIdx.x;
  if (e >= T.n) return;
  if (mask && !mask[e]) return;
  float o[OD];
  reset_env<TASK>(C, T, S, e, o);
  if (obs) {
    const int stride = T.obs_dim;     // OD, or OD + 14 in torque mode
#pragma unroll
    for (int k = 0; k < OD; ++k) obs[(size_t)e * stride + k] = o[k];
    if (T.torque_mode) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) { obs[(size_t)e * stride + OD + j] = T.init_q[j]; obs[(size_t)e * stride + OD + NJ + j] = 0.f; }
    }
  }
}

// FK of a host-supplied batch (armsim_fk_host): q [n,7] row-major -> pos [n,3], rot [n,9]
template <int ROBOT>
__global__ void fk_kernel(const __grid_constant__ ChainParams C, int n, const float* __restrict__ qin, float* __restrict__ pos,
                          float* __restrict__ rot) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float q[NJ], p[3], R[9], P[NJ][3], Z[NJ][3];
#pragma unroll
  for (int j = 0; j < NJ; ++j) q[j] = qin[(size_t)e * NJ + j];
  RobotFK<ROBOT>::template run<false>(C, q, p, R, P, Z);
#pragma unroll
  for (int i = 0; i < 3; ++i) pos[(size_t)e * 3 + i] = p[i];
  if (rot) {
#pragma unroll
    for (int i = 0; i < 9; ++i) rot[(size_t)e * 9 + i] = R[i];
  }
}
