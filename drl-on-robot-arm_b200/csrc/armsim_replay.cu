// armsim_replay.cu -- device-resident trajectory replay with HER "future" relabelling (C-ABI: include/armsim.h,
// section "trajectory replay").  Replaces the reference's utils/rl_utils.py:91-199 (Trajectory +
// ReplayBuffer_Trajectory_reach / _push), whose sample() is a Python loop x256 over a deque with O(len) indexing.
//
// Layout (HBM, one allocation): a ring of W lockstep rows; row k holds, for every env of the shard,
//     act[k][n][A]  rew[k][n]  done[k][n]  fobs[k][n][O]  (this step's observation, pre-auto-reset = next_state)
//     oobs[k][n][O] (the observation the NEXT step starts from = post-reset obs)
// all row-major exactly as the step kernel emits them, so a store is five contiguous, fully coalesced copies.
// states[i] of a trajectory that started at absolute step s is oobs[s-1+i] for i < L and fobs[s+L-1] for i = L.
// Finished episodes are appended to a trajectory table (env, start, length) through an atomic cursor; sampling draws
// uniformly over the table (the reference draws uniformly over trajectories, then over steps, rl_utils.py:126-127).
// Every counter (absolute step, table cursor, sample call number) lives on the device so that store / sample can be
// captured in a CUDA graph and replayed.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "armsim.h"
#include "armsim_device.cuh"

static thread_local char r_err[256] = "";
static int rfail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(r_err, sizeof(r_err), fmt, ap);
  va_end(ap);
  return code;
}
#define RCU(call)                                                                                    \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) return rfail(ARMSIM_E_CUDA, "%s: %s", #call, cudaGetErrorString(_e));     \
  } while (0)

struct ReplayDev {
  int n, O, A, W, cap, kind;
  float *act, *rew, *fobs, *oobs;
  uint8_t* done;
  long long* ep_start;          // [n] absolute step at which the env's current episode started
  int* t_env;                   // [cap]
  long long* t_start;           // [cap]
  int* t_len;                   // [cap]
  unsigned long long* counters; // [0] absolute step (rows written), [1] trajectories appended, [2] sample calls, [3] sample calls on an empty table
  unsigned int* blocks_done;
  uint32_t seed_lo, seed_hi;
};

struct ArmReplay {
  ReplayDev d{};
  void* block = nullptr;
  size_t bytes = 0;
  int device = 0;
};

struct ReplayDeviceGuard {       // launch on the replay's device whatever the caller's current device is
  int prev = -1;
  bool switched = false;
  explicit ReplayDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~ReplayDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

constexpr int RB = 256;

// one lockstep row: five coalesced copies by the whole grid; the LAST block to finish then commits the trajectories of
// the envs that terminated in this row IN ENV ORDER (block-wide prefix sum over the done flags -- deterministic table
// slots, so a resumed run samples the same transitions) and advances the absolute step counter.
__global__ void __launch_bounds__(RB)
replay_store_kernel(const ReplayDev D, const float* __restrict__ action, const float* __restrict__ reward,
                    const uint8_t* __restrict__ done, const float* __restrict__ final_obs, const float* __restrict__ obs_out) {
  __shared__ unsigned int s_warp[RB / 32];
  __shared__ bool s_last;
  const unsigned long long now = D.counters[0];
  const size_t row = (size_t)(now % (unsigned long long)D.W);
  const int n = D.n;
  const int tid = blockIdx.x * RB + threadIdx.x, nth = gridDim.x * RB;
  float* fo = D.fobs + row * n * D.O;
  float* oo = D.oobs + row * n * D.O;
  for (int i = tid; i < n * D.O; i += nth) { fo[i] = final_obs[i]; oo[i] = obs_out[i]; }
  float* ac = D.act + row * n * D.A;
  for (int i = tid; i < n * D.A; i += nth) ac[i] = action[i];
  for (int e = tid; e < n; e += nth) {
    D.rew[row * n + e] = reward[e];
    D.done[row * n + e] = done[e];
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(D.blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ONE pass: thread t owns the env range [t*chunk, (t+1)*chunk); it counts its commits, the block does a single
  // exclusive scan over the 256 counts, and each thread then writes its commits to consecutive slots -- env order is
  // preserved (deterministic table), with two block barriers in total instead of two per 256 envs.
  const unsigned long long base = D.counters[1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = (n + RB - 1) / RB;
  const int lo = min(n, (int)threadIdx.x * chunk), hi = min(n, lo + chunk);
  unsigned int mine = 0;
  for (int e = lo; e < hi; ++e) {
    if (done[e] != 0) {
      const long long L = (long long)now - D.ep_start[e] + 1;
      mine += (L >= 1 && L < (long long)D.W) ? 1u : 0u;       // an episode longer than the ring cannot be replayed
    }
  }
  unsigned int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned int before = incl - mine, total = 0;
  for (int w = 0; w < RB / 32; ++w) {
    const unsigned int c = s_warp[w];
    if (w < warp) before += c;
    total += c;
  }
  {
    unsigned int k = 0;
    for (int e = lo; e < hi; ++e) {
      if (done[e] != 0) {
        const long long st = D.ep_start[e];
        const long long L = (long long)now - st + 1;
        if (L >= 1 && L < (long long)D.W) {
          const unsigned long long slot = (base + before + k) % (unsigned long long)D.cap;
          D.t_env[slot] = e;
          D.t_start[slot] = st;
          D.t_len[slot] = (int)L;
          ++k;
        }
        D.ep_start[e] = (long long)now + 1;
      }
    }
  }
  if (threadIdx.x == 0) {
    D.counters[1] = base + total;
    *D.blocks_done = 0;
    __threadfence();
    D.counters[0] = now + 1;
  }
}

// (re)start every env's episode at the current absolute step with obs0 as states[0]
__global__ void replay_begin_kernel(const ReplayDev D, const float* __restrict__ obs0) {
  const unsigned long long now = D.counters[0];
  // states[0] of an episode starting at `now` is oobs[now - 1]
  const size_t row = (size_t)((now + (unsigned long long)D.W - 1) % (unsigned long long)D.W);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  float* oo = D.oobs + row * D.n * D.O;
  for (int i = tid; i < D.n * D.O; i += nth) oo[i] = obs0[i];
  for (int e = tid; e < D.n; e += nth) D.ep_start[e] = (long long)now;
}

__device__ __forceinline__ const float* traj_state(const ReplayDev& D, int env, long long start, int len, int i) {
  // states[i], i in [0, len]
  if (i < len) {
    const size_t row = (size_t)(((start - 1 + i) % D.W + D.W) % D.W);
    return D.oobs + (row * D.n + env) * D.O;
  }
  const size_t row = (size_t)((start + len - 1) % D.W);
  return D.fobs + (row * D.n + env) * D.O;
}

// The transition of (trajectory slot, step, goal_step) with the reference's relabelling (goal_step < 0: no HER).
//   reach (rl_utils.py:133-141): goal = states[goal_step][:3]; dis = |next[:3] - goal|; reward = dis > thr ? -0.1 : 1;
//     done = dis <= thr; state = [state[:3], goal]; next = [next[:3], goal]
//   push  (rl_utils.py:180-188): state = [state[:3], goal, state[6:10]]; next = [next[:3], goal, state[6:10]]  (sic: the
//     cube slot is overwritten by the EE-derived goal and next keeps STATE's target slot -- reference quirk, kept)
__device__ __forceinline__ void emit_transition(const ReplayDev& D, int slot, int step, int goal_step, float thr, int b,
                                                float* __restrict__ states, float* __restrict__ actions,
                                                float* __restrict__ next_states, float* __restrict__ rewards,
                                                float* __restrict__ dones) {
  const int env = D.t_env[slot], len = D.t_len[slot];
  const long long start = D.t_start[slot];
  const float* s0 = traj_state(D, env, start, len, step);
  const float* s1 = traj_state(D, env, start, len, step + 1);
  const size_t row = (size_t)((start + step) % D.W);
  const float* a = D.act + (row * D.n + env) * D.A;
  for (int k = 0; k < D.A; ++k) actions[(size_t)b * D.A + k] = a[k];
  float r = D.rew[row * D.n + env];
  float dn = D.done[row * D.n + env] ? 1.f : 0.f;
  float* so = states + (size_t)b * D.O;
  float* no = next_states + (size_t)b * D.O;
  for (int k = 0; k < D.O; ++k) { so[k] = s0[k]; no[k] = s1[k]; }
  if (goal_step >= 0) {
    const float* g = traj_state(D, env, start, len, goal_step);
    const float g0 = g[0], g1 = g[1], g2 = g[2];
    const float d0 = __fsub_rn(s1[0], g0), d1 = __fsub_rn(s1[1], g1), d2 = __fsub_rn(s1[2], g2);
    // np.sqrt(np.sum(np.square(.))) on float32 data: products and the left-to-right sum are rounded to f32
    const float dis = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    const bool far = dis > thr;
    r = far ? -0.1f : 1.0f;
    dn = far ? 0.f : 1.f;
    so[3] = g0; so[4] = g1; so[5] = g2;
    no[3] = g0; no[4] = g1; no[5] = g2;
    if (D.kind == 1) {
      for (int k = 6; k < D.O; ++k) no[k] = s0[k];
    }
  }
  rewards[b] = r;
  dones[b] = dn;
}

__global__ void replay_gather_kernel(const ReplayDev D, int batch, const int* __restrict__ slot, const int* __restrict__ step,
                                     const int* __restrict__ goal_step, float thr, float* states, float* actions,
                                     float* next_states, float* rewards, float* dones) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  emit_transition(D, slot[b], step[b], goal_step[b], thr, b, states, actions, next_states, rewards, dones);
}

// ReplayBuffer_Trajectory_*.sample (rl_utils.py:119-152): per sample -- uniform trajectory, uniform step, with
// probability her_ratio a uniform FUTURE step in (step, len] as the goal.  Draws: Philox4x32-10, counter =
// (sample id, attempt, call number), key = seed.
// Which trajectories can be drawn: the table is a ring of `cap` entries in commit order, i.e. sorted by the row in which
// the episode ENDED; an entry is intact while the row holding its states[0] (start - 1) has not been overwritten.  Every
// entry that ended before the ring's oldest row is gone, so a binary search over the end rows gives the first slot
// `lo` that can still be intact and candidates are drawn uniformly from [lo, ntraj): inside that range only episodes
// that straddle the ring's oldest row (at most one per env) are rejected and redrawn, which keeps the draw uniform
// over the intact trajectories like the reference's random.sample(self.buffer, 1) (rl_utils.py:126).  After 16 failed
// attempts (probability <= (episode length / window)^16) a sample falls back to a uniformly drawn trajectory of the
// newest row's commits, which are always intact.
// An EMPTY table (nothing committed yet, or nothing intact) has no transition to give: the reference raises
// (random.sample on an empty deque); here the batch is zero-filled with done = 1 and counters[3] counts the call, so
// a host that did not gate on size() can see it (armsim_replay_info) -- VectorTrainer gates on the minimum local size.
__global__ void replay_sample_kernel(const ReplayDev D, int batch, int use_her, float thr, float her_ratio, float* states,
                                     float* actions, float* next_states, float* rewards, float* dones, int* picks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long call = D.counters[2];
  const unsigned long long now = D.counters[0], ntraj = D.counters[1];
  const unsigned long long cap = (unsigned long long)D.cap;
  const unsigned long long first = ntraj > cap ? ntraj - cap : 0ull;     // oldest logical index still in the table
  // first logical index whose episode ended at or after the ring's oldest row (end rows are non-decreasing)
  const long long oldest_row = (long long)now - (long long)D.W;          // rows >= oldest_row are in the ring
  unsigned long long lo = first, hi = ntraj;
  while (lo < hi) {
    const unsigned long long mid = (lo + hi) >> 1;
    const size_t s = (size_t)(mid % cap);
    const long long end = D.t_start[s] + (long long)D.t_len[s] - 1;
    if (end < oldest_row) lo = mid + 1; else hi = mid;
  }
  const unsigned long long span = ntraj - lo;
  const bool empty = span == 0ull;
  if (b < batch) {
    if (empty) {
      for (int k = 0; k < D.O; ++k) { states[(size_t)b * D.O + k] = 0.f; next_states[(size_t)b * D.O + k] = 0.f; }
      for (int k = 0; k < D.A; ++k) actions[(size_t)b * D.A + k] = 0.f;
      rewards[b] = 0.f;
      dones[b] = 1.f;
      if (picks) { picks[3 * b] = -1; picks[3 * b + 1] = -1; picks[3 * b + 2] = -1; }
    } else {
      int slot = -1;
      uint32_t rnd[4] = {0, 0, 0, 0};
      for (uint32_t attempt = 0; attempt < 16u && slot < 0; ++attempt) {
        philox4x32_10((uint32_t)b, attempt, (uint32_t)call, (uint32_t)(call >> 32), D.seed_lo, D.seed_hi, rnd);
        const int cand = (int)((lo + (((unsigned long long)rnd[0] * span) >> 32)) % cap);
        // states[0] lives in row start-1: intact while start - 1 >= now - W
        if (D.t_start[cand] - 1 >= oldest_row && D.t_len[cand] > 0) slot = cand;
      }
      if (slot < 0) {
        // the commits of the newest row: walk back from ntraj-1 while the end row is the same (bounded by n)
        const size_t last = (size_t)((ntraj - 1) % cap);
        const long long end_new = D.t_start[last] + (long long)D.t_len[last] - 1;
        unsigned long long cnt = 1;
        while (cnt < span && cnt < (unsigned long long)D.n) {
          const size_t s = (size_t)((ntraj - 1 - cnt) % cap);
          if (D.t_start[s] + (long long)D.t_len[s] - 1 != end_new) break;
          ++cnt;
        }
        slot = (int)((ntraj - 1 - (((unsigned long long)rnd[0] * cnt) >> 32)) % cap);
      }
      const int len = D.t_len[slot];
      const int step = (int)(((unsigned long long)rnd[1] * (unsigned long long)len) >> 32);        // randint(len)
      int goal = -1;
      const float coin = (float)(rnd[2] >> 8) * 5.9604644775390625e-08f;
      if (use_her && coin <= her_ratio)                                                            // uniform() <= her_ratio
        goal = step + 1 + (int)(((unsigned long long)rnd[3] * (unsigned long long)(len - step)) >> 32);  // randint(step+1, len+1)
      emit_transition(D, slot, step, goal, thr, b, states, actions, next_states, rewards, dones);
      if (picks) { picks[3 * b] = slot; picks[3 * b + 1] = step; picks[3 * b + 2] = goal; }
    }
  }
  // the last block bumps the call counter so a replayed CUDA graph draws fresh numbers
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(D.blocks_done + 1, 1u) == gridDim.x - 1) {
    D.blocks_done[1] = 0;
    if (empty) D.counters[3] += 1ull;
    __threadfence();
    D.counters[2] = call + 1;
  }
}

extern "C" {

const char* armsim_replay_last_error(void) { return r_err; }

int armsim_replay_create(const ArmReplayConfig* cfg, ArmReplay** out) {
  if (!cfg || !out) return rfail(ARMSIM_E_INVALID, "armsim_replay_create: null argument");
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(ArmReplayConfig)) return rfail(ARMSIM_E_INVALID, "armsim_replay_create: struct_size mismatch");
  if (cfg->n_envs <= 0 || cfg->obs_dim < 6 || cfg->act_dim <= 0 || cfg->window < 2 || cfg->table_cap <= 0 || cfg->kind < 0 || cfg->kind > 1)
    return rfail(ARMSIM_E_INVALID, "armsim_replay_create: bad shape (obs_dim must be >= 6: [ee, goal, ...])");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return rfail(ARMSIM_E_CUDA, "armsim_replay_create: no usable CUDA device");
  if (cfg->device < 0 || cfg->device >= ndev) return rfail(ARMSIM_E_INVALID, "armsim_replay_create: bad device");
  RCU(cudaSetDevice(cfg->device));
  ArmReplay* r = new (std::nothrow) ArmReplay();
  if (!r) return rfail(ARMSIM_E_NOMEM, "armsim_replay_create: host allocation failed");
  r->device = cfg->device;
  ReplayDev& d = r->d;
  d.n = cfg->n_envs; d.O = cfg->obs_dim; d.A = cfg->act_dim; d.W = cfg->window; d.cap = cfg->table_cap; d.kind = cfg->kind;
  d.seed_lo = (uint32_t)cfg->seed; d.seed_hi = (uint32_t)(cfg->seed >> 32);
  auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t n = d.n, W = d.W, cap = d.cap;
  const size_t b_act = pad(W * n * d.A * 4), b_rew = pad(W * n * 4), b_done = pad(W * n), b_obs = pad(W * n * d.O * 4);
  const size_t b_eps = pad(n * 8), b_tenv = pad(cap * 4), b_tst = pad(cap * 8), b_tlen = pad(cap * 4), b_cnt = 256;
  const size_t total = b_act + b_rew + b_done + 2 * b_obs + b_eps + b_tenv + b_tst + b_tlen + b_cnt;
  if (cudaMalloc(&r->block, total) != cudaSuccess) {
    cudaGetLastError();
    delete r;
    return rfail(ARMSIM_E_NOMEM, "armsim_replay_create: cudaMalloc of %zu bytes failed", total);
  }
  cudaMemset(r->block, 0, total);
  r->bytes = total;
  char* p = (char*)r->block;
  d.act = (float*)p; p += b_act;
  d.rew = (float*)p; p += b_rew;
  d.done = (uint8_t*)p; p += b_done;
  d.fobs = (float*)p; p += b_obs;
  d.oobs = (float*)p; p += b_obs;
  d.ep_start = (long long*)p; p += b_eps;
  d.t_env = (int*)p; p += b_tenv;
  d.t_start = (long long*)p; p += b_tst;
  d.t_len = (int*)p; p += b_tlen;
  d.counters = (unsigned long long*)p;
  d.blocks_done = (unsigned int*)(p + 64);
  *out = r;
  return ARMSIM_OK;
}

void armsim_replay_destroy(ArmReplay* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  cudaDeviceSynchronize();
  if (r->block) cudaFree(r->block);
  delete r;
}

int armsim_replay_begin(ArmReplay* r, const float* obs0_dev, void* stream) {
  if (!r || !obs0_dev) return rfail(ARMSIM_E_INVALID, "armsim_replay_begin: null argument");
  ReplayDeviceGuard guard(r->device);
  const int grid = (r->d.n * r->d.O + 255) / 256;
  replay_begin_kernel<<<grid < 1184 ? grid : 1184, 256, 0, (cudaStream_t)stream>>>(r->d, obs0_dev);
  RCU(cudaGetLastError());
  return ARMSIM_OK;
}

int armsim_replay_store(ArmReplay* r, const float* action_dev, const float* reward_dev, const uint8_t* done_dev,
                        const float* final_obs_dev, const float* obs_out_dev, void* stream) {
  if (!r || !action_dev || !reward_dev || !done_dev || !final_obs_dev || !obs_out_dev)
    return rfail(ARMSIM_E_INVALID, "armsim_replay_store: null argument");
  ReplayDeviceGuard guard(r->device);
  int grid = (r->d.n * r->d.O + RB - 1) / RB;
  if (grid > 1184) grid = 1184;       // 148 SMs x 8 resident CTAs, grid-stride beyond
  replay_store_kernel<<<grid, RB, 0, (cudaStream_t)stream>>>(r->d, action_dev, reward_dev, done_dev, final_obs_dev, obs_out_dev);
  RCU(cudaGetLastError());
  return ARMSIM_OK;
}

int armsim_replay_sample(ArmReplay* r, int32_t batch, int32_t use_her, float dis_threshold, float her_ratio, float* states_dev,
                         float* actions_dev, float* next_states_dev, float* rewards_dev, float* dones_dev, int32_t* picks_dev,
                         void* stream) {
  if (!r || batch <= 0 || !states_dev || !actions_dev || !next_states_dev || !rewards_dev || !dones_dev)
    return rfail(ARMSIM_E_INVALID, "armsim_replay_sample: bad argument");
  ReplayDeviceGuard guard(r->device);
  replay_sample_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r->d, batch, use_her, dis_threshold, her_ratio, states_dev,
                                                                          actions_dev, next_states_dev, rewards_dev, dones_dev, picks_dev);
  RCU(cudaGetLastError());
  return ARMSIM_OK;
}

int armsim_replay_gather(ArmReplay* r, int32_t batch, const int32_t* slot_dev, const int32_t* step_dev, const int32_t* goal_step_dev,
                         float dis_threshold, float* states_dev, float* actions_dev, float* next_states_dev, float* rewards_dev,
                         float* dones_dev, void* stream) {
  if (!r || batch <= 0 || !slot_dev || !step_dev || !goal_step_dev) return rfail(ARMSIM_E_INVALID, "armsim_replay_gather: bad argument");
  ReplayDeviceGuard guard(r->device);
  replay_gather_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r->d, batch, slot_dev, step_dev, goal_step_dev, dis_threshold,
                                                                          states_dev, actions_dev, next_states_dev, rewards_dev, dones_dev);
  RCU(cudaGetLastError());
  return ARMSIM_OK;
}

/* host read-back of the counters / trajectory table (synchronises): info[0] = rows stored, [1] = trajectories appended,
 * [2] = sample calls, [3] = sample calls that found no intact trajectory (zero-filled batches) */
int armsim_replay_info(ArmReplay* r, int64_t info[4]) {
  if (!r || !info) return rfail(ARMSIM_E_INVALID, "armsim_replay_info: null argument");
  RCU(cudaSetDevice(r->device));
  RCU(cudaDeviceSynchronize());
  unsigned long long c[4];
  RCU(cudaMemcpy(c, r->d.counters, sizeof(c), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) info[i] = (int64_t)c[i];
  return ARMSIM_OK;
}

/* checkpointing: the whole ring + table + cursors as one opaque blob (valid for an identical ArmReplayConfig) */
int64_t armsim_replay_state_bytes(ArmReplay* r) { return r ? (int64_t)r->bytes : (int64_t)ARMSIM_E_INVALID; }

int armsim_replay_get_state(ArmReplay* r, void* host_dst, int64_t bytes) {
  if (!r || !host_dst || bytes != (int64_t)r->bytes) return rfail(ARMSIM_E_STATE, "armsim_replay_get_state: need exactly %zu bytes", r ? r->bytes : (size_t)0);
  RCU(cudaSetDevice(r->device));
  RCU(cudaDeviceSynchronize());
  RCU(cudaMemcpy(host_dst, r->block, r->bytes, cudaMemcpyDeviceToHost));
  return ARMSIM_OK;
}

int armsim_replay_set_state(ArmReplay* r, const void* host_src, int64_t bytes) {
  if (!r || !host_src || bytes != (int64_t)r->bytes) return rfail(ARMSIM_E_STATE, "armsim_replay_set_state: need exactly %zu bytes", r ? r->bytes : (size_t)0);
  RCU(cudaSetDevice(r->device));
  RCU(cudaDeviceSynchronize());
  RCU(cudaMemcpy(r->block, host_src, r->bytes, cudaMemcpyHostToDevice));
  return ARMSIM_OK;
}

int armsim_replay_table(ArmReplay* r, int32_t* env_host, int64_t* start_host, int32_t* len_host, int32_t count) {
  if (!r || count < 0 || count > r->d.cap) return rfail(ARMSIM_E_INVALID, "armsim_replay_table: bad count");
  RCU(cudaSetDevice(r->device));
  RCU(cudaDeviceSynchronize());
  if (env_host) RCU(cudaMemcpy(env_host, r->d.t_env, (size_t)count * 4, cudaMemcpyDeviceToHost));
  if (start_host) RCU(cudaMemcpy(start_host, r->d.t_start, (size_t)count * 8, cudaMemcpyDeviceToHost));
  if (len_host) RCU(cudaMemcpy(len_host, r->d.t_len, (size_t)count * 4, cudaMemcpyDeviceToHost));
  return ARMSIM_OK;
}

}  // extern "C"
