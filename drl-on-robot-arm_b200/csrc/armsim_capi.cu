// armsim_capi.cu -- the C-ABI of include/armsim.h on top of the sm_100a kernels.
// No torch types, no exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "armsim.h"
#include "armsim_defaults.h"
#include "armsim_kernels.cuh"
#include "policy_kernels.cuh"
#include "armsim_robot_models.h"

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return fail(ARMSIM_E_CUDA, "%s: %s", #call, cudaGetErrorString(_e));    \
  } while (0)

// The device-pointer entry points launch on the caller's stream; the handle's device must be current for the launch
// (single-process multi-GPU callers).  Switches only when needed and restores the caller's device.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

struct ArmSim {
  ArmsimConfig cfg;
  int n = 0, obs_dim = 0, act_dim = 3, mapping = ARMSIM_MAP_LANE;
  ChainParams chain;
  TaskParams task;
  DynParams dyn;
  StatePtrs S{};
  void* state_block = nullptr;   // one allocation holding every SoA field
  // *_host path: one pinned block + one device block, outputs contiguous so the D2H is a single copy
  char* h_pin = nullptr;         // mapped + portable pinned block: [actions | obs | reward | done | success | flag]
  char* d_io = nullptr;          // device twin of the same layout (large batches: DMA copies instead of zero-copy)
  size_t off_obs = 0, off_reward = 0, off_done = 0, off_success = 0, out_bytes = 0, act_bytes = 0;
  unsigned long long* d_stats = nullptr;   // {episodes, successes, return sum (2^-16 fixed point)} of armsim_track_episodes
  unsigned int* d_cta_seq = nullptr;   // per-block launch counters of the doorbells (device)
  volatile unsigned int* h_flags = nullptr;   // per-block doorbells (mapped host memory, tail of h_pin)
  unsigned int seq = 0;                // host-path launches so far == the value every doorbell shows when one is done
  int grid = 0;
  bool zero_copy = true;
  cudaGraphExec_t host_graph = nullptr;   // the zero-copy host step as an instantiated graph (parameters never change)
  bool host_graph_tried = false;
  bool host_pending = false;              // armsim_step_host_async issued, armsim_step_host_wait not yet
  // resident step server (armsim_host_server): the kernel stays on the GPU between host steps
  bool server_enabled = false, server_live = false;
  unsigned long long server_idle_ns = 0;
  volatile unsigned int* h_cmd = nullptr;    // mapped host word: sequence number of the step requested, or SERVER_STOP
  unsigned int* d_relay = nullptr;           // device word through which block 0 republishes h_cmd
  unsigned int* h_relay_init = nullptr;      // pinned staging word for resetting d_relay
  double server_last_use = 0.0;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  float* fk_scratch = nullptr;   // armsim_fk_host staging (q | pos | rot), grown on demand
  int fk_cap = 0;
};

static void rpy_to_mat(const double rpy[3], double R[9]) {
  const double cr = cos(rpy[0]), sr = sin(rpy[0]), cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

// FK(q) of the chain in fp64 on the host (used once per handle for the constant start-of-episode EE pose)
template <class Model>
static void host_fk(const Model& m, const double q[NJ], double p[3], double R[9]) {
  auto mul = [](const double A[9], const double B[9], double Cm[9]) {
    double Tm[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Tm[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(Cm, Tm, sizeof(Tm));
  };
  rpy_to_mat(m.base_rpy, R);
  for (int i = 0; i < 3; ++i) p[i] = m.base_xyz[i];
  for (int j = 0; j < NJ; ++j) {
    for (int i = 0; i < 3; ++i) p[i] += R[3 * i] * m.xyz[j][0] + R[3 * i + 1] * m.xyz[j][1] + R[3 * i + 2] * m.xyz[j][2];
    double Rf[9];
    rpy_to_mat(m.rpy[j], Rf);
    mul(R, Rf, R);
    const double cq = cos(q[j]), sq = sin(q[j]);
    const double Rz[9] = {cq, -sq, 0, sq, cq, 0, 0, 0, 1};
    mul(R, Rz, R);
  }
}

template <class Model>
static void fill_chain(const Model& m, ChainParams& c) {
  double R[9];
  rpy_to_mat(m.base_rpy, R);
  for (int i = 0; i < 9; ++i) c.Rb[i] = (float)R[i];
  for (int i = 0; i < 3; ++i) c.tb[i] = (float)m.base_xyz[i];
  for (int j = 0; j < NJ; ++j) {
    rpy_to_mat(m.rpy[j], R);
    for (int i = 0; i < 9; ++i) c.Rf[j][i] = (float)R[i];
    for (int i = 0; i < 3; ++i) c.t[j][i] = (float)m.xyz[j][i];
    c.lower[j] = (float)m.lower[j];
    c.upper[j] = (float)m.upper[j];
  }
}

// torque mode: spatial inertia of each link about its frame origin, limits, gravity as a base acceleration
template <class Model>
static void fill_dyn(const Model& m, const ArmsimConfig& cfg, DynParams& d) {
  memset(&d, 0, sizeof(d));
  for (int j = 0; j < NJ; ++j) {
    const double mass = m.mass[j], *c = m.com[j], *I = m.inertia[j];
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double full[6] = {I[0] + mass * (cc - c[0] * c[0]), I[1] - mass * c[0] * c[1], I[2] - mass * c[0] * c[2],
                            I[3] + mass * (cc - c[1] * c[1]), I[4] - mass * c[1] * c[2], I[5] + mass * (cc - c[2] * c[2])};
    for (int k = 0; k < 6; ++k) d.Ibar[j][k] = (float)full[k];
    for (int k = 0; k < 3; ++k) d.h[j][k] = (float)(mass * c[k]);
    d.mass[j] = (float)mass;
    d.effort[j] = (float)m.effort[j];
    d.maxvel[j] = (float)m.velocity[j];
    d.damping[j] = (float)m.damping[j];
  }
  double Rb[9];
  rpy_to_mat(m.base_rpy, Rb);
  for (int i = 0; i < 3; ++i)   // Rb^T (-g)
    d.abase[i] = (float)(-(Rb[i] * cfg.gravity[0] + Rb[3 + i] * cfg.gravity[1] + Rb[6 + i] * cfg.gravity[2]));
  d.dt = (float)cfg.sim_dt;
}

static void quat_from_euler(const double rpy[3], float q[4]) {
  const double hr = rpy[0] * 0.5, hp = rpy[1] * 0.5, hy = rpy[2] * 0.5;
  const double cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy = sin(hy);
  q[0] = (float)(sr * cp * cy - cr * sp * sy);
  q[1] = (float)(cr * sp * cy + sr * cp * sy);
  q[2] = (float)(cr * cp * sy - sr * sp * cy);
  q[3] = (float)(cr * cp * cy + sr * sp * sy);
}

static int field_width(int32_t f) {
  switch (f) {
    case ARMSIM_F_Q: case ARMSIM_F_QD: return 7;
    case ARMSIM_F_GOAL: case ARMSIM_F_CUBE_POS: case ARMSIM_F_CUBE_LINVEL: case ARMSIM_F_CUBE_ANGVEL: return 3;
    case ARMSIM_F_CUBE_QUAT: return 4;
    case ARMSIM_F_STEP: case ARMSIM_F_EPISODE: case ARMSIM_F_LAST_DIST: case ARMSIM_F_GRIP: case ARMSIM_F_IK_ITERS:
    case ARMSIM_F_EP_RETURN: case ARMSIM_F_EXPLORE_COUNT: return 1;
    default: return -1;
  }
}

static void* field_ptr(ArmSim* s, int32_t f) {
  const size_t n = (size_t)s->n;
  switch (f) {
    case ARMSIM_F_Q: return s->S.q;
    case ARMSIM_F_QD: return s->S.qd;
    case ARMSIM_F_GOAL: return s->S.goal;
    case ARMSIM_F_STEP: return s->S.step;
    case ARMSIM_F_EPISODE: return s->S.episode;
    case ARMSIM_F_CUBE_POS: return s->S.cube;
    case ARMSIM_F_CUBE_QUAT: return s->S.cube + 3 * n;
    case ARMSIM_F_CUBE_LINVEL: return s->S.cube + 7 * n;
    case ARMSIM_F_CUBE_ANGVEL: return s->S.cube + 10 * n;
    case ARMSIM_F_LAST_DIST: return s->S.last_dist;
    case ARMSIM_F_GRIP: return s->S.grip;
    case ARMSIM_F_IK_ITERS: return s->S.ik_iters;
    case ARMSIM_F_EP_RETURN: return s->S.ep_return;
    case ARMSIM_F_EXPLORE_COUNT: return s->S.explore_count;
    default: return nullptr;
  }
}

// Step kernels go out with the programmatic-stream-serialization attribute (programmatic dependent launch): a step
// launched right behind another kernel of the stream is scheduled onto idle SMs while that kernel still runs and
// parks in griddepcontrol.wait (first instruction of the kernel, before any global read) until the predecessor has
// completed and flushed -- the ~1 us launch gap between back-to-back steps overlaps the previous step's tail.
// ARMSIM_PDL=0 in the environment turns it off.
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ARMSIM_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Dynamic shared memory of a kernel that may step a cube (per-thread contact slots, cube_model.cuh): 64 KB per block,
// above the 48 KB default, so every such kernel function opts in once.
template <class F>
static void ensure_smem(F kern, size_t bytes) {
  // static + dynamic shared memory above 48 KB needs the opt-in even when the dynamic part alone is below it (torque
  // kernels: 12 KB static staging + 41.5 KB contact slots), so every kernel with contact slots opts in -- once per
  // device (the attribute belongs to the function on the CURRENT device) and safely from several host threads
  if (bytes == 0) return;
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& p : done)
    if (p.first == (const void*)kern && p.second == dev) return;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  done.emplace_back((const void*)kern, dev);
}

template <class... KArgs, class... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), int grid, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  ensure_smem(kern, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(LANE_BLOCK);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int TASK>
constexpr size_t task_smem() { return TaskTraits<TASK>::HAS_CUBE ? (size_t)cube::scratch_bytes<TASK == ARMSIM_TASK_PICK>() : 0; }

extern "C" {

int32_t armsim_abi_version(void) { return ARMSIM_ABI_VERSION; }
const char* armsim_last_error(void) { return g_err; }

int armsim_default_config(int32_t task, ArmsimConfig* cfg) {
  int rc = armsim_fill_default_config(task, cfg);
  if (rc) return fail(rc, "armsim_default_config: bad task %d or null cfg", task);
  return ARMSIM_OK;
}

int32_t armsim_obs_dim(const ArmSim* s) { return s ? s->obs_dim : ARMSIM_E_INVALID; }
int32_t armsim_action_dim(const ArmSim* s) { return s ? s->act_dim : ARMSIM_E_INVALID; }
int32_t armsim_n_envs(const ArmSim* s) { return s ? s->n : ARMSIM_E_INVALID; }
int32_t armsim_mapping(const ArmSim* s) { return s ? s->mapping : ARMSIM_E_INVALID; }
int64_t armsim_launch_count(const ArmSim* s) { return s ? s->launches : ARMSIM_E_INVALID; }

void armsim_destroy(ArmSim* s) {
  if (!s) return;
  cudaSetDevice(s->cfg.device);
  if (s->server_live && s->h_cmd) __atomic_store_n((unsigned int*)s->h_cmd, 0xffffffffu, __ATOMIC_RELEASE);   // SERVER_STOP
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->state_block) cudaFree(s->state_block);
  if (s->d_io) cudaFree(s->d_io);
  if (s->fk_scratch) cudaFree(s->fk_scratch);
  if (s->host_graph) cudaGraphExecDestroy(s->host_graph);
  if (s->d_cta_seq) cudaFree(s->d_cta_seq);
  if (s->d_stats) cudaFree(s->d_stats);
  if (s->h_pin) cudaFreeHost(s->h_pin);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

static int launch_reset(ArmSim* s, const uint8_t* mask_dev, float* obs_dev, cudaStream_t st) {
  const int grid = (s->n + LANE_BLOCK - 1) / LANE_BLOCK;
  switch (s->cfg.task) {
    case ARMSIM_TASK_REACH: ensure_smem(reset_lane_kernel<ARMSIM_TASK_REACH>, task_smem<ARMSIM_TASK_REACH>()); reset_lane_kernel<ARMSIM_TASK_REACH><<<grid, LANE_BLOCK, task_smem<ARMSIM_TASK_REACH>(), st>>>(s->chain, s->task, s->S, mask_dev, obs_dev); break;
    case ARMSIM_TASK_PUSH: ensure_smem(reset_lane_kernel<ARMSIM_TASK_PUSH>, task_smem<ARMSIM_TASK_PUSH>()); reset_lane_kernel<ARMSIM_TASK_PUSH><<<grid, LANE_BLOCK, task_smem<ARMSIM_TASK_PUSH>(), st>>>(s->chain, s->task, s->S, mask_dev, obs_dev); break;
    case ARMSIM_TASK_PICK: ensure_smem(reset_lane_kernel<ARMSIM_TASK_PICK>, task_smem<ARMSIM_TASK_PICK>()); reset_lane_kernel<ARMSIM_TASK_PICK><<<grid, LANE_BLOCK, task_smem<ARMSIM_TASK_PICK>(), st>>>(s->chain, s->task, s->S, mask_dev, obs_dev); break;
    case ARMSIM_TASK_KUKA_REACH: ensure_smem(reset_lane_kernel<ARMSIM_TASK_KUKA_REACH>, task_smem<ARMSIM_TASK_KUKA_REACH>()); reset_lane_kernel<ARMSIM_TASK_KUKA_REACH><<<grid, LANE_BLOCK, task_smem<ARMSIM_TASK_KUKA_REACH>(), st>>>(s->chain, s->task, s->S, mask_dev, obs_dev); break;
    default: return fail(ARMSIM_E_INVALID, "bad task");
  }
  s->launches += 1;
  CU(cudaGetLastError());
  return ARMSIM_OK;
}

// kernels are instantiated per (task, robot): the built-in robots use the generated straight-line FK, custom chains
// the parameter-driven one
#define ARMSIM_STEP_CASE(TASK, ROBOT)                                                                           \
  case (TASK) * 4 + (ROBOT):                                                                                    \
    if (grid > (TaskTraits<TASK>::HAS_CUBE ? DENSE_GRID_THRESHOLD_CUBE : DENSE_GRID_THRESHOLD))                       \
      lerr = launch_k(step_lane_kernel<TASK, ROBOT, BUILD_DENSE>, grid, task_smem<TASK>(), st, H.flags == nullptr, s->chain, s->task, s->S, a, o, r, d, su, fo, H); \
    else if (!TaskTraits<TASK>::HAS_CUBE && grid > PREFETCH_MAX_GRID)                                             \
      lerr = launch_k(step_lane_kernel<TASK, ROBOT, TaskTraits<TASK>::HAS_CUBE ? BUILD_LATENCY : BUILD_WAVE>, grid, task_smem<TASK>(), st, H.flags == nullptr, s->chain, s->task, s->S, a, o, r, d, su, fo, H); \
    else                                                                                                          \
      lerr = launch_k(step_lane_kernel<TASK, ROBOT, BUILD_LATENCY>, grid, task_smem<TASK>(), st, H.flags == nullptr, s->chain, s->task, s->S, a, o, r, d, su, fo, H); \
    break;

#define ARMSIM_TORQUE_CASE(TASK, ROBOT)                                                                                \
  case (TASK) * 4 + (ROBOT):                                                                                           \
    lerr = launch_k(step_torque_kernel<TASK, ROBOT>, grid, task_smem<TASK>(), st, H.flags == nullptr, s->chain, s->task, s->dyn, s->S, a, o, r, d, su, fo, H); \
    break;

static int launch_step(ArmSim* s, const float* a, float* o, float* r, uint8_t* d, uint8_t* su, cudaStream_t st,
                       const HostNotify H = HostNotify{nullptr, nullptr}, float* fo = nullptr) {
  const int grid = s->grid;
  cudaError_t lerr = cudaSuccess;
  if (s->cfg.mode == ARMSIM_MODE_TORQUE) {
    switch (s->cfg.task * 4 + s->cfg.robot) {
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_KUKA_IIWA)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_DIANA_S1)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_CUSTOM)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_KUKA_IIWA)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_DIANA_S1)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_CUSTOM)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_KUKA_IIWA)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_DIANA_S1)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_CUSTOM)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_KUKA_IIWA)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_DIANA_S1)
      ARMSIM_TORQUE_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_CUSTOM)
      default: return fail(ARMSIM_E_INVALID, "bad task / robot");
    }
    if (lerr != cudaSuccess) return fail(ARMSIM_E_CUDA, "step launch: %s", cudaGetErrorString(lerr));
    s->launches += 1;
    CU(cudaGetLastError());
    return ARMSIM_OK;
  }
  switch (s->cfg.task * 4 + s->cfg.robot) {
    ARMSIM_STEP_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_STEP_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_STEP_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_CUSTOM)
    ARMSIM_STEP_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_STEP_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_STEP_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_CUSTOM)
    ARMSIM_STEP_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_STEP_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_STEP_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_CUSTOM)
    ARMSIM_STEP_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_STEP_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_STEP_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_CUSTOM)
    default: return fail(ARMSIM_E_INVALID, "bad task / robot");
  }
  if (lerr != cudaSuccess) return fail(ARMSIM_E_CUDA, "step launch: %s", cudaGetErrorString(lerr));
  s->launches += 1;
  CU(cudaGetLastError());
  return ARMSIM_OK;
}

int armsim_create(const ArmsimConfig* cfg, ArmSim** out) {
  if (!cfg || !out) return fail(ARMSIM_E_INVALID, "armsim_create: null argument");
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(ArmsimConfig))
    return fail(ARMSIM_E_INVALID, "armsim_create: struct_size %d != %zu (ABI mismatch)", cfg->struct_size, sizeof(ArmsimConfig));
  if (cfg->n_envs <= 0) return fail(ARMSIM_E_INVALID, "armsim_create: n_envs must be > 0");
  if (cfg->task < 0 || cfg->task > ARMSIM_TASK_KUKA_REACH) return fail(ARMSIM_E_INVALID, "armsim_create: bad task %d", cfg->task);
  if (cfg->mode != ARMSIM_MODE_IK_TELEPORT && cfg->mode != ARMSIM_MODE_TORQUE) return fail(ARMSIM_E_INVALID, "armsim_create: bad mode %d", cfg->mode);
  if (cfg->mode == ARMSIM_MODE_TORQUE && !(cfg->sim_dt > 0.0)) return fail(ARMSIM_E_INVALID, "armsim_create: torque mode needs sim_dt > 0");
  if (cfg->mapping != ARMSIM_MAP_AUTO && cfg->mapping != ARMSIM_MAP_LANE) return fail(ARMSIM_E_INVALID, "armsim_create: bad mapping %d (one lane per arm is the only mapping)", cfg->mapping);
  if (cfg->ik_max_iters < 0 || cfg->max_steps < 0) return fail(ARMSIM_E_INVALID, "armsim_create: negative iteration / step limit");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(ARMSIM_E_CUDA, "armsim_create: no usable CUDA device (%s); libarmsim has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(ARMSIM_E_INVALID, "armsim_create: device %d out of range [0,%d)", cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));

  ArmSim* s = new (std::nothrow) ArmSim();
  if (!s) return fail(ARMSIM_E_NOMEM, "armsim_create: host allocation failed");
  s->cfg = *cfg;
  s->cfg.custom_chain = nullptr;
  s->n = cfg->n_envs;
  double ee0[3], R0[9];
  if (cfg->robot == ARMSIM_ROBOT_KUKA_IIWA) { fill_chain(ARMSIM_MODEL_KUKA_IIWA, s->chain); fill_dyn(ARMSIM_MODEL_KUKA_IIWA, *cfg, s->dyn); host_fk(ARMSIM_MODEL_KUKA_IIWA, cfg->init_q, ee0, R0); }
  else if (cfg->robot == ARMSIM_ROBOT_DIANA_S1) { fill_chain(ARMSIM_MODEL_DIANA_S1, s->chain); fill_dyn(ARMSIM_MODEL_DIANA_S1, *cfg, s->dyn); host_fk(ARMSIM_MODEL_DIANA_S1, cfg->init_q, ee0, R0); }
  else if (cfg->robot == ARMSIM_ROBOT_CUSTOM && cfg->custom_chain) { fill_chain(*cfg->custom_chain, s->chain); fill_dyn(*cfg->custom_chain, *cfg, s->dyn); host_fk(*cfg->custom_chain, cfg->init_q, ee0, R0); }
  else { delete s; return fail(ARMSIM_E_INVALID, "armsim_create: bad robot %d (custom needs custom_chain)", cfg->robot); }

  TaskParams& T = s->task;
  memset(&T, 0, sizeof(T));
  T.task = cfg->task; T.n = s->n; T.max_steps = cfg->max_steps; T.auto_reset = cfg->auto_reset ? 1 : 0;
  T.ik_max_iters = cfg->ik_max_iters; T.napply = cfg->task == ARMSIM_TASK_PICK ? 6 : NJ;
  T.clamp = cfg->clamp_joint_limits ? 1 : 0;
  T.torque_mode = cfg->mode == ARMSIM_MODE_TORQUE ? 1 : 0;
  s->act_dim = T.torque_mode ? ARMSIM_TORQUE_DIM : ARMSIM_ACT_DIM;
  T.obs_dim = s->obs_dim = (cfg->task == ARMSIM_TASK_REACH ? 6 : (cfg->task == ARMSIM_TASK_KUKA_REACH ? 3 : 9)) + (T.torque_mode ? 2 * NJ : 0);
  T.dv = (float)cfg->dv; T.reach_dis = (float)cfg->reach_dis; T.ik_damping = (float)cfg->ik_damping; T.ik_residual = (float)cfg->ik_residual;
  for (int i = 0; i < 3; ++i) {
    T.ws_lo[i] = (float)cfg->ws_lo[i]; T.ws_hi[i] = (float)cfg->ws_hi[i];
    T.goal_lo[i] = (float)cfg->goal_lo[i]; T.goal_span[i] = (float)(cfg->goal_hi[i] - cfg->goal_lo[i]);
  }
  quat_from_euler(cfg->target_rpy, T.tquat);
  {
    double tRd[9];
    rpy_to_mat(cfg->target_rpy, tRd);   // getQuaternionFromEuler = Rz(yaw) Ry(pitch) Rx(roll), the URDF rpy convention
    for (int i = 0; i < 9; ++i) T.tR[i] = (float)tRd[i];
  }
  for (int j = 0; j < NJ; ++j) T.init_q[j] = (float)cfg->init_q[j];
  for (int i = 0; i < 3; ++i) T.init_ee[i] = (float)ee0[i];
  for (int i = 0; i < 9; ++i) T.init_R[i] = (float)R0[i];
  T.seed_lo = (uint32_t)cfg->seed; T.seed_hi = (uint32_t)(cfg->seed >> 32);
  T.gid_offset = cfg->env_id_offset;
  s->mapping = ARMSIM_MAP_LANE;

  // state: one block, every field padded to 256 B so each SoA row starts on its own lines
  const size_t n = (size_t)s->n;
  auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t sz_q = pad(7 * n * 4), sz_goal = pad(3 * n * 4), sz_i = pad(n * 4), sz_b = pad(n), sz_cube = pad(13 * n * 4);
  const size_t total = 2 * sz_q + sz_goal + 3 * sz_i + sz_b + sz_cube + 4 * sz_i;
  if (cudaMalloc(&s->state_block, total) != cudaSuccess) {
    cudaGetLastError();
    delete s;
    return fail(ARMSIM_E_NOMEM, "armsim_create: cudaMalloc of %zu state bytes failed", total);
  }
  cudaMemset(s->state_block, 0, total);
  char* b = (char*)s->state_block;
  s->S.q = (float*)b; b += sz_q;
  s->S.qd = (float*)b; b += sz_q;
  s->S.goal = (float*)b; b += sz_goal;
  s->S.step = (int*)b; b += sz_i;
  s->S.episode = (int*)b; b += sz_i;
  s->S.ik_iters = (int*)b; b += sz_i;
  s->S.done = (uint8_t*)b; b += sz_b;
  s->S.cube = (float*)b; b += sz_cube;
  s->S.last_dist = (float*)b; b += sz_i;
  s->S.grip = (float*)b; b += sz_i;
  s->S.ep_return = (float*)b; b += sz_i;
  s->S.explore_count = (unsigned int*)b; b += sz_i;

  // host-path staging
  s->act_bytes = pad(n * s->act_dim * 4);
  s->off_obs = 0;
  s->off_reward = pad(n * s->obs_dim * 4);
  s->off_done = s->off_reward + pad(n * 4);
  s->off_success = s->off_done + pad(n);
  s->out_bytes = s->off_success + pad(n);
  // zero-copy (kernel reads actions from / writes results to mapped host memory, completion by doorbell) while the
  // per-step payload is small enough for PCIe latency, not bandwidth, to dominate; DMA copies beyond that
  s->zero_copy = n <= 65536;
  s->grid = (s->n + LANE_BLOCK - 1) / LANE_BLOCK;
  const size_t flag_bytes = pad((size_t)s->grid * sizeof(unsigned int));
  if (cudaMalloc((void**)&s->d_io, s->act_bytes + s->out_bytes) != cudaSuccess ||
      cudaMalloc((void**)&s->d_stats, 3 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMemset(s->d_stats, 0, 3 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMalloc((void**)&s->d_cta_seq, flag_bytes + 256) != cudaSuccess ||
      cudaMemset(s->d_cta_seq, 0, flag_bytes + 256) != cudaSuccess ||
      cudaHostAlloc((void**)&s->h_pin, s->act_bytes + s->out_bytes + flag_bytes + 512, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    armsim_destroy(s);
    return fail(ARMSIM_E_NOMEM, "armsim_create: staging allocation failed");
  }
  s->h_flags = (volatile unsigned int*)(s->h_pin + s->act_bytes + s->out_bytes);
  memset((void*)s->h_flags, 0, flag_bytes + 512);
  s->h_cmd = (volatile unsigned int*)(s->h_pin + s->act_bytes + s->out_bytes + flag_bytes);
  s->h_relay_init = (unsigned int*)(s->h_pin + s->act_bytes + s->out_bytes + flag_bytes + 256);
  s->d_relay = (unsigned int*)((char*)s->d_cta_seq + flag_bytes);
  int rc = launch_reset(s, nullptr, nullptr, s->stream);
  if (rc == ARMSIM_OK && cudaStreamSynchronize(s->stream) != cudaSuccess) rc = fail(ARMSIM_E_CUDA, "armsim_create: initial reset failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc != ARMSIM_OK) { armsim_destroy(s); return rc; }
  *out = s;
  return ARMSIM_OK;
}

static int server_launch(ArmSim* s);
static int server_quiesce(ArmSim* s);
#define ARMSIM_QUIESCE(s)                   \
  do {                                      \
    int _q = server_quiesce(s);             \
    if (_q) return _q;                      \
  } while (0)

int armsim_reset(ArmSim* s, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_reset: null handle");
  ARMSIM_QUIESCE(s);
  DeviceGuard guard(s->cfg.device);
  return launch_reset(s, mask_dev, obs_dev, (cudaStream_t)stream);
}

int armsim_step(ArmSim* s, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint8_t* success_dev,
                void* stream) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_step: null handle");
  ARMSIM_QUIESCE(s);
  if (!action_dev || !obs_dev || !reward_dev || !done_dev || !success_dev) return fail(ARMSIM_E_INVALID, "armsim_step: null buffer");
  DeviceGuard guard(s->cfg.device);
  return launch_step(s, action_dev, obs_dev, reward_dev, done_dev, success_dev, (cudaStream_t)stream);
}

// Wait until every block's doorbell shows `seq`.  Polls the mapped words; every so often asks the driver whether the
// stream died so a faulting kernel turns into an error code instead of a hang.
static int wait_doorbells(ArmSim* s, unsigned int seq) {
  const volatile unsigned int* f = s->h_flags;
  const int grid = s->grid;
  int b = 0;                                   // doorbells [0, b) already seen at seq
  for (unsigned long long spins = 1;; ++spins) {
    while (b < grid && f[b] == seq) ++b;
    if (b == grid) {
      // the outputs were written before the doorbells (device side: fence + store); order OUR loads of them after the
      // doorbell loads too -- free on x86 (TSO), required on weakly ordered hosts (aarch64 / Grace)
      __atomic_thread_fence(__ATOMIC_ACQUIRE);
      return ARMSIM_OK;
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
    if ((spins & 0xFFFFull) == 0) {
      cudaError_t q = cudaStreamQuery(s->stream);
      if (q == cudaSuccess) {
        while (b < grid && f[b] == seq) ++b;
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        if (b == grid) return ARMSIM_OK;
        if (s->server_enabled) {        // the resident kernel timed out with this command (partly) unserved: start it again
          s->server_live = false;
          int rc = server_launch(s);
          if (rc) return rc;
          continue;
        }
        return fail(ARMSIM_E_CUDA, "armsim_step_host: kernel finished without ringing every doorbell");
      }
      if (q != cudaErrorNotReady) return fail(ARMSIM_E_CUDA, "armsim_step_host: %s", cudaGetErrorString(q));
    }
  }
}

int armsim_host_buffers(ArmSim* s, float** action, float** obs, float** reward, uint8_t** done, uint8_t** success) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_host_buffers: null handle");
  char* h_out = s->h_pin + s->act_bytes;
  if (action) *action = (float*)s->h_pin;
  if (obs) *obs = (float*)(h_out + s->off_obs);
  if (reward) *reward = (float*)(h_out + s->off_reward);
  if (done) *done = (uint8_t*)(h_out + s->off_done);
  if (success) *success = (uint8_t*)(h_out + s->off_success);
  return ARMSIM_OK;
}

int armsim_step_ex(ArmSim* s, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint8_t* success_dev,
                   float* final_obs_dev, void* stream) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_step_ex: null handle");
  ARMSIM_QUIESCE(s);
  if (!action_dev || !obs_dev || !reward_dev || !done_dev || !success_dev) return fail(ARMSIM_E_INVALID, "armsim_step_ex: null buffer");
  DeviceGuard guard(s->cfg.device);
  return launch_step(s, action_dev, obs_dev, reward_dev, done_dev, success_dev, (cudaStream_t)stream, HostNotify{nullptr, nullptr},
                     final_obs_dev);
}

int armsim_step_tracked(ArmSim* s, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                        uint8_t* success_dev, float* final_obs_dev, void* stream) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_step_tracked: null handle");
  ARMSIM_QUIESCE(s);
  if (!action_dev || !obs_dev || !reward_dev || !done_dev || !success_dev) return fail(ARMSIM_E_INVALID, "armsim_step_tracked: null buffer");
  if (s->cfg.mode == ARMSIM_MODE_TORQUE) return fail(ARMSIM_E_INVALID, "armsim_step_tracked: IK-teleport mode only (use armsim_step_ex + armsim_track_episodes)");
  DeviceGuard guard(s->cfg.device);
  HostNotify H{nullptr, nullptr};
  H.track_stats = s->d_stats;
  return launch_step(s, action_dev, obs_dev, reward_dev, done_dev, success_dev, (cudaStream_t)stream, H, final_obs_dev);
}

// ---------------------------------------------------------------------------------------------- resident step server
static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define ARMSIM_SERVER_CASE(TASK, ROBOT)                                                                                  \
  case (TASK) * 4 + (ROBOT):                                                                                             \
    ensure_smem(step_server_kernel<TASK, ROBOT>, task_smem<TASK>());                                                     \
    step_server_kernel<TASK, ROBOT><<<s->grid, LANE_BLOCK, task_smem<TASK>(), s->stream>>>(s->chain, s->task, s->S, (const float*)s->h_pin, o, r, d, su, H, ctl); \
    break;

// Put the resident kernel on the handle's stream.  Blocks resume from their completed-step counters (d_cta_seq), so a
// command that was pending when the previous instance timed out runs exactly once.
static int server_launch(ArmSim* s) {
  char* h_out = s->h_pin + s->act_bytes;
  float* o = (float*)(h_out + s->off_obs);
  float* r = (float*)(h_out + s->off_reward);
  uint8_t *d = (uint8_t*)(h_out + s->off_done), *su = (uint8_t*)(h_out + s->off_success);
  HostNotify H{s->d_cta_seq, (unsigned int*)s->h_flags};
  ServerCtl ctl{s->h_cmd, s->d_relay, s->server_idle_ns};
  if (*s->h_cmd == SERVER_STOP) *s->h_cmd = s->seq;
  *s->h_relay_init = 0u;                                   // "nothing requested yet" for every block (sequence numbers start at 1)
  CU(cudaMemcpyAsync(s->d_relay, s->h_relay_init, sizeof(unsigned int), cudaMemcpyHostToDevice, s->stream));
  switch (s->cfg.task * 4 + s->cfg.robot) {
    ARMSIM_SERVER_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_REACH, ARMSIM_ROBOT_CUSTOM)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_PUSH, ARMSIM_ROBOT_CUSTOM)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_PICK, ARMSIM_ROBOT_CUSTOM)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_KUKA_IIWA)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_DIANA_S1)
    ARMSIM_SERVER_CASE(ARMSIM_TASK_KUKA_REACH, ARMSIM_ROBOT_CUSTOM)
    default: return fail(ARMSIM_E_INVALID, "bad task / robot");
  }
  CU(cudaGetLastError());
  s->launches += 1;
  s->server_live = true;
  s->server_last_use = now_s();
  return ARMSIM_OK;
}

// Make the resident kernel leave (every entry point that touches the handle's device state outside the host step calls
// this first).  Cheap when no server is running.
static int server_quiesce(ArmSim* s) {
  if (!s->server_live) return ARMSIM_OK;
  if (s->host_pending) return fail(ARMSIM_E_STATE, "a host step is in flight: call armsim_step_host_wait first");
  __atomic_store_n((unsigned int*)s->h_cmd, SERVER_STOP, __ATOMIC_RELEASE);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  s->server_live = false;
  __atomic_store_n((unsigned int*)s->h_cmd, s->seq, __ATOMIC_RELEASE);
  if (e != cudaSuccess) return fail(ARMSIM_E_CUDA, "step server: %s", cudaGetErrorString(e));
  return ARMSIM_OK;
}

int armsim_host_server(ArmSim* s, int32_t idle_us) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_host_server: null handle");
  if (idle_us < 0 || idle_us > 1000000) return fail(ARMSIM_E_INVALID, "armsim_host_server: idle_us must be in [0, 1000000]");
  if (idle_us > 0 && (!s->zero_copy || s->cfg.mode != ARMSIM_MODE_IK_TELEPORT))
    return fail(ARMSIM_E_INVALID, "armsim_host_server: needs the zero-copy host path (n_envs <= 65536) in IK-teleport mode");
  DeviceGuard guard(s->cfg.device);
  int rc = server_quiesce(s);
  if (rc) return rc;
  s->server_enabled = idle_us > 0;
  s->server_idle_ns = (unsigned long long)idle_us * 1000ull;
  return ARMSIM_OK;
}

// First half of the host step: stage the actions (unless they already sit in the pinned block) and put the fused
// launch in flight.  Zero-copy path: one graph launch, nothing else; DMA path (n > 65536): H2D + launch + D2H queued.
static int host_step_submit(ArmSim* s, const float* action_host) {
  const size_t n = (size_t)s->n;
  char* h_out = s->h_pin + s->act_bytes;
  char* d_out = s->d_io + s->act_bytes;
  if (s->host_pending) return fail(ARMSIM_E_STATE, "armsim_step_host_async: the previous step has not been waited for");
  // callers that work in the handle's own pinned block (armsim_host_buffers) skip the staging memcpys
  if ((const char*)action_host != s->h_pin) memcpy(s->h_pin, action_host, n * s->act_dim * 4);
  if (s->zero_copy && s->server_enabled) {
    // resident kernel: no CUDA call on the fast path -- one release store of the step's sequence number
    const double t = now_s();
    if (!s->server_live || (t - s->server_last_use) * 1e9 > 0.5 * (double)s->server_idle_ns) {
      if (s->server_live && cudaStreamQuery(s->stream) == cudaSuccess) s->server_live = false;   // it timed out meanwhile
      if (!s->server_live) {
        int rc = server_launch(s);
        if (rc) return rc;
      }
    }
    s->server_last_use = t;
    ++s->seq;
    __atomic_store_n((unsigned int*)s->h_cmd, s->seq, __ATOMIC_RELEASE);
  } else if (s->zero_copy) {
    HostNotify H{s->d_cta_seq, (unsigned int*)s->h_flags};
    float* o = (float*)(h_out + s->off_obs);
    float* r = (float*)(h_out + s->off_reward);
    uint8_t *d = (uint8_t*)(h_out + s->off_done), *su = (uint8_t*)(h_out + s->off_success);
    if (!s->host_graph_tried) {   // first call: record the launch once; its parameters are the same for every later step
      s->host_graph_tried = true;
      cudaGraph_t g = nullptr;
      if (getenv("ARMSIM_HOST_GRAPH") == nullptr || getenv("ARMSIM_HOST_GRAPH")[0] != '0') {
        const int64_t launches0 = s->launches;
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          const int rc0 = launch_step(s, (const float*)s->h_pin, o, r, d, su, s->stream, H);
          if (cudaStreamEndCapture(s->stream, &g) != cudaSuccess || rc0 != ARMSIM_OK ||
              cudaGraphInstantiate(&s->host_graph, g, 0) != cudaSuccess)
            s->host_graph = nullptr;
          if (g) cudaGraphDestroy(g);
        }
        s->launches = launches0;
        cudaGetLastError();
      }
    }
    if (s->host_graph) {
      CU(cudaGraphLaunch(s->host_graph, s->stream));
      s->launches += 1;
    } else {
      int rc = launch_step(s, (const float*)s->h_pin, o, r, d, su, s->stream, H);
      if (rc) return rc;
    }
    ++s->seq;
  } else {
    CU(cudaMemcpyAsync(s->d_io, s->h_pin, n * s->act_dim * 4, cudaMemcpyHostToDevice, s->stream));
    int rc = launch_step(s, (const float*)s->d_io, (float*)(d_out + s->off_obs), (float*)(d_out + s->off_reward),
                         (uint8_t*)(d_out + s->off_done), (uint8_t*)(d_out + s->off_success), s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_out, d_out, s->out_bytes, cudaMemcpyDeviceToHost, s->stream));
  }
  s->host_pending = true;
  return ARMSIM_OK;
}

// Second half: block until the results are readable in the pinned block, then hand them to the caller's buffers.
static int host_step_wait(ArmSim* s, float* obs_host, float* reward_host, uint8_t* done_host, uint8_t* success_host) {
  const size_t n = (size_t)s->n;
  char* h_out = s->h_pin + s->act_bytes;
  if (!s->host_pending) return fail(ARMSIM_E_STATE, "armsim_step_host_wait: no step in flight");
  s->host_pending = false;
  if (s->zero_copy) {
    int rc = wait_doorbells(s, s->seq);
    if (rc) return rc;
  } else {
    CU(cudaStreamSynchronize(s->stream));
  }
  if ((char*)obs_host != h_out + s->off_obs) memcpy(obs_host, h_out + s->off_obs, n * s->obs_dim * 4);
  if ((char*)reward_host != h_out + s->off_reward) memcpy(reward_host, h_out + s->off_reward, n * 4);
  if ((char*)done_host != h_out + s->off_done) memcpy(done_host, h_out + s->off_done, n);
  if ((char*)success_host != h_out + s->off_success) memcpy(success_host, h_out + s->off_success, n);
  return ARMSIM_OK;
}

int armsim_step_host(ArmSim* s, const float* action_host, float* obs_host, float* reward_host, uint8_t* done_host,
                     uint8_t* success_host) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_step_host: null handle");
  if (!action_host || !obs_host || !reward_host || !done_host || !success_host) return fail(ARMSIM_E_INVALID, "armsim_step_host: null buffer");
  CU(cudaSetDevice(s->cfg.device));
  int rc = host_step_submit(s, action_host);
  if (rc) return rc;
  return host_step_wait(s, obs_host, reward_host, done_host, success_host);
}

int armsim_step_host_async(ArmSim* s, const float* action_host) {
  if (!s || !action_host) return fail(ARMSIM_E_INVALID, "armsim_step_host_async: null argument");
  CU(cudaSetDevice(s->cfg.device));
  return host_step_submit(s, action_host);
}

int armsim_step_host_wait(ArmSim* s, float* obs_host, float* reward_host, uint8_t* done_host, uint8_t* success_host) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_step_host_wait: null handle");
  if (!obs_host || !reward_host || !done_host || !success_host) return fail(ARMSIM_E_INVALID, "armsim_step_host_wait: null buffer");
  return host_step_wait(s, obs_host, reward_host, done_host, success_host);
}

int armsim_explore(ArmSim* s, const float* actor_out_dev, float noise_std, float clip, float* action_out_dev, void* stream) {
  if (!s || !actor_out_dev || !action_out_dev) return fail(ARMSIM_E_INVALID, "armsim_explore: null argument");
  ARMSIM_QUIESCE(s);
  if (!(noise_std >= 0.0f)) return fail(ARMSIM_E_INVALID, "armsim_explore: noise_std must be >= 0");
  DeviceGuard guard(s->cfg.device);
  explore_kernel<<<s->grid, LANE_BLOCK, 0, (cudaStream_t)stream>>>(s->task, s->S, s->act_dim, actor_out_dev, noise_std, clip, action_out_dev);
  s->launches += 1;
  CU(cudaGetLastError());
  return ARMSIM_OK;
}

int armsim_policy_act(ArmSim* s, const float* obs_dev, const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* w3, const float* b3, int32_t hidden, float action_bound, float noise_std, float clip,
                      float* action_out_dev, void* stream) {
  if (!s || !obs_dev || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !action_out_dev) return fail(ARMSIM_E_INVALID, "armsim_policy_act: null argument");
  ARMSIM_QUIESCE(s);
  if (hidden != POLICY_H) return fail(ARMSIM_E_INVALID, "armsim_policy_act: hidden must be %d (got %d)", POLICY_H, hidden);
  if (s->obs_dim > POLICY_MAX_S || s->act_dim > POLICY_MAX_A) return fail(ARMSIM_E_INVALID, "armsim_policy_act: obs_dim %d / action_dim %d too wide", s->obs_dim, s->act_dim);
  DeviceGuard guard(s->cfg.device);
  const PolicyParams P{w1, b1, w2, b2, w3, b3, s->obs_dim, s->act_dim, action_bound};
  const int grid = (s->n + POLICY_ROWS - 1) / POLICY_ROWS;
  if (noise_std >= 0.0f) {
    ensure_smem(policy_mlp_kernel<true>, POLICY_SMEM);
    policy_mlp_kernel<true><<<grid, POLICY_THREADS, POLICY_SMEM, (cudaStream_t)stream>>>(s->task, s->S, s->n, P, obs_dev, noise_std, clip, action_out_dev);
  } else {
    ensure_smem(policy_mlp_kernel<false>, POLICY_SMEM);
    policy_mlp_kernel<false><<<grid, POLICY_THREADS, POLICY_SMEM, (cudaStream_t)stream>>>(s->task, s->S, s->n, P, obs_dev, 0.f, clip, action_out_dev);
  }
  s->launches += 1;
  CU(cudaGetLastError());
  return ARMSIM_OK;
}

int armsim_track_episodes(ArmSim* s, const float* reward_dev, const uint8_t* done_dev, const uint8_t* success_dev, void* stream) {
  if (!s || !reward_dev || !done_dev || !success_dev) return fail(ARMSIM_E_INVALID, "armsim_track_episodes: null argument");
  ARMSIM_QUIESCE(s);
  DeviceGuard guard(s->cfg.device);
  track_episodes_kernel<<<s->grid, LANE_BLOCK, 0, (cudaStream_t)stream>>>(s->n, s->S, reward_dev, done_dev, success_dev, s->d_stats);
  s->launches += 1;
  CU(cudaGetLastError());
  return ARMSIM_OK;
}

int armsim_episode_stats(ArmSim* s, double out[3]) {
  if (!s || !out) return fail(ARMSIM_E_INVALID, "armsim_episode_stats: null argument");
  ARMSIM_QUIESCE(s);
  CU(cudaSetDevice(s->cfg.device));
  unsigned long long h[3];
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h, s->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
  out[0] = (double)h[0];
  out[1] = (double)h[1];
  out[2] = (double)(long long)h[2] / 65536.0;
  return ARMSIM_OK;
}

int armsim_set_episode_stats(ArmSim* s, const double in[3]) {
  if (!s || !in) return fail(ARMSIM_E_INVALID, "armsim_set_episode_stats: null argument");
  ARMSIM_QUIESCE(s);
  CU(cudaSetDevice(s->cfg.device));
  const unsigned long long h[3] = {(unsigned long long)llround(in[0]), (unsigned long long)llround(in[1]),
                                   (unsigned long long)llround(in[2] * 65536.0)};
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(s->d_stats, h, sizeof(h), cudaMemcpyHostToDevice));
  return ARMSIM_OK;
}

int armsim_reset_host(ArmSim* s, const uint8_t* mask_host, float* obs_host) {
  if (!s) return fail(ARMSIM_E_INVALID, "armsim_reset_host: null handle");
  ARMSIM_QUIESCE(s);
  CU(cudaSetDevice(s->cfg.device));
  const size_t n = (size_t)s->n;
  char* h_out = s->h_pin + s->act_bytes;
  char* d_out = s->d_io + s->act_bytes;
  uint8_t* mask_dev = nullptr;
  if (mask_host) {  // the done slot of the output block doubles as the mask upload
    memcpy(h_out + s->off_done, mask_host, n);
    CU(cudaMemcpyAsync(d_out + s->off_done, h_out + s->off_done, n, cudaMemcpyHostToDevice, s->stream));
    mask_dev = (uint8_t*)(d_out + s->off_done);
  }
  float* obs_dev = (float*)(d_out + s->off_obs);
  if (mask_host && obs_host) {  // keep rows of un-reset envs as the caller passed them
    memcpy(h_out + s->off_obs, obs_host, n * s->obs_dim * 4);
    CU(cudaMemcpyAsync(obs_dev, h_out + s->off_obs, n * s->obs_dim * 4, cudaMemcpyHostToDevice, s->stream));
  }
  int rc = launch_reset(s, mask_dev, obs_dev, s->stream);
  if (rc) return rc;
  if (obs_host) CU(cudaMemcpyAsync(h_out + s->off_obs, obs_dev, n * s->obs_dim * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (obs_host) memcpy(obs_host, h_out + s->off_obs, n * s->obs_dim * 4);
  return ARMSIM_OK;
}

int armsim_set_state(ArmSim* s, int32_t field, const void* host_src, size_t bytes) {
  if (!s || !host_src) return fail(ARMSIM_E_INVALID, "armsim_set_state: null argument");
  ARMSIM_QUIESCE(s);
  const int w = field_width(field);
  if (w < 0 || field == ARMSIM_F_IK_ITERS) return fail(ARMSIM_E_STATE, "armsim_set_state: field %d not writable", field);
  const size_t n = (size_t)s->n;
  if (bytes != n * w * 4) return fail(ARMSIM_E_STATE, "armsim_set_state: field %d expects %zu bytes, got %zu", field, n * w * 4, bytes);
  CU(cudaSetDevice(s->cfg.device));
  std::vector<uint32_t> soa(n * w);
  const uint32_t* src = (const uint32_t*)host_src;
  for (size_t e = 0; e < n; ++e)
    for (int k = 0; k < w; ++k) soa[(size_t)k * n + e] = src[e * w + k];
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(field_ptr(s, field), soa.data(), bytes, cudaMemcpyHostToDevice));
  if (field == ARMSIM_F_Q || field == ARMSIM_F_STEP) CU(cudaMemset(s->S.done, 0, n));  // injected envs are live again
  return ARMSIM_OK;
}

int armsim_get_state(ArmSim* s, int32_t field, void* host_dst, size_t bytes) {
  if (!s || !host_dst) return fail(ARMSIM_E_INVALID, "armsim_get_state: null argument");
  ARMSIM_QUIESCE(s);
  const int w = field_width(field);
  if (w < 0) return fail(ARMSIM_E_STATE, "armsim_get_state: unknown field %d", field);
  const size_t n = (size_t)s->n;
  if (bytes != n * w * 4) return fail(ARMSIM_E_STATE, "armsim_get_state: field %d expects %zu bytes, got %zu", field, n * w * 4, bytes);
  CU(cudaSetDevice(s->cfg.device));
  std::vector<uint32_t> soa(n * w);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(soa.data(), field_ptr(s, field), bytes, cudaMemcpyDeviceToHost));
  uint32_t* dst = (uint32_t*)host_dst;
  for (size_t e = 0; e < n; ++e)
    for (int k = 0; k < w; ++k) dst[e * w + k] = soa[(size_t)k * n + e];
  return ARMSIM_OK;
}

int armsim_fk_host(ArmSim* s, const float* q_host, int32_t n, float* pos_host, float* rot_host) {
  if (!s || !q_host || !pos_host || n <= 0) return fail(ARMSIM_E_INVALID, "armsim_fk_host: bad argument");
  ARMSIM_QUIESCE(s);
  CU(cudaSetDevice(s->cfg.device));
  if (n > s->fk_cap) {                       // persistent staging: the Env shims call this on every first reset()
    if (s->fk_scratch) cudaFree(s->fk_scratch);
    s->fk_scratch = nullptr; s->fk_cap = 0;
    const int cap = n < 64 ? 64 : n;
    CU(cudaMalloc((void**)&s->fk_scratch, (size_t)cap * 19 * 4));
    s->fk_cap = cap;
  }
  float *dq = s->fk_scratch, *dp = dq + (size_t)s->fk_cap * 7, *dr = dp + (size_t)s->fk_cap * 3;
  cudaMemcpyAsync(dq, q_host, (size_t)n * 7 * 4, cudaMemcpyHostToDevice, s->stream);
  const int g = (n + 127) / 128;
  if (s->cfg.robot == ARMSIM_ROBOT_KUKA_IIWA) fk_kernel<ARMSIM_ROBOT_KUKA_IIWA><<<g, 128, 0, s->stream>>>(s->chain, n, dq, dp, rot_host ? dr : nullptr);
  else if (s->cfg.robot == ARMSIM_ROBOT_DIANA_S1) fk_kernel<ARMSIM_ROBOT_DIANA_S1><<<g, 128, 0, s->stream>>>(s->chain, n, dq, dp, rot_host ? dr : nullptr);
  else fk_kernel<ARMSIM_ROBOT_CUSTOM><<<g, 128, 0, s->stream>>>(s->chain, n, dq, dp, rot_host ? dr : nullptr);
  s->launches += 1;
  cudaMemcpyAsync(pos_host, dp, (size_t)n * 3 * 4, cudaMemcpyDeviceToHost, s->stream);
  if (rot_host) cudaMemcpyAsync(rot_host, dr, (size_t)n * 9 * 4, cudaMemcpyDeviceToHost, s->stream);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) return fail(ARMSIM_E_CUDA, "armsim_fk_host: %s", cudaGetErrorString(e));
  return ARMSIM_OK;
}

}  // extern "C"
