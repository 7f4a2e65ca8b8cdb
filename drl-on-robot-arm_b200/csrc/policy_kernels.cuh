// policy_kernels.cuh -- the acting policy of the rollout as ONE launch (sm_100a, fp32).
//
// What it replaces: `action = agent.take_action(state)` + exploration noise of the reference's rollout loops
// (main.py:196-200, :116-117) where take_action is PolicyNet.forward (algo/TD3/net_mlp.py:29-40):
//     a = tanh(fc3(relu(fc2(relu(fc1(s)))))) * action_bound,      fc1 [H,S], fc2 [H,H], fc3 [A,H], H = 256.
// Issued through PyTorch that is three SIMT sgemms + seven bias / activation kernels per lockstep step (~45 us for 4096
// envs on a B200, 11x the fused env step it feeds); here it is one kernel that reads the nn.Linear parameters where
// PyTorch keeps them (row-major [out, in], fp32) and optionally adds the exploration noise of armsim_explore in its
// epilogue.  The agents' update (algo/) stays PyTorch; this is the inference side of the rollout only.
//
// fp32 FFMA on purpose: 4096 x 256 x 256 MACs is GEMM-shaped, but the rollout must act with the SAME policy PyTorch
// trains (fp32 master weights, allow_tf32 off); one TF32 tensor-core pass is 1e-3 off and a 3xTF32 split costs what the
// FFMA form costs at this size.  Mapping: a block of 256 threads owns 32 batch rows.
//   layer 1: thread j = hidden unit j, 32 rows x S inputs from a shared obs tile         -> hT[j][row]  (transposed)
//   layer 2: 8 rows x 4 columns of accumulators per thread; fc2 is staged in 32-input tiles, transposed on the way into
//            shared memory (wt[k][out]) and double-buffered through registers; per k: 2 broadcast LDS.128 (8 row
//            values) + 1 LDS.128 (4 weights) feed 32 FFMA                                -> hT[col][row]
//   layer 3: 32 x A threads, one 256-long dot product each, tanh, bound, (noise, clip)   -> action [n, A]
#pragma once
#include "armsim_device.cuh"

constexpr int POLICY_H = 256;        // hidden width (opt.hidden_dim of the reference, config.py)
constexpr int POLICY_ROWS = 32;      // batch rows per block
constexpr int POLICY_KC = 32;        // layer-2 inputs per staged tile
constexpr int POLICY_HS = 36;        // row stride of hT in floats: 16-byte aligned rows, conflict-free 128-bit stores
constexpr int POLICY_MAX_S = 16, POLICY_MAX_A = 4;
constexpr size_t POLICY_SMEM = (size_t)(POLICY_H * POLICY_HS + 2 * POLICY_KC * POLICY_H + POLICY_ROWS * POLICY_MAX_S +
                                        POLICY_MAX_A * POLICY_H) * sizeof(float);

struct PolicyParams {
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  int S, A;
  float bound;
};

template <bool EXPLORE>
__global__ void __launch_bounds__(POLICY_H, 1)
policy_mlp_kernel(const __grid_constant__ TaskParams T, const StatePtrs St, int n, const PolicyParams P,
                  const float* __restrict__ obs, float noise_std, float clip, float* __restrict__ out) {
  extern __shared__ __align__(16) float policy_smem[];
  float* hT = policy_smem;                                   // [H][HS]
  float* wt = hT + POLICY_H * POLICY_HS;                     // [2][KC][H]
  float* xs = wt + 2 * POLICY_KC * POLICY_H;                 // [ROWS][MAX_S]
  float* w3s = xs + POLICY_ROWS * POLICY_MAX_S;              // [MAX_A][H]
  const int tid = threadIdx.x;
  const int base = blockIdx.x * POLICY_ROWS;
  const int S = P.S, A = P.A;

  // ---- stage the obs tile (rows past n read as zero) and fc3
  for (int i = tid; i < POLICY_ROWS * S; i += POLICY_H) {
    const int r = i / S, s = i - r * S;
    xs[r * POLICY_MAX_S + s] = (base + r < n) ? __ldg(obs + (size_t)base * S + i) : 0.f;
  }
  for (int i = tid; i < A * POLICY_H; i += POLICY_H) w3s[i] = __ldg(P.w3 + i);
  // first fc2 tile on its way while layer 1 runs: thread t owns output row t, 32 consecutive inputs = 8 x 16 bytes
  float4 pre[POLICY_KC / 4];
  {
    const float4* src = reinterpret_cast<const float4*>(P.w2 + (size_t)tid * POLICY_H);
#pragma unroll
    for (int i = 0; i < POLICY_KC / 4; ++i) pre[i] = __ldg(src + i);
  }
  __syncthreads();

  // ---- layer 1: hidden unit tid for all 32 rows
  {
    float acc[POLICY_ROWS];
    const float b = __ldg(P.b1 + tid);
#pragma unroll
    for (int r = 0; r < POLICY_ROWS; ++r) acc[r] = b;
    for (int s = 0; s < S; ++s) {
      const float w = __ldg(P.w1 + (size_t)tid * S + s);
#pragma unroll
      for (int r = 0; r < POLICY_ROWS; ++r) acc[r] = fmaf(xs[r * POLICY_MAX_S + s], w, acc[r]);
    }
    float4* dst = reinterpret_cast<float4*>(hT + tid * POLICY_HS);
#pragma unroll
    for (int r = 0; r < POLICY_ROWS / 4; ++r)
      dst[r] = make_float4(fmaxf(acc[4 * r], 0.f), fmaxf(acc[4 * r + 1], 0.f), fmaxf(acc[4 * r + 2], 0.f), fmaxf(acc[4 * r + 3], 0.f));
  }
  // tile 0 into buffer 0 (transposed: wt[k][out])
#pragma unroll
  for (int i = 0; i < POLICY_KC / 4; ++i) {
    wt[(4 * i + 0) * POLICY_H + tid] = pre[i].x; wt[(4 * i + 1) * POLICY_H + tid] = pre[i].y;
    wt[(4 * i + 2) * POLICY_H + tid] = pre[i].z; wt[(4 * i + 3) * POLICY_H + tid] = pre[i].w;
  }
  __syncthreads();

  // ---- layer 2: rows rg*8 .. +8, columns cg*4 .. +4
  const int cg = tid & 63, rg = tid >> 6;
  float acc[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
  constexpr int NT = POLICY_H / POLICY_KC;
#pragma unroll 1
  for (int c = 0; c < NT; ++c) {
    if (c + 1 < NT) {
      const float4* src = reinterpret_cast<const float4*>(P.w2 + (size_t)tid * POLICY_H + (c + 1) * POLICY_KC);
#pragma unroll
      for (int i = 0; i < POLICY_KC / 4; ++i) pre[i] = __ldg(src + i);
    }
    const float* wb = wt + (c & 1) * (POLICY_KC * POLICY_H) + cg * 4;
    const float* hb = hT + (c * POLICY_KC) * POLICY_HS + rg * 8;
#pragma unroll 8
    for (int kk = 0; kk < POLICY_KC; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(wb + kk * POLICY_H);
      const float4 h0 = *reinterpret_cast<const float4*>(hb + kk * POLICY_HS);
      const float4 h1 = *reinterpret_cast<const float4*>(hb + kk * POLICY_HS + 4);
      const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        acc[r][0] = fmaf(hv[r], w.x, acc[r][0]); acc[r][1] = fmaf(hv[r], w.y, acc[r][1]);
        acc[r][2] = fmaf(hv[r], w.z, acc[r][2]); acc[r][3] = fmaf(hv[r], w.w, acc[r][3]);
      }
    }
    if (c + 1 < NT) {
      float* wn = wt + ((c + 1) & 1) * (POLICY_KC * POLICY_H);
#pragma unroll
      for (int i = 0; i < POLICY_KC / 4; ++i) {
        wn[(4 * i + 0) * POLICY_H + tid] = pre[i].x; wn[(4 * i + 1) * POLICY_H + tid] = pre[i].y;
        wn[(4 * i + 2) * POLICY_H + tid] = pre[i].z; wn[(4 * i + 3) * POLICY_H + tid] = pre[i].w;
      }
    }
    __syncthreads();
  }
  // bias + relu, back into hT as hT[col][row] (every thread is past its last read of layer-1 activations)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = cg * 4 + q;
    const float b = __ldg(P.b2 + col);
    float4* dst = reinterpret_cast<float4*>(hT + col * POLICY_HS + rg * 8);
    dst[0] = make_float4(fmaxf(acc[0][q] + b, 0.f), fmaxf(acc[1][q] + b, 0.f), fmaxf(acc[2][q] + b, 0.f), fmaxf(acc[3][q] + b, 0.f));
    dst[1] = make_float4(fmaxf(acc[4][q] + b, 0.f), fmaxf(acc[5][q] + b, 0.f), fmaxf(acc[6][q] + b, 0.f), fmaxf(acc[7][q] + b, 0.f));
  }
  __syncthreads();

  // ---- layer 3 + tanh * bound (+ exploration noise): thread (a, r) = (tid / 32, tid % 32)
  const int r = tid & 31, a = tid >> 5;
  const int e = base + r;
  const bool mine = a < A && e < n;
  float act = 0.f;
  unsigned int draw = 0u;
  if (mine) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float* w = w3s + a * POLICY_H;
#pragma unroll 4
    for (int k = 0; k < POLICY_H; k += 4) {
      s0 = fmaf(hT[(k + 0) * POLICY_HS + r], w[k + 0], s0); s1 = fmaf(hT[(k + 1) * POLICY_HS + r], w[k + 1], s1);
      s2 = fmaf(hT[(k + 2) * POLICY_HS + r], w[k + 2], s2); s3 = fmaf(hT[(k + 3) * POLICY_HS + r], w[k + 3], s3);
    }
    act = tanhf((s0 + s1) + (s2 + s3) + __ldg(P.b3 + a)) * P.bound;
    if (EXPLORE) draw = St.explore_count[e];
  }
  if (EXPLORE) {
    __syncthreads();                                   // every component has read its env's draw counter
    if (mine) {
      if (a == 0) St.explore_count[e] = draw + 1u;
      float z[4];
      explore_normals(T, T.gid_offset + (unsigned long long)e, draw, 0u, z);
      const float za = a == 0 ? z[0] : (a == 1 ? z[1] : (a == 2 ? z[2] : z[3]));
      act = fmaf(noise_std, za, act);
      if (clip > 0.0f) act = fminf(fmaxf(act, -clip), clip);
    }
  }
  if (mine) out[(size_t)e * A + a] = act;
}
