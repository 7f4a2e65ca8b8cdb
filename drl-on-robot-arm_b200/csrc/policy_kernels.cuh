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
// Layer 2 (4096 x 256 x 256 MACs, 97 % of the work) runs on the tensor cores in the 3xTF32 form: every fp32 operand is
// split x = hi + lo (hi = x rounded to TF32, lo = x - hi, exact) and  a.b ~ a_hi.b_hi + a_hi.b_lo + a_lo.b_hi  with
// fp32 accumulation -- the dropped a_lo.b_lo term is 2^-22 relative, so the result matches PyTorch's fp32 module to a
// few 1e-7 (the rollout must act with the policy that is being trained; allow_tf32 is off in PyTorch, and a single TF32
// pass would be 1e-3 off).  Measured on B200 at 4096 envs: PyTorch module + armsim_explore 44 us, an all-FFMA version
// of this kernel 32 us (FFMA issues once per ~1.5-2 cycles per sub-partition, tools/micro/ffma_rate.cu).
// `mma.sync` (m16n8k8), not tcgen05: one 32 x 256 x 256 problem per block is far below a UMMA tile pipeline's set-up
// cost, and the fragments come straight from the transposed activations this kernel already keeps in shared memory.
// Mapping: a block of 512 threads (16 warps: four per scheduler to hide LDS / mma latency) owns 32 batch rows.
//   layer 1: thread (j, half) = hidden unit j, 16 of the 32 rows x S inputs (FFMA), relu, split -> hi[j][.], lo[j][.]
//            (row r of a column sits at position (r & 7) * 4 + (r >> 3): the four rows g, g+8, g+16, g+24 one lane
//            needs for its two A fragments are one LDS.128)
//   layer 2: warp w owns columns 16w..16w+15 of all 32 rows (2 x 2 mma tiles); fc2 streams through a 3-stage cp.async
//            pipeline of [256 out][32 in] tiles (row stride 36 floats: B fragments are conflict-free straight from
//            nn.Linear's [out, in] layout, no transpose), split hi / lo per fragment in registers
//   layer 3: 32 x A threads, one 256-long fp32 dot product each, tanh, bound, (noise, clip) -> action [n, A]
#pragma once
#include "armsim_device.cuh"

constexpr int POLICY_H = 256;        // hidden width (opt.hidden_dim of the reference, config.py)
constexpr int POLICY_ROWS = 32;      // batch rows per block
constexpr int POLICY_KC = 32;        // layer-2 inputs per staged tile
constexpr int POLICY_HS = 40;        // row stride of the transposed activations (floats): A fragments conflict-free
constexpr int POLICY_WS = 36;        // row stride of a weight tile [out][KC] (floats): B fragments conflict-free
constexpr int POLICY_STAGES = 3;
constexpr int POLICY_THREADS = 512;
constexpr int POLICY_MAX_S = 16, POLICY_MAX_A = 4;
constexpr size_t POLICY_SMEM = (size_t)(2 * POLICY_H * POLICY_HS + POLICY_STAGES * POLICY_H * POLICY_WS + POLICY_ROWS * POLICY_MAX_S +
                                        POLICY_MAX_A * POLICY_H) * sizeof(float);

struct PolicyParams {
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  int S, A;
  float bound;
};

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// D += A (16x8, row) * B (8x8, col), TF32 operands, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

template <bool EXPLORE>
__global__ void __launch_bounds__(POLICY_THREADS, 1)
policy_mlp_kernel(const __grid_constant__ TaskParams T, const StatePtrs St, int n, const PolicyParams P,
                  const float* __restrict__ obs, float noise_std, float clip, float* __restrict__ out) {
  extern __shared__ __align__(16) float policy_smem[];
  float* hhi = policy_smem;                                  // [H][HS]  layer-1 activations, TF32-rounded; later h2 (fp32)
  float* hlo = hhi + POLICY_H * POLICY_HS;                   // [H][HS]  their remainders
  float* wt = hlo + POLICY_H * POLICY_HS;                    // [STAGES][H out][WS]  fc2 tiles as PyTorch stores them
  float* xs = wt + POLICY_STAGES * POLICY_H * POLICY_WS;     // [ROWS][MAX_S]
  float* w3s = xs + POLICY_ROWS * POLICY_MAX_S;              // [MAX_A][H]
  const int tid = threadIdx.x;
  const int base = blockIdx.x * POLICY_ROWS;
  const int S = P.S, A = P.A;
  constexpr int NT = POLICY_H / POLICY_KC;

  // fc2 tile c -> stage c % STAGES: two threads per output row, each copies 64 of the 128 bytes W2[row][32c .. 32c+31]
  // with four 16-byte cp.async
  const int wrow = tid >> 1, whalf = tid & 1;
  auto issue_tile = [&](int c) {
    if (c < NT) {
      const float* src = P.w2 + (size_t)wrow * POLICY_H + c * POLICY_KC + whalf * 16;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(wt + ((c % POLICY_STAGES) * POLICY_H + wrow) * POLICY_WS + whalf * 16);
#pragma unroll
      for (int i = 0; i < POLICY_KC / 8; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 4 * i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_tile(0);
  issue_tile(1);

  // ---- stage the obs tile (rows past n read as zero) and fc3; this thread's fc1 row into registers
  const int hu = tid & (POLICY_H - 1), rhalf = tid >> 8;        // layer 1: hidden unit, which 16 rows
  float w1r[POLICY_MAX_S];
#pragma unroll
  for (int s = 0; s < POLICY_MAX_S; ++s) w1r[s] = s < S ? __ldg(P.w1 + (size_t)hu * S + s) : 0.f;
  const float b1v = __ldg(P.b1 + hu);
  for (int i = tid; i < POLICY_ROWS * S; i += POLICY_THREADS) {
    const int r = i / S, s = i - r * S;
    xs[r * POLICY_MAX_S + s] = (base + r < n) ? __ldg(obs + (size_t)base * S + i) : 0.f;
  }
  for (int i = tid; i < A * POLICY_H; i += POLICY_THREADS) w3s[i] = __ldg(P.w3 + i);
  __syncthreads();

  // ---- layer 1: hidden unit hu, rows 16 * rhalf .. + 15 (= positions 4g + 2 rhalf, + 1 of every g)
  {
    float acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = b1v;
    for (int s = 0; s < S; ++s) {
      const float w = w1r[0];
#pragma unroll
      for (int r = 0; r < 16; ++r) acc[r] = fmaf(xs[(rhalf * 16 + r) * POLICY_MAX_S + s], w, acc[r]);
#pragma unroll
      for (int q = 0; q + 1 < POLICY_MAX_S; ++q) w1r[q] = w1r[q + 1];      // rotate: the next input's weight to slot 0
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {                     // rows 16 rhalf + g and + g + 8 -> positions 4g + 2 rhalf, + 1
      const float x0 = fmaxf(acc[g], 0.f), x1 = fmaxf(acc[g + 8], 0.f);
      const float h0 = tf32_round(x0), h1 = tf32_round(x1);
      *reinterpret_cast<float2*>(hhi + hu * POLICY_HS + 4 * g + 2 * rhalf) = make_float2(h0, h1);
      *reinterpret_cast<float2*>(hlo + hu * POLICY_HS + 4 * g + 2 * rhalf) = make_float2(x0 - h0, x1 - h1);
    }
  }

  // ---- layer 2 on the tensor cores: warp w -> columns 16w .. 16w+15, rows 0..31 = 2 (M) x 2 (N) mma tiles
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  float acc[2][2][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][q][i] = 0.f;
#pragma unroll 1
  for (int c = 0; c < NT; ++c) {
    asm volatile("cp.async.wait_group %0;" ::"n"(POLICY_STAGES - 2) : "memory");      // this thread's part of tile c landed
    __syncthreads();            // everybody's part landed (c = 0: layer 1 too); stage (c + 2) % 3 is no longer being read
    issue_tile(c + POLICY_STAGES - 1);
    const float* wb = wt + ((c % POLICY_STAGES) * POLICY_H + warp * 16 + g) * POLICY_WS + t4;
#pragma unroll
    for (int ks = 0; ks < POLICY_KC / 8; ++ks) {
      const int k0 = c * POLICY_KC + ks * 8;
      // rows (g, g+8, g+16, g+24) at k0+t4 and at k0+t4+4: fragment registers a0 a1 (a2 a3) of M tile 0 and 1
      const float4 h0 = *reinterpret_cast<const float4*>(hhi + (k0 + t4) * POLICY_HS + 4 * g);
      const float4 h1 = *reinterpret_cast<const float4*>(hhi + (k0 + t4 + 4) * POLICY_HS + 4 * g);
      const float4 l0 = *reinterpret_cast<const float4*>(hlo + (k0 + t4) * POLICY_HS + 4 * g);
      const float4 l1 = *reinterpret_cast<const float4*>(hlo + (k0 + t4 + 4) * POLICY_HS + 4 * g);
      const float ahi[2][4] = {{h0.x, h0.y, h1.x, h1.y}, {h0.z, h0.w, h1.z, h1.w}};
      const float alo[2][4] = {{l0.x, l0.y, l1.x, l1.y}, {l0.z, l0.w, l1.z, l1.w}};
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float b0 = wb[q * 8 * POLICY_WS + ks * 8], b1 = wb[q * 8 * POLICY_WS + ks * 8 + 4];
        const float b0h = tf32_round(b0), b1h = tf32_round(b1);
        const float b0l = b0 - b0h, b1l = b1 - b1h;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#ifndef POLICY_SINGLE_PASS                                   // (tuning experiment: one TF32 pass, 1e-3 accuracy)
          mma_tf32(acc[m][q], alo[m], b0h, b1h);            // small terms first
          mma_tf32(acc[m][q], ahi[m], b0l, b1l);
#endif
          mma_tf32(acc[m][q], ahi[m], b0h, b1h);
        }
      }
    }
  }
  __syncthreads();              // every warp is past its last fragment read of hhi
  // bias + relu, back into hhi as fp32 h2[col][position(row)].
  // accumulator fragment: c0 (row g, col 2t), c1 (row g, col 2t+1), c2 (row g+8, col 2t), c3 (row g+8, col 2t+1)
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int col = warp * 16 + q * 8 + 2 * t4;
    const float bb0 = __ldg(P.b2 + col), bb1 = __ldg(P.b2 + col + 1);
#pragma unroll
    for (int m = 0; m < 2; ++m) {                     // rows 16m + g -> position 4g + 2m, rows 16m + g + 8 -> 4g + 2m + 1
      float* d0 = hhi + col * POLICY_HS + 4 * g + 2 * m;
      *reinterpret_cast<float2*>(d0) = make_float2(fmaxf(acc[m][q][0] + bb0, 0.f), fmaxf(acc[m][q][2] + bb0, 0.f));
      *reinterpret_cast<float2*>(d0 + POLICY_HS) = make_float2(fmaxf(acc[m][q][1] + bb1, 0.f), fmaxf(acc[m][q][3] + bb1, 0.f));
    }
  }
  __syncthreads();
  const float* hT = hhi;

  // ---- layer 3 + tanh * bound (+ exploration noise): thread (a, r) = (tid / 32, tid % 32)
  const int r = tid & 31, a = tid >> 5;
  const int rp = (r & 7) * 4 + (r >> 3);              // where row r sits inside a column of hT
  const int e = base + r;
  const bool mine = a < A && e < n;
  float act = 0.f;
  unsigned int draw = 0u;
  if (mine) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float* w = w3s + a * POLICY_H;
#pragma unroll 4
    for (int k = 0; k < POLICY_H; k += 4) {
      s0 = fmaf(hT[(k + 0) * POLICY_HS + rp], w[k + 0], s0); s1 = fmaf(hT[(k + 1) * POLICY_HS + rp], w[k + 1], s1);
      s2 = fmaf(hT[(k + 2) * POLICY_HS + rp], w[k + 2], s2); s3 = fmaf(hT[(k + 3) * POLICY_HS + rp], w[k + 3], s3);
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
