// ffma_rate.cu -- issue-rate microbenchmark behind DESIGN's "one warp per scheduler" analysis: cycles per warp-level
// instruction for independent FFMA (3 register operands, no reuse / with a reused operand), packed fma.rn.f32x2 and
// FMNMX streams, with W warps per SM sub-partition.   nvcc -arch=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<unsigned long long&>(d))
               : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float s0, float s1) {
  float a[16], b[16], c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = s0 + i + threadIdx.x; b[i] = s1 - i + threadIdx.x * 0.001f; c[i] = i * 0.5f + threadIdx.x; }
  float2 p[8], q[8], r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = make_float2(a[2 * i], a[2 * i + 1]); q[i] = make_float2(b[2 * i], b[2 * i + 1]); r[i] = make_float2(c[2 * i], c[2 * i + 1]); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 16 independent FFMA, three distinct registers each
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fmaf(a[i], b[i], c[i]);
    } else if (MODE == 1) {   // one multiplicand shared by 4 consecutive FFMA (what a register-tiled GEMM issues)
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fmaf(a[i / 4], b[i % 4], c[i]);
    } else if (MODE == 2) {   // 8 packed f32x2 FMAs = 16 FMAs per lane
#pragma unroll
      for (int i = 0; i < 8; ++i) ffma2(r[i], p[i], q[i]);
    } else if (MODE == 3) {   // 16 FMNMX (alu pipe)
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fmaxf(a[i], fminf(b[i], c[i]));
    } else if (MODE == 4) {   // dependent FFMA chain
#pragma unroll
      for (int i = 0; i < 16; ++i) c[0] = fmaf(a[i], b[i], c[0]);
    } else if (MODE == 5) {   // 8 FFMA + 8 FMNMX interleaved (two pipes)
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = fmaf(a[i], b[i], c[i]); c[8 + i] = fmaxf(a[8 + i], c[8 + i]); }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int warps = 1; warps <= 4; warps *= 2) {
    const int iters = 4096;
    k<MODE><<<148, 128 * warps>>>(out, cyc, iters, 1.0f, 2.0f);
    k<MODE><<<148, 128 * warps>>>(out, cyc, iters, 1.0f, 2.0f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SMSP %d : %.2f cycles per warp-instruction (%.2f per SMSP issue slot)\n", name, warps,
           (double)c / iters / per_iter, (double)c / iters / per_iter / warps);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA 3 distinct regs, independent", 16);
  run<1>("FFMA shared multiplicand (reuse)", 16);
  run<2>("fma.rn.f32x2 (2 FMAs per lane per instr)", 8);
  run<3>("FMNMX pairs", 32);
  run<4>("FFMA dependent chain", 16);
  run<5>("FFMA + FMNMX interleaved", 16);
  return 0;
}
