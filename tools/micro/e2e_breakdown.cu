// tools/micro/e2e_breakdown.cu -- where the host-buffer step's time goes on this box (measurement tool, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/e2e_breakdown tools/micro/e2e_breakdown.cu && gpurun_out/e2e_breakdown
// Each variant: launch -> (PCIe reads) -> (compute stand-in) -> (PCIe writes) -> doorbell in mapped host memory -> host
// poll.  Prints microseconds per round trip (median of `reps`).
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Notify { unsigned* counter; unsigned* flag; unsigned seq; };

__device__ __forceinline__ void ring(const Notify& H) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(H.counter, 1u) == gridDim.x - 1) {
      *H.counter = 0;
      __threadfence_system();
      *(volatile unsigned*)H.flag = H.seq;
    }
  }
}

// reads `nin` floats per thread-strided from `in`, spins `spin` clocks, writes `nout` floats to `out`
__global__ void __launch_bounds__(128) k_step(const float* __restrict__ in, float* __restrict__ out, int nin, int nout, int spin,
                                              Notify H) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int i = t; i < nin; i += nt) acc += in[i];
  if (spin) { const long long t0 = clock64(); while (clock64() - t0 < spin) { } }
  for (int i = t; i < nout; i += nt) out[i] = acc + (float)i;
  ring(H);
}

// per-CTA doorbells: no device-wide atomic, ONE system fence per CTA (cumulative through the CTA barrier); the host
// polls gridDim.x consecutive words.  seq comes from a device counter so the launch parameters never change (graphs).
__global__ void __launch_bounds__(128) k_step_cta(const float* __restrict__ in, float* __restrict__ out, int nin, int nout, int spin,
                                                  unsigned* flags, const unsigned* seq_dev, unsigned seq_arg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int i = t; i < nin; i += nt) acc += in[i];
  if (spin) { const long long t0 = clock64(); while (clock64() - t0 < spin) { } }
  for (int i = t; i < nout; i += nt) out[i] = acc + (float)i;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned s = seq_dev ? *seq_dev : seq_arg;
    __threadfence_system();
    *(volatile unsigned*)(flags + blockIdx.x) = s;
  }
}
__global__ void k_bump(unsigned* seq_dev) { *seq_dev += 1; }

// persistent server: waits for host doorbell `cmd` (mapped host memory) to reach seq, does one step, rings back.
// Bounded: leaves after `max_rounds` rounds or when a poll exceeds ~2 s, so it can never hang the box.
__global__ void __launch_bounds__(128) k_server(const float* __restrict__ in, float* __restrict__ out, int nin, int nout, int spin,
                                                volatile unsigned* cmd, unsigned* counter, unsigned* flag, int max_rounds) {
  __shared__ int s_quit;
  for (int round = 1; round <= max_rounds; ++round) {
    if (threadIdx.x == 0) {
      s_quit = 0;
      const long long t0 = clock64();
      while (*cmd < (unsigned)round) { if (clock64() - t0 > 4000000000ll) { s_quit = 1; break; } }
    }
    __syncthreads();
    if (s_quit) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    float acc = 0.f;
    for (int i = t; i < nin; i += nt) acc += ((const volatile float*)in)[i];
    if (spin) { const long long t0 = clock64(); while (clock64() - t0 < spin) { } }
    for (int i = t; i < nout; i += nt) out[i] = acc + (float)i;
    Notify H{counter, flag, (unsigned)round};
    ring(H);
    __syncthreads();
  }
}

// ping-pong only: ONE thread polls the host word and echoes it back (pure PCIe poll round trip)
__global__ void k_pingpong(volatile unsigned* cmd, volatile unsigned* flag, int max_rounds) {
  for (int round = 1; round <= max_rounds; ++round) {
    const long long t0 = clock64();
    while (*cmd < (unsigned)round) { if (clock64() - t0 > 4000000000ll) return; }
    *flag = (unsigned)round;
  }
}

// server v2: block 0 / thread 0 is the only poller of host memory; it relays the command through a device word that
// the other blocks poll in L2.  Per-block doorbells back to the host.
__global__ void __launch_bounds__(128) k_server2(const float* __restrict__ in, float* __restrict__ out, int nin, int nout, int spin,
                                                 volatile unsigned* cmd, volatile unsigned* relay, unsigned* flags, int max_rounds) {
  __shared__ int s_quit;
  for (int round = 1; round <= max_rounds; ++round) {
    if (threadIdx.x == 0) {
      s_quit = 0;
      const long long t0 = clock64();
      if (blockIdx.x == 0) {
        while (*cmd < (unsigned)round) { if (clock64() - t0 > 4000000000ll) { s_quit = 1; break; } }
        *relay = s_quit ? 0xffffffffu : (unsigned)round;
      } else {
        unsigned v;
        while ((v = *relay) < (unsigned)round) { if (clock64() - t0 > 6000000000ll) { s_quit = 1; break; } }
        if (v == 0xffffffffu) s_quit = 1;
      }
    }
    __syncthreads();
    if (s_quit) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    float acc = 0.f;
    for (int i = t; i < nin; i += nt) acc += __ldcv(in + i);
    if (spin) { const long long t0 = clock64(); while (clock64() - t0 < spin) { } }
    for (int i = t; i < nout; i += nt) out[i] = acc + (float)i;
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence_system(); *(volatile unsigned*)(flags + blockIdx.x) = (unsigned)round; }
  }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const int n = 4096, grid = 32, reps = argc > 1 ? atoi(argv[1]) : 2000;
  const int nin = n * 3, nout = n * 6 + n + n / 2;   // actions ; obs + reward + done/success bytes as floats
  float *h_in, *h_out, *d_in, *d_out;
  unsigned *h_flag, *d_counter, *h_cmd;
  CK(cudaSetDevice(0));
  CK(cudaHostAlloc(&h_in, nin * 4, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_out, nout * 4, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_flag, 256, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_cmd, 256, cudaHostAllocMapped));
  CK(cudaMalloc(&d_in, nin * 4)); CK(cudaMalloc(&d_out, nout * 4)); CK(cudaMalloc(&d_counter, 4));
  CK(cudaMemset(d_counter, 0, 4)); CK(cudaMemset(d_in, 0, nin * 4));
  memset(h_in, 0, nin * 4);
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const int spin = 5000;   // ~2.5 us of stand-in compute
  unsigned seq = 0;
  volatile unsigned* vf = h_flag;
  *vf = 0;

  auto bench = [&](const char* name, auto&& body) {
    std::vector<double> t(reps);
    for (int k = 0; k < 50; ++k) body();
    for (int k = 0; k < reps; ++k) { const double t0 = now(); body(); t[k] = now() - t0; }
    std::sort(t.begin(), t.end());
    printf("%-64s median %6.2f us   p10 %6.2f   p90 %6.2f\n", name, 1e6 * t[reps / 2], 1e6 * t[reps / 10], 1e6 * t[reps * 9 / 10]);
  };
  auto wait = [&](unsigned s) { while (*vf != s) { __builtin_ia32_pause(); } };

  bench("launch only (cudaLaunchKernel returns)", [&] { k_step<<<grid, 128, 0, st>>>(d_in, d_out, 0, 0, 0, Notify{d_counter, h_flag, ++seq}); });
  CK(cudaStreamSynchronize(st));
  bench("empty kernel + doorbell", [&] { k_step<<<grid, 128, 0, st>>>(d_in, d_out, 0, 0, 0, Notify{d_counter, h_flag, ++seq}); wait(seq); });
  bench("empty kernel + cudaStreamSynchronize", [&] { k_step<<<grid, 128, 0, st>>>(d_in, d_out, 0, 0, 0, Notify{d_counter, h_flag, ++seq}); cudaStreamSynchronize(st); });
  bench("compute 2.5us (device in/out) + doorbell", [&] { k_step<<<grid, 128, 0, st>>>(d_in, d_out, nin, nout, spin, Notify{d_counter, h_flag, ++seq}); wait(seq); });
  bench("zero-copy read 48K + compute + device out + doorbell", [&] { k_step<<<grid, 128, 0, st>>>(h_in, d_out, nin, nout, spin, Notify{d_counter, h_flag, ++seq}); wait(seq); });
  bench("device in + compute + zero-copy write 122K + doorbell", [&] { k_step<<<grid, 128, 0, st>>>(d_in, h_out, nin, nout, spin, Notify{d_counter, h_flag, ++seq}); wait(seq); });
  bench("zero-copy read + compute + zero-copy write + doorbell", [&] { k_step<<<grid, 128, 0, st>>>(h_in, h_out, nin, nout, spin, Notify{d_counter, h_flag, ++seq}); wait(seq); });
  bench("memcpyAsync H2D + kernel + memcpyAsync D2H + streamSync", [&] {
    cudaMemcpyAsync(d_in, h_in, nin * 4, cudaMemcpyHostToDevice, st);
    k_step<<<grid, 128, 0, st>>>(d_in, d_out, nin, nout, spin, Notify{d_counter, h_flag, ++seq});
    cudaMemcpyAsync(h_out, d_out, nout * 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st); });
  bench("zero-copy read + compute + device out + memcpyAsync D2H + sync", [&] {
    k_step<<<grid, 128, 0, st>>>(h_in, d_out, nin, nout, spin, Notify{d_counter, h_flag, ++seq});
    cudaMemcpyAsync(h_out, d_out, nout * 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st); });
  // CUDA graph of the single kernel (launch cost of a graph vs a kernel)
  {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    k_step<<<grid, 128, 0, st>>>(h_in, h_out, nin, nout, spin, Notify{d_counter, h_flag, 0x7fffffffu});
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    bench("graph launch of the zero-copy kernel + doorbell(poll any change)", [&] { *vf = 0; cudaGraphLaunch(ge, st); wait(0x7fffffffu); });
  }
  CK(cudaStreamSynchronize(st));
  {
    unsigned* h_flags; CK(cudaHostAlloc(&h_flags, 4096, cudaHostAllocMapped)); memset(h_flags, 0, 4096);
    volatile unsigned* vfl = h_flags;
    auto waitall = [&](unsigned s) { for (;;) { bool ok = true; for (int b = 0; b < grid; ++b) ok &= (vfl[b] == s); if (ok) break; __builtin_ia32_pause(); } };
    bench("per-CTA flags: empty kernel + doorbells", [&] { k_step_cta<<<grid, 128, 0, st>>>(d_in, d_out, 0, 0, 0, h_flags, nullptr, ++seq); waitall(seq); });
    bench("per-CTA flags: zero-copy read + compute + zero-copy write", [&] { k_step_cta<<<grid, 128, 0, st>>>(h_in, h_out, nin, nout, spin, h_flags, nullptr, ++seq); waitall(seq); });
    bench("per-CTA flags: device in + compute + zero-copy write", [&] { k_step_cta<<<grid, 128, 0, st>>>(d_in, h_out, nin, nout, spin, h_flags, nullptr, ++seq); waitall(seq); });
    bench("per-CTA flags: zero-copy read + compute + device out", [&] { k_step_cta<<<grid, 128, 0, st>>>(h_in, d_out, nin, nout, spin, h_flags, nullptr, ++seq); waitall(seq); });
    unsigned* d_seq; CK(cudaMalloc(&d_seq, 4)); CK(cudaMemset(d_seq, 0, 4));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    k_bump<<<1, 1, 0, st>>>(d_seq);
    k_step_cta<<<grid, 128, 0, st>>>(h_in, h_out, nin, nout, spin, h_flags, d_seq, 0);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    unsigned gs = 0;
    bench("per-CTA flags: GRAPH(bump + zero-copy step)", [&] { cudaGraphLaunch(ge, st); waitall(++gs); });
    // H2D by DMA inside the graph, zero-copy write back
    cudaGraph_t g2; cudaGraphExec_t ge2;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    k_bump<<<1, 1, 0, st>>>(d_seq);
    cudaMemcpyAsync(d_in, h_in, nin * 4, cudaMemcpyHostToDevice, st);
    k_step_cta<<<grid, 128, 0, st>>>(d_in, h_out, nin, nout, spin, h_flags, d_seq, 0);
    CK(cudaStreamEndCapture(st, &g2));
    CK(cudaGraphInstantiate(&ge2, g2, 0));
    bench("per-CTA flags: GRAPH(bump + DMA H2D + step, zero-copy write)", [&] { cudaGraphLaunch(ge2, st); waitall(++gs); });
  }
  CK(cudaStreamSynchronize(st));
  // persistent server
  {
    volatile unsigned* vc = h_cmd; *vc = 0; *vf = 0;
    const int rounds = reps + 50;
    CK(cudaMemset(d_counter, 0, 4));
    k_server<<<grid, 128, 0, st>>>(h_in, h_out, nin, nout, spin, h_cmd, d_counter, h_flag, rounds);
    unsigned r = 0;
    std::vector<double> t(reps);
    for (int k = 0; k < 50; ++k) { *vc = ++r; __sync_synchronize(); wait(r); }
    for (int k = 0; k < reps; ++k) { const double t0 = now(); *vc = ++r; __sync_synchronize(); wait(r); t[k] = now() - t0; }
    CK(cudaStreamSynchronize(st));
    std::sort(t.begin(), t.end());
    printf("%-64s median %6.2f us   p10 %6.2f   p90 %6.2f\n", "persistent server: host doorbell -> zero-copy step -> doorbell", 1e6 * t[reps / 2], 1e6 * t[reps / 10], 1e6 * t[reps * 9 / 10]);
  }
  {
    volatile unsigned* vc = h_cmd; *vc = 0; *vf = 0;
    const int rounds = reps + 50;
    k_pingpong<<<1, 1, 0, st>>>(h_cmd, h_flag, rounds);
    unsigned r = 0;
    std::vector<double> t(reps);
    for (int k = 0; k < 50; ++k) { *vc = ++r; __sync_synchronize(); wait(r); }
    for (int k = 0; k < reps; ++k) { const double t0 = now(); *vc = ++r; __sync_synchronize(); wait(r); t[k] = now() - t0; }
    CK(cudaStreamSynchronize(st));
    std::sort(t.begin(), t.end());
    printf("%-64s median %6.2f us   p10 %6.2f   p90 %6.2f\n", "ping-pong: host word -> 1 GPU thread -> host word", 1e6 * t[reps / 2], 1e6 * t[reps / 10], 1e6 * t[reps * 9 / 10]);
  }
  for (int variant = 0; variant < 2; ++variant) {
    unsigned* h_flags; CK(cudaHostAlloc(&h_flags, 4096, cudaHostAllocMapped)); memset(h_flags, 0, 4096);
    unsigned* d_relay; CK(cudaMalloc(&d_relay, 4)); CK(cudaMemset(d_relay, 0, 4));
    volatile unsigned* vfl = h_flags;
    auto waitall = [&](unsigned s) { for (;;) { bool ok = true; for (int b = 0; b < grid; ++b) ok &= (vfl[b] == s); if (ok) break; __builtin_ia32_pause(); } };
    volatile unsigned* vc = h_cmd; *vc = 0;
    const int rounds = reps + 50;
    const int sp = variant ? spin : 0;
    k_server2<<<grid, 128, 0, st>>>(h_in, h_out, variant ? nin : 0, variant ? nout : 0, sp, h_cmd, d_relay, h_flags, rounds);
    unsigned r = 0;
    std::vector<double> t(reps);
    for (int k = 0; k < 50; ++k) { *vc = ++r; __sync_synchronize(); waitall(r); }
    for (int k = 0; k < reps; ++k) { const double t0 = now(); *vc = ++r; __sync_synchronize(); waitall(r); t[k] = now() - t0; }
    CK(cudaStreamSynchronize(st));
    std::sort(t.begin(), t.end());
    printf("%-64s median %6.2f us   p10 %6.2f   p90 %6.2f\n", variant ? "server v2 (1 poller + relay): zero-copy step" : "server v2 (1 poller + relay): empty step", 1e6 * t[reps / 2], 1e6 * t[reps / 10], 1e6 * t[reps * 9 / 10]);
  }
  printf("done\n");
  return 0;
}
