// FFMA / FFMA2 latency and throughput micro-benchmark (one CTA, clock64-based).
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool PACKED>
__global__ void bench(float* out, long long* cycles, int iters, float seed) {
  float2 acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = make_float2(seed + k, seed - k);
  const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(seed * 1e-6f, -seed * 1e-6f);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (PACKED) acc[k] = __ffma2_rn(acc[k], a, b);
        else { acc[k].x = fmaf(acc[k].x, a.x, b.x); }
      }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) s += acc[k].x + acc[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K, bool PACKED>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long));
  const int iters = 2000;
  bench<K, PACKED><<<1, threads>>>(out, cyc, iters, 1.5f);
  bench<K, PACKED><<<1, threads>>>(out, cyc, iters, 1.5f);
  long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double n = (double)iters * 8 * K;           // instructions per warp
  const int warps_per_smsp = threads / 128 > 0 ? threads / 128 : 1;
  printf("%-6s K=%2d threads=%4d: %.2f cycles/instr/warp, %.3f instr/cycle/SMSP\n", name, K, threads, c / n,
         n * warps_per_smsp * (threads >= 128 ? 1 : threads / 32.0 / 4) / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {32, 128, 256, 512, 1024}) {
    run<1, true>("FFMA2", threads); run<2, true>("FFMA2", threads); run<4, true>("FFMA2", threads);
    run<8, true>("FFMA2", threads); run<16, true>("FFMA2", threads);
    run<1, false>("FFMA", threads); run<4, false>("FFMA", threads); run<8, false>("FFMA", threads); run<16, false>("FFMA", threads);
  }
  return 0;
}
