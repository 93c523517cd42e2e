// FFMA2 with three distinct register-pair operands (as in a register-tiled GEMM), with/without operand reuse.
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int MODE>
__global__ void bench(const float2* in, float* out, long long* cycles, int iters) {
  float2 acc[K], w[K], x[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { acc[k] = in[k]; w[k] = in[K + k + threadIdx.x % 2]; x[k] = in[2 * K + k + threadIdx.x % 3]; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (MODE == 0) acc[k] = __ffma2_rn(w[k], x[k], acc[k]);                 // 3 distinct pairs, no reuse
        else if (MODE == 1) acc[k] = __ffma2_rn(w[k], x[(k / 4) * 4], acc[k]);  // x shared by 4 consecutive FFMA2 (reuse)
        else { acc[k].x = fmaf(w[k].x, x[k].x, acc[k].x); acc[k].y = fmaf(w[k].y, x[k].y, acc[k].y); }  // scalar FFMA pair
      }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) s += acc[k].x + acc[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K, int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc; float2* in;
  cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long)); cudaMalloc(&in, 4096 * sizeof(float2));
  cudaMemset(in, 0, 4096 * sizeof(float2));
  const int iters = 2000;
  bench<K, MODE><<<1, threads>>>(in, out, cyc, iters);
  bench<K, MODE><<<1, threads>>>(in, out, cyc, iters);
  long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double fma_per_warp = (double)iters * 4 * K * 2;      // scalar FMAs per lane
  printf("%-14s K=%2d threads=%4d: %.3f FMA/lane/cycle/SMSP (peak 1.0)\n", name, K, threads,
         fma_per_warp * (threads / 128.0) / c);
  cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<8, 0>("ffma2 3-distinct", threads); run<16, 0>("ffma2 3-distinct", threads); run<24, 0>("ffma2 3-distinct", threads);
    run<16, 1>("ffma2 reuse-x", threads); run<24, 1>("ffma2 reuse-x", threads);
    run<8, 2>("ffma scalar", threads); run<16, 2>("ffma scalar", threads);
  }
  return 0;
}
