// Does an FFMA2 block the issue port for both of its pipe cycles?  Mix independent integer ops in.
#include <cstdio>
#include <cuda_runtime.h>

template <int NALU>
__global__ void bench(const float2* in, float* out, long long* cycles, int iters) {
  constexpr int K = 16;
  float2 acc[K], w[K];
  float2 x = in[100 + threadIdx.x % 3];
  int ia[8];
#pragma unroll
  for (int k = 0; k < K; ++k) { acc[k] = in[k]; w[k] = in[K + k + threadIdx.x % 2]; }
#pragma unroll
  for (int k = 0; k < 8; ++k) ia[k] = threadIdx.x + k;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        acc[k] = __ffma2_rn(w[k], x, acc[k]);
#pragma unroll
        for (int a = 0; a < NALU; ++a) ia[(k + a) & 7] = (ia[(k + a) & 7] ^ (i + k)) + 0x9e37;   // LOP3 + IADD on the ALU pipe
      }
  }
  long long t1 = clock64();
  float s = 0.f;
  int is = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) s += acc[k].x + acc[k].y;
#pragma unroll
  for (int k = 0; k < 8; ++k) is += ia[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + is;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NALU>
void run(int threads) {
  float* out; long long* cyc; float2* in;
  cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long)); cudaMalloc(&in, 4096 * sizeof(float2));
  cudaMemset(in, 0, 4096 * sizeof(float2));
  const int iters = 2000;
  bench<NALU><<<1, threads>>>(in, out, cyc, iters);
  bench<NALU><<<1, threads>>>(in, out, cyc, iters);
  long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double nf = (double)iters * 4 * 16;
  printf("alu_ops_per_ffma2=%d threads=%4d: %.2f cycles per FFMA2 per SMSP\n", NALU, threads, c / (nf * (threads / 128.0)));
  cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
  for (int threads : {128, 256, 512}) { run<0>(threads); run<1>(threads); run<2>(threads); run<3>(threads); }
  return 0;
}
