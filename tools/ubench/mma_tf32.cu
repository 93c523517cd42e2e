// Legacy mma.sync TF32 / BF16 throughput on sm_100a (one CTA per SM, clock64).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int K, bool BF16>
__global__ void bench(const unsigned* in, float* out, long long* cycles, int iters) {
  float c[K][4];
  unsigned a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = in[threadIdx.x % 7 + i];
  b[0] = in[threadIdx.x % 5]; b[1] = in[threadIdx.x % 3 + 9];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[k][i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (BF16) mma_bf16(c[k], a, b); else mma_tf32(c[k], a, b);
      }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K, bool BF16>
void run(int threads) {
  float* out; long long* cyc; unsigned* in;
  cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long)); cudaMalloc(&in, 4096 * sizeof(unsigned));
  cudaMemset(in, 0, 4096 * sizeof(unsigned));
  const int iters = 2000;
  bench<K, BF16><<<1, threads>>>(in, out, cyc, iters);
  bench<K, BF16><<<1, threads>>>(in, out, cyc, iters);
  long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double mmas = (double)iters * 4 * K * (threads / 32);
  const double macs = mmas * 16 * 8 * (BF16 ? 16 : 8);
  printf("%s K=%d threads=%4d: %.2f cycles/mma/warp, %.0f MAC/cycle/SM\n", BF16 ? "bf16 m16n8k16" : "tf32 m16n8k8 ", K, threads,
         c / ((double)iters * 4 * K), macs / c);
  cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
  for (int threads : {32, 128, 256, 512}) {
    run<1, false>(threads); run<4, false>(threads); run<8, false>(threads);
    run<4, true>(threads); run<8, true>(threads);
  }
  return 0;
}
