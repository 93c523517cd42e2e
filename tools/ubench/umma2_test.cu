// Standalone validation of the cta_group::2 (CTA pair) tcgen05 machinery before it goes into the engine:
//   D[256][N] = A[256][K] * B[N][K]^T, kind::f16, FP32 accumulate, K-major SWIZZLE_NONE operand images.
// A cluster of two CTAs: CTA r stages rows [128 r, 128 r + 128) of A and rows [N/2 r, N/2 r + N/2) of B in its own shared
// memory (same offsets in both CTAs), both allocate tensor memory collectively, the peer tells the leader that its
// operands are staged with a remote mbarrier arrive, the leader issues the M = 256 MMAs and commits with a multicast to
// both CTAs' barriers; each CTA reads its 128 accumulator rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench/umma2_test.cu -o tools/ubench/umma2_test
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// byte offset of element (r, k) in an FP16 [rows][K] K-major SWIZZLE_NONE image: 8 x 16 B core matrices, LBO 128 B, SBO K/8 * 128 B
__host__ __device__ inline int img_off(int r, int k, int K) { return (r >> 3) * (K / 8) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

template <int N, int K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) umma2_test(const __half* A, const __half* B, float* D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* As = smem;                        // [128][K] halves
  unsigned char* Bs = smem + 128 * K * 2;          // [N/2][K] halves
  __shared__ __align__(8) uint64_t bar_ready, bar_done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;

  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(As + img_off(r, k, K)) = A[((size_t)pair * 256 + rank * 128 + r) * K + k];
  }
  for (int i = tid; i < N / 2 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(Bs + img_off(r, k, K)) = B[(size_t)(rank * (N / 2) + r) * K + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&bar_ready)));
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();                               // both CTAs' barriers exist before anyone signals across
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    if (rank == 1) {                                // peer: operands staged -> arrive on the leader's barrier
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(&bar_ready)));
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    } else {
      asm volatile(
          "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra W_%=;\nDONE_%=:\n}\n" ::"r"(
              smem_u32(&bar_ready))
          : "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t desc0 = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((K / 8 * 128) >> 4) << 32) | ((uint64_t)1 << 46);
      const uint64_t da = desc0 | (smem_u32(As) >> 4), db = desc0 | (smem_u32(Bs) >> 4);
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t acc = ks > 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem),
            "l"(da + ks * 16), "l"(db + ks * 16), "r"(idesc), "r"(acc), "r"(0u)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)),
                   "h"((uint16_t)3)
                   : "memory");
    }
  }
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra W_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(&bar_done))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[((size_t)pair * 256 + rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                               // both CTAs have drained their accumulators
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N < 32 ? 32 : N));
}

template <int N, int K>
int run(int pairs) {
  const int M = 256 * pairs;
  std::vector<__half> A((size_t)M * K), B((size_t)N * K);
  std::vector<float> D((size_t)M * N);
  srand(1);
  for (auto& v : A) v = __float2half((rand() / (float)RAND_MAX - 0.5f) * 2.f);
  for (auto& v : B) v = __float2half((rand() / (float)RAND_MAX - 0.5f) * 2.f);
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  const size_t smem = (size_t)(128 * K + N / 2 * K) * 2;
  cudaFuncSetAttribute(umma2_test<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma2_test<N, K><<<2 * pairs, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d K=%d pairs=%d CUDA error: %s\n", N, K, pairs, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)__half2float(A[(size_t)m * K + k]) * __half2float(B[(size_t)n * K + k]);
      const double d = fabs(D[(size_t)m * N + n] - ref);
      if (!(d <= err)) err = d;                    // NaN-propagating max
    }
  printf("cta_group::2  M=256 N=%3d K=%3d pairs=%d: max|D - fp64(A*B)| = %.3e\n", N, K, pairs, err);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return !(err < 1e-3);
}

int main() {
  int bad = 0;
  bad |= run<64, 64>(1);
  bad |= run<64, 64>(37);
  bad |= run<128, 64>(3);
  bad |= run<64, 128>(3);
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}
