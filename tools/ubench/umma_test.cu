// Standalone validation of the tcgen05 / TMEM machinery used by the engine:
// D[128][N] = A[128][K] * B[N][K]^T with kind::tf32, operands in the K-major SWIZZLE_NONE ("interleave")
// shared-memory layout, accumulator in TMEM, single pass and error-compensated 3xTF32.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int M = 128;
__host__ __device__ constexpr int tmem_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes (32 floats, 128 B contiguous);
// core matrices adjacent in K are LBO = 128 B apart, 8-row groups are SBO = (K/4)*128 B apart.
__host__ __device__ inline int core_off(int r, int k, int K) { return (r >> 3) * (K / 4) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  return d;                        // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int N, int K>
__global__ void __launch_bounds__(128) umma_test(const float* A, const float* B, float* D, int three_pass) {
  extern __shared__ __align__(128) float smem[];
  float* As = smem;                 // [128*K]
  float* Al = As + M * K;
  float* Bs = Al + M * K;           // [N*K]
  float* Bl = Bs + N * K;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    const float v = A[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    As[core_off(r, k, K)] = hi;
    Al[core_off(r, k, K)] = __uint_as_float(__float_as_uint(v - hi) & 0xffffe000u);
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const float v = B[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    Bs[core_off(r, k, K)] = hi;
    Bl[core_off(r, k, K)] = __uint_as_float(__float_as_uint(v - hi) & 0xffffe000u);
  }
  if (tid == 0) asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(tmem_cols(N)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t sbo = (K / 4) * 128;
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t da = make_desc(smem_u32(As) + ks * 256, 128, sbo), dal = make_desc(smem_u32(Al) + ks * 256, 128, sbo);
      const uint64_t db = make_desc(smem_u32(Bs) + ks * 256, 128, sbo), dbl = make_desc(smem_u32(Bl) + ks * 256, 128, sbo);
      umma_tf32(tmem, da, db, idesc, acc);
      acc = 1;
      if (three_pass) {
        umma_tf32(tmem, dal, db, idesc, 1);
        umma_tf32(tmem, da, dbl, idesc, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra W_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(&bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w owns TMEM lanes 32w..32w+31 = rows; 32x32b.x8 -> 8 consecutive columns per thread
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(tmem_cols(N)));
}

template <int N, int K>
int run() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
  cudaFuncSetAttribute(umma_test<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int bad = 0;
  for (int pass3 = 0; pass3 < 2; ++pass3) {
    cudaMemset(dD, 0xff, D.size() * 4);
    umma_test<N, K><<<1, 128, smem>>>(dA, dB, dD, pass3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d pass3=%d CUDA error: %s\n", N, K, pass3, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double e_full = 0, e_tf = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double full = 0, tf = 0;
        for (int k = 0; k < K; ++k) {
          const float a = A[m * K + k], b = B[n * K + k];
          uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); ua &= 0xffffe000u; ub &= 0xffffe000u;
          float ta, tb; memcpy(&ta, &ua, 4); memcpy(&tb, &ub, 4);
          full += (double)a * b; tf += (double)ta * tb;
        }
        e_full = fmax(e_full, fabs(D[m * N + n] - full));
        e_tf = fmax(e_tf, fabs(D[m * N + n] - tf));
      }
    printf("N=%3d K=%3d %s: max|D - fp64(A*B)| = %.3e   max|D - fp64(trunc_tf32 A * trunc_tf32 B)| = %.3e\n", N, K,
           pass3 ? "3xTF32" : "1xTF32", e_full, e_tf);
    if (pass3 ? e_full > 1e-4 : e_tf > 1e-4) bad = 1;
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad;
}

int main() {
  int bad = 0;
  bad |= run<64, 64>();
  bad |= run<64, 128>();
  bad |= run<192, 64>();
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}
