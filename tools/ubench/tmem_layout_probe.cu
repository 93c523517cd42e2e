// Which (TMEM lane, column) does each thread of a warp see with the 16-lane fragment shapes of tcgen05.ld / tcgen05.st?
// Every warp w of a 128-thread CTA fills its lane quadrant with 32x32b stores (value = lane * 256 + column), reads it back
// with .16x128b.x1 / .16x64b.x1 at lane offsets 0 and 16, and the host prints the tables; then the reverse: a .16x128b.x1
// store of (thread * 4 + register) read back with 32x32b loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dpdfnet_b200/csrc tools/ubench/tmem_layout_probe.cu -o tools/ubench/tmem_layout_probe
#include <cstdio>
#include <cstdint>
#include "../../dpdfnet_b200/csrc/tc_common.cuh"

using namespace dpdf::tc;

__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc<32>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16);
  // fill: lane i, columns 0..7
  for (int c = 0; c < 8; c += 4)
    tmem_st4(base + c, (warp * 32 + lane) * 256 + c, (warp * 32 + lane) * 256 + c + 1, (warp * 32 + lane) * 256 + c + 2, (warp * 32 + lane) * 256 + c + 3);
  tmem_st_wait();
  __syncwarp();
  uint32_t r0, r1;
  for (int off = 0; off < 32; off += 16) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(base + ((uint32_t)off << 16)));
    tmem_ld_wait();
    out[((0 * 2 + off / 16) * 128 + tid) * 2] = r0;
    out[((0 * 2 + off / 16) * 128 + tid) * 2 + 1] = r1;
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(base + ((uint32_t)off << 16)));
    tmem_ld_wait();
    out[((1 * 2 + off / 16) * 128 + tid) * 2] = r0;
    out[((1 * 2 + off / 16) * 128 + tid) * 2 + 1] = 0;
  }
  __syncwarp();
  // reverse: 16x128b.x1 store at lane offset 0, columns 8..11, then 32x32b load
  asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1,%2};" ::"r"(base + 8), "r"(0x10000u + lane * 4), "r"(0x10000u + lane * 4 + 1) : "memory");
  tmem_st_wait();
  __syncwarp();
  uint32_t g[4];
  tmem_ld4_nowait(base + 8, g);
  tmem_ld_wait();
  for (int i = 0; i < 4; ++i) out[4 * 256 + tid * 4 + i] = g[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, (4 * 256 + 512) * 4);
  cudaMemset(d, 0xff, (4 * 256 + 512) * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[4 * 256 + 512];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[2] = {"16x128b.x1", "16x64b.x1"};
  for (int s = 0; s < 2; ++s)
    for (int o = 0; o < 2; ++o) {
      printf("== ld %s lane offset %d, warp 1 (quadrant lanes 32..63): thread -> (lane, col) per register\n", names[s], o * 16);
      for (int t = 32; t < 64; ++t) {
        const uint32_t a = h[((s * 2 + o) * 128 + t) * 2], b = h[((s * 2 + o) * 128 + t) * 2 + 1];
        printf("  t%2d: r0=(%3u,%u)", t - 32, a >> 8, a & 255);
        if (s == 0) printf(" r1=(%3u,%u)", b >> 8, b & 255);
        if ((t & 3) == 3) printf("\n");
      }
    }
  printf("== st 16x128b.x1 (value = 0x10000 + thread*4 + reg) read back 32x32b.x4 at cols 8..11, warp 1: lane -> values (thread,reg)\n");
  for (int t = 32; t < 64; ++t) {
    printf("  lane %2d:", t - 32);
    for (int i = 0; i < 4; ++i) {
      const uint32_t v = h[4 * 256 + t * 4 + i];
      if (v >> 16 == 1) printf(" (t%2u,r%u)", (v & 0xffff) >> 2, v & 3);
      else printf(" (%08x)", v);
    }
    printf("\n");
  }
  return 0;
}
