// Phase timeline of k_dprnn_post_tc (SM clock stamps of CTA 0) and launch time against the tile count.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPT_TIMELINE tools/ubench/post_tc_timeline.cu -o tools/ubench/post_tc_timeline
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dpdfnet_b200/csrc/k_dprnn_tc.cu"
namespace dpdf { int intra_tc_dup(const Engine&, int) { return 1; } int intra_tc_dup_erb(const Engine&, int) { return 1; } }   // the launcher (unused here) refers to it

int main(int argc, char** argv) {
  using namespace dpdf;
  const int B = argc > 1 ? atoi(argv[1]) : 1024, Fp = 48;
  init_dprnn_tc_kernels();
  const size_t rows = (size_t)B * Fp;
  float *hcat, *x, *out, *hstate, *w, *small;
  IoDesc* io;
  long long* tl;
  cudaMalloc(&hcat, rows * 128 * 4); cudaMalloc(&x, rows * 64 * 4); cudaMalloc(&out, rows * 64 * 4);
  cudaMalloc(&hstate, rows * 64 * 4); cudaMalloc(&w, 9 * 16384); cudaMalloc(&small, 1024 * 4);
  cudaMalloc(&io, sizeof(IoDesc)); cudaMalloc(&tl, 16 * 8);
  cudaMemset(hcat, 0, rows * 128 * 4); cudaMemset(x, 0, rows * 64 * 4); cudaMemset(hstate, 0, rows * 64 * 4);
  cudaMemset(w, 0, 9 * 16384); cudaMemset(small, 0, 4096); cudaMemset(io, 0, sizeof(IoDesc)); cudaMemset(tl, 0, 128);
  PostTcParams p{};
  p.io = io; p.B = B; p.tl = tl;
  PostTcBranch& b = p.br[0];
  b.hcat = hcat; b.xin = x; b.xout = out; b.hstate = hstate; b.per_slot = (long long)Fp * 64; b.Fp = Fp;
  b.tc_fc_w = w; b.tc_gates = w + 2 * 4096; b.tc_fc2_w = w + 8 * 4096;
  b.fc_b = small; b.ln_g = small + 64; b.ln_b = small + 128; b.bias = small + 192; b.fc2_b = small + 448; b.ln2_g = small + 512; b.ln2_b = small + 576;
  p.br[1] = b;
  p.tiles0 = (int)((rows + 127) / 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
#ifdef PT_RES      // persistent form with resident weights: one CTA per SM, stamps of the third tile of the last CTA
    {
      static int* ctr = nullptr;
      if (!ctr) { cudaMalloc(&ctr, 16); cudaMemset(ctr, 0, 16); }
      p.ctr = ctr; p.ctas_erb = 0; p.tiles1 = 0;
      const int G = p.tiles0 < 148 ? p.tiles0 : 148;
      cudaFuncSetAttribute(k_dprnn_post_res, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POST_RES_SMEM);
      k_dprnn_post_res<<<G, TC_NT, POST_RES_SMEM>>>(p);
    }
#elif defined(PT_PAIR)     // CTA pairs: clusters of two, even tile count
    {
      p.tiles0 = (p.tiles0 + 1) & ~1;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(p.tiles0); cfg.blockDim = dim3(TC_NT); cfg.dynamicSmemBytes = POST_TC_SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_dprnn_post_tc<1>, p);
    }
#else
    k_dprnn_post_tc<0><<<p.tiles0, TC_NT, POST_TC_SMEM>>>(p);
#endif
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("B=%d tiles=%d launch %d: %s, %.1f us\n", B, p.tiles0, it, cudaGetErrorString(err), ms * 1e3);
  }
  long long h[16];
  cudaMemcpy(h, tl, 128, cudaMemcpyDeviceToHost);
  printf("cycles: hcat staged+sync %lld | phase1 mma %lld | epilogue1 %lld | phase2 mma %lld | epilogue2 %lld | phase3 mma %lld | epilogue3+store %lld | total %lld\n",
         h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[7] - h[0]);
  return 0;
}
