// Phase timeline of k_sepconv_tc (SM clock stamps of CTA 0) for a plain depthwise + pointwise problem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSCT_TIMELINE tools/ubench/sepconv_tc_timeline.cu -o tools/ubench/sepconv_tc_timeline
#include <cstdio>
#include <cstdlib>
#include "../../dpdfnet_b200/csrc/k_conv_tc.cu"

int main(int argc, char** argv) {
  using namespace dpdf;
  const int B = argc > 1 ? atoi(argv[1]) : 1024, F = 48;
  init_conv_tc_kernels();
  float *in, *out, *w;
  long long* tl;
  cudaMalloc(&in, (size_t)B * F * 64 * 4); cudaMalloc(&out, (size_t)B * F * 64 * 4); cudaMalloc(&w, 1 << 20); cudaMalloc(&tl, 64);
  cudaMemset(in, 0, (size_t)B * F * 64 * 4); cudaMemset(w, 0, 1 << 20);
  SepTcParams p{};
  p.nprob = 1; p.B = B; p.tl = tl;
  SepProblem& q = p.prob[0];
  q.mode = 0; q.in1 = in; q.in2 = nullptr; q.dw = w; q.pw = w; q.tc_pw = w + 4096; q.bias = w; q.out = out;
  q.Fin = F; q.Fout = F; q.stride = 1; q.up = 1; q.tile0 = 0;
  const int tiles = (B * F + 127) / 128;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    k_sepconv_tc<512><<<tiles, 512, SCT_SMEM>>>(p);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("B=%d tiles=%d launch %d: %s, %.1f us\n", B, tiles, it, cudaGetErrorString(err), ms * 1e3);
  }
  long long h[8];
  cudaMemcpy(h, tl, 64, cudaMemcpyDeviceToHost);
  printf("cycles: setup (barriers, TMEM alloc, bias) %lld | prologue loads + dw conv + split %lld | fence+sync %lld | mma + commit wait %lld | epilogue %lld | store %lld | total %lld\n",
         h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[6] - h[0]);
  return 0;
}
