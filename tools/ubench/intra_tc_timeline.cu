// Per-step timeline of k_dprnn_intra_tc (SM clock stamps of CTA 0): where a step's ~cycles go.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DITC_TIMELINE -I dpdfnet_b200/csrc \
//        tools/ubench/intra_tc_timeline.cu -o tools/ubench/intra_tc_timeline
#include <cstdio>
#include <cstdlib>
#include <vector>
#ifndef ITC_DUP
#define ITC_DUP 1    // -DITC_DUP=2|4: row-duplicated tiles (128 / D streams per CTA)
#endif
#ifndef ITC_SR
#define ITC_SR 0     // -DITC_SR=1 (with ITC_DUP > 1): split rows, hi | lo halves in the stream's rows, two MMA passes; 2 (ITC_DUP = 4): fragment form
#endif
#include "../../dpdfnet_b200/csrc/k_dprnn_intra_tc.cu"

int main(int argc, char** argv) {
  using namespace dpdf;
  const int B = argc > 1 ? atoi(argv[1]) : 1024, T = 48;
  init_dprnn_intra_tc_kernels();
  float *x, *hcat, *w, *bias;
  long long* tl;
  cudaMalloc(&x, (size_t)B * T * 64 * 4);
  cudaMalloc(&hcat, (size_t)B * T * 128 * 4);
  cudaMalloc(&w, 2 * 4 * 192 * 64 * 2);
  cudaMalloc(&bias, 2 * 4 * 64 * 4);
  cudaMalloc(&tl, T * 12 * 8);
  cudaMemset(x, 0, (size_t)B * T * 64 * 4);
  cudaMemset(w, 0, 2 * 4 * 192 * 64 * 2);
  cudaMemset(bias, 0, 2 * 4 * 64 * 4);
  cudaMemset(tl, 0, T * 12 * 8);
  IntraTcParams p{};
  p.x[0] = p.x[1] = x; p.hcat[0] = p.hcat[1] = hcat; p.Fp[0] = T; p.Fp[1] = 8;
  p.wimg[0] = p.wimg[1] = w; p.wimg_f[0] = p.wimg_f[1] = w; p.bias[0] = p.bias[1] = bias; p.B = B; p.tiles[0] = (B * ITC_DUP + 127) / 128; p.tiles[1] = (B + 127) / 128;
#ifdef ITC_TIMELINE
  p.tl = tl;
#endif
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    k_dprnn_intra_tc<ITC_DUP, 1, ITC_SR><<<2 * (p.tiles[0] + p.tiles[1]), ITC_NT, INTRA_TC_SMEM>>>(p);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("launch %d: %s, %.1f us\n", it, cudaGetErrorString(err), ms * 1e3);
  }
  std::vector<long long> h(T * 12);
  cudaMemcpy(h.data(), tl, T * 12 * 8, cudaMemcpyDeviceToHost);
  printf("step: [gate warp 5] wait_mma ldtm math(+sts/stg) store_x | [issuer] barrier->h_issued x_issued | step_total  (cycles)\n");
  for (int t = 1; t < T - 1; ++t) {
    const long long* a = &h[t * 12];
    const long long* n = &h[(t + 1) * 12];
    printf("%2d: wait %5lld ldtm %5lld math(+copy,stx) %5lld | last slice handed over -> issuer_wake %5lld h_issue %5lld x_issue %5lld | step %5lld\n", t, a[1] - a[0],
           a[2] - a[1], a[3] - a[2], a[5] - a[4], a[6] - a[5], a[7] - a[6], n[0] - a[0]);
  }
  return 0;
}
