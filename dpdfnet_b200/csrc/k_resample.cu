// Batched streaming polyphase resampler (the reference resamples every chunk on the
// host with a stateless librosa call, stream.py:112,163-164 / audio.py:20-27).
//
//     y[m] = sum_j x[j] * h[half + m*down - j*up]        (zero-phase FIR of length 2*half + 1, rational rate up/down)
//
// i.e. the arithmetic of scipy.signal.resample_poly / upfirdn, with the filter designed on the host
// (dpdfnet_b200/resample.py) and kept in device memory.  B streams advance in lock step; each keeps the last `hist`
// input samples, so the concatenation of the chunk outputs (plus flush) equals the one-shot result and there are no
// chunk-boundary artefacts.  HBM-bound: one thread per output sample, taps read from L1/L2, coalesced stores.
#include <cstdio>
#include <cstring>

#include "engine.h"

using namespace dpdf;

struct dpdf_resampler {
  int up = 1, down = 1, half = 0, ntaps = 0, hist = 0, max_streams = 0, max_chunk = 0, device = 0;
  float* h_dev = nullptr;        // [2*half + 1]
  float* hist_dev = nullptr;     // [max_streams][hist]: the last `hist` input samples, oldest first
  float* hist_tmp = nullptr;
  long long n_in = 0;            // input samples consumed so far (all streams)
  long long m_out = 0;           // output samples emitted so far
};

namespace {

struct RsParams {
  const float* in;               // [B][in_stride], n_new valid samples per row
  float* out;                    // [B][out_stride]
  const float* hist;             // [B][H]
  const float* h;
  long long in_stride, out_stride;
  long long n0;                  // input samples consumed before this call
  long long m0;                  // first output index of this call
  int n_new, n_out, H, up, down, half, B;
};

__global__ void __launch_bounds__(256) k_resample(RsParams p) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_out) return;
  const long long m = p.m0 + i;
  const long long c = m * p.down;                                   // centre on the up-sampled grid
  // j ranges over inputs with |c - j*up| <= half
  long long j_lo = (c - p.half + p.up - 1) / p.up;                  // ceil((c - half) / up), c - half may be negative
  if (c - p.half < 0) j_lo = -((p.half - c) / p.up);
  const long long j_hi = (c + p.half) / p.up;
  const long long n_tot = p.n0 + p.n_new;
  const float* xin = p.in + (size_t)b * p.in_stride;
  const float* xh = p.hist + (size_t)b * p.H;
  float acc = 0.f;
  for (long long j = j_lo > 0 ? j_lo : 0; j <= j_hi && j < n_tot; ++j) {
    const float x = j >= p.n0 ? __ldg(xin + (j - p.n0)) : (j >= p.n0 - p.H ? xh[j - (p.n0 - p.H)] : 0.f);
    acc = fmaf(x, __ldg(p.h + p.half + (int)(c - j * p.up)), acc);
  }
  p.out[(size_t)b * p.out_stride + i] = acc;
}

// hist_new[b][k] = sample (n_tot - H + k) of the stream, taken from the new chunk or from the old history
__global__ void k_resample_hist(const float* in, long long in_stride, const float* hist, float* hist_new, int H, int n_new, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, k = idx % H;
  const int src = k + n_new - H;                                     // index into the new chunk (negative: old history)
  float v = 0.f;
  if (src >= 0) v = in[(size_t)b * in_stride + src];
  else if (src + H >= 0) v = hist[(size_t)b * H + src + H];
  hist_new[idx] = v;
}

int rs_fail(int code, const char* msg) { return set_error(code, msg); }

}  // namespace

extern "C" int dpdf_resampler_create(int32_t up, int32_t down, const float* taps_host, int32_t ntaps, int32_t max_streams,
                                     int32_t device, dpdf_resampler** out) {
  if (!out) return rs_fail(DPDF_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (up <= 0 || down <= 0 || !taps_host || ntaps <= 0 || (ntaps & 1) == 0 || max_streams <= 0)
    return rs_fail(DPDF_ERR_INVALID, "resampler: need up, down > 0, an odd number of taps and max_streams > 0");
  if (cudaSetDevice(device) != cudaSuccess) return rs_fail(DPDF_ERR_CUDA, "resampler: cudaSetDevice failed");
  dpdf_resampler* r = new dpdf_resampler();
  r->up = up; r->down = down; r->ntaps = ntaps; r->half = ntaps / 2; r->max_streams = max_streams; r->device = device;
  r->hist = 2 * ((r->half + up - 1) / up) + (down + up - 1) / up + 4;
  if (cudaMalloc(&r->h_dev, ntaps * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&r->hist_dev, (size_t)max_streams * r->hist * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&r->hist_tmp, (size_t)max_streams * r->hist * sizeof(float)) != cudaSuccess) {
    cudaFree(r->h_dev); cudaFree(r->hist_dev); cudaFree(r->hist_tmp);
    delete r;
    return rs_fail(DPDF_ERR_NOMEM, "resampler: cudaMalloc failed");
  }
  cudaMemcpy(r->h_dev, taps_host, ntaps * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemset(r->hist_dev, 0, (size_t)max_streams * r->hist * sizeof(float));
  *out = r;
  return 0;
}

extern "C" int dpdf_resampler_destroy(dpdf_resampler* r) {
  if (!r) return 0;
  cudaSetDevice(r->device);
  cudaFree(r->h_dev); cudaFree(r->hist_dev); cudaFree(r->hist_tmp);
  delete r;
  return 0;
}

extern "C" int dpdf_resampler_reset(dpdf_resampler* r, void* cuda_stream) {
  if (!r) return rs_fail(DPDF_ERR_INVALID, "NULL resampler");
  cudaSetDevice(r->device);
  r->n_in = 0; r->m_out = 0;
  if (cudaMemsetAsync(r->hist_dev, 0, (size_t)r->max_streams * r->hist * sizeof(float), static_cast<cudaStream_t>(cuda_stream)) != cudaSuccess)
    return rs_fail(DPDF_ERR_CUDA, "resampler: memset failed");
  return 0;
}

// Outputs a call with n_new more input samples would emit (flush: everything up to ceil(n_total * up / down)).
extern "C" int64_t dpdf_resampler_pending(const dpdf_resampler* r, int32_t n_new, int32_t flush) {
  if (!r || n_new < 0) return -1;
  const long long n_tot = r->n_in + n_new;
  long long m1;
  if (flush) m1 = (n_tot * r->up + r->down - 1) / r->down;
  else {
    const long long lim = (n_tot - 1) * r->up - r->half;           // last centre whose whole support has arrived
    m1 = lim < 0 ? 0 : lim / r->down + 1;
    const long long cap = (n_tot * r->up + r->down - 1) / r->down;
    if (m1 > cap) m1 = cap;
  }
  return m1 > r->m_out ? m1 - r->m_out : 0;
}

extern "C" int dpdf_resampler_process(dpdf_resampler* r, const float* in, int64_t in_stride, int32_t n_new, float* out,
                                      int64_t out_stride, int32_t B, int32_t flush, int64_t* n_out, void* cuda_stream) {
  if (!r || !n_out || (n_new > 0 && !in)) return rs_fail(DPDF_ERR_INVALID, "resampler: NULL argument");
  if (B <= 0 || B > r->max_streams || n_new < 0) return rs_fail(DPDF_ERR_INVALID, "resampler: bad batch or chunk size");
  cudaSetDevice(r->device);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const long long cnt = dpdf_resampler_pending(r, n_new, flush);
  *n_out = cnt;
  if (cnt > 0) {
    if (!out || out_stride < cnt) return rs_fail(DPDF_ERR_INVALID, "resampler: output buffer too small");
    RsParams p{in, out, r->hist_dev, r->h_dev, in_stride, out_stride, r->n_in, r->m_out, n_new, (int)cnt, r->hist, r->up, r->down, r->half, B};
    dim3 grid((unsigned)((cnt + 255) / 256), B);
    k_resample<<<grid, 256, 0, st>>>(p);
  }
  if (n_new > 0) {
    k_resample_hist<<<(B * r->hist + 255) / 256, 256, 0, st>>>(in, in_stride, r->hist_dev, r->hist_tmp, r->hist, n_new, B);
    float* t = r->hist_dev; r->hist_dev = r->hist_tmp; r->hist_tmp = t;
  }
  if (cudaGetLastError() != cudaSuccess) return rs_fail(DPDF_ERR_CUDA, "resampler: kernel launch failed");
  r->n_in += n_new;
  r->m_out += cnt;
  return 0;
}
