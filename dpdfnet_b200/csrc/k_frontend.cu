// Analysis (a1-a4) and synthesis (a11-a13) kernels, plus slot reset / priming.
//
// analysis : framed windowed real DFT (x wnorm) fused with ERB-band energy / per-bin log-magnitude,
//            the two running normalisers and the pushes into the mask / feature rings.
//            reference: stream.py:119-126, onnx_model/dpdfnet.py:814-852, layers.py:485-572 (16 k),
//            layers.py:621-730 + dpdfnet_48khz_hr.py:903 (48 k)
// synthesis: ERB / per-bin mask on the 2-frame-delayed spectrum, 5-tap complex deep filter with
//            2-frame-delayed coefficients, inverse real DFT x window, overlap-add.
//            reference: layers.py:414-445, onnx_model/multiframe.py:140-154,200-232, stream.py:138-156
#include <algorithm>

#include "engine.h"

namespace dpdf {

constexpr int ABT = 8;   // streams per CTA in the analysis kernel
constexpr int SBT = 4;   // streams per CTA in the synthesis kernel

struct AnaParams {
  const IoDesc* io;
  Dims d;
  State st;
  const float* dft_fwd;      // [win][F][2]
  const float* band_inv_w;   // [32]
  const int* band_start;     // [33]
  int B;
};

// Single-group form (throughput-bound grids): thread = bin over all win samples.
__global__ void __launch_bounds__(512) k_analysis_one(AnaParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  const int win = p.d.win, hop = p.d.hop, F = p.d.F;
  float2* xs2 = reinterpret_cast<float2*>(smem);                 // [win][ABT] {x,x}
  float* pws = smem + 2 * win * ABT;                              // [ABT][F]
  const IoDesc* io = p.io;
  const int b0 = blockIdx.x * ABT;
  const int nb = min(ABT, p.B - b0);
  const int tid = threadIdx.x, NT = blockDim.x;
  const bool pcm_mode = io->mode == 0;

  __shared__ int s_slot[ABT], s_flag[ABT], s_pos[ABT];
  if (tid < ABT) {
    int b = b0 + tid;
    s_slot[tid] = b < p.B ? io_slot(io, b) : 0;
    s_flag[tid] = b < p.B ? io_flags(io, b) : 0;
    s_pos[tid] = b < p.B ? p.st.pos[s_slot[tid]] : 0;
  }
  __syncthreads();

  float2 X[ABT];
#pragma unroll
  for (int bb = 0; bb < ABT; ++bb) X[bb] = make_float2(0.f, 0.f);

  if (pcm_mode) {
    const long long toff = (long long)io->t_in * hop;
    for (int i = tid; i < ABT * win; i += NT) {
      int bb = i / win, n = i % win;
      float v = 0.f;
      if (bb < nb) {
        v = n < hop ? p.st.in_hist[(size_t)s_slot[bb] * hop + n]
                    : __ldg(io->in + (size_t)(b0 + bb) * io->in_stride + toff + (n - hop));
      }
      xs2[n * ABT + bb] = make_float2(v, v);
    }
    __syncthreads();
    for (int i = tid; i < nb * hop; i += NT) {      // history <- this hop (after all reads above)
      int bb = i / hop, n = i % hop;
      p.st.in_hist[(size_t)s_slot[bb] * hop + n] = xs2[(n + hop) * ABT + bb].x;
    }
    if (tid < F) {
      // the (cos, sin) basis streams from L2 (412 KB at 16 kHz, larger than L1): keep 16 loads in flight
      const float2* basis = reinterpret_cast<const float2*>(p.dft_fwd) + tid;
      for (int n0 = 0; n0 < win; n0 += 16) {
        float2 cs[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) cs[u] = __ldg(basis + (size_t)(n0 + u) * F);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float4* xr = reinterpret_cast<const float4*>(xs2 + (n0 + u) * ABT);
#pragma unroll
          for (int q = 0; q < ABT / 2; ++q) {
            float4 xx = xr[q];
            X[2 * q] = ffma2(cs[u], lo2(xx), X[2 * q]);
            X[2 * q + 1] = ffma2(cs[u], hi2(xx), X[2 * q + 1]);
          }
        }
      }
    }
  } else if (tid < F) {
    for (int bb = 0; bb < nb; ++bb) {
      float2 v = __ldg(reinterpret_cast<const float2*>(io->in) + (size_t)(b0 + bb) * F + tid);
      X[bb] = make_float2(v.x * p.d.wnorm, v.y * p.d.wnorm);
    }
  }

  const float a = 0.98f, one_m_a = 0.02f;     // float32(0.98), float32(1 - 0.98)
  if (tid < F) {
    const int k = tid;
#pragma unroll
    for (int bb = 0; bb < ABT; ++bb) {
      if (bb >= nb) break;
      const int slot = s_slot[bb], fl = s_flag[bb], pos = s_pos[bb];
      if (fl & DPDF_FLAG_ZERO_SPEC_) X[bb] = make_float2(0.f, 0.f);
      const float re = X[bb].x, im = X[bb].y;
      reinterpret_cast<float2*>(p.st.mask_ring)[((size_t)slot * 3 + pos % 3) * F + k] = X[bb];
      const float pw = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
      const bool kill = (fl & (DPDF_FLAG_WARMUP_ | DPDF_FLAG_ZERO_FEAT_)) != 0;
      if (p.d.hr48) {
        const float feat = 10.0f * log10f(sqrtf(pw) + 1e-10f);
        float* mup = p.st.mu + (size_t)slot * p.d.fe_feat + k;
        const float mu = __fadd_rn(__fmul_rn(a, *mup), __fmul_rn(one_m_a, feat));
        *mup = mu;
        p.st.erb_ring[((size_t)slot * 3 + pos % 3) * p.d.fe_feat + k] = kill ? 0.f : (feat - mu) / 40.0f;
      } else {
        pws[bb * F + k] = pw;
      }
      if (k < NDF) {
        const float mag = sqrtf(pw);
        float* sp = p.st.s + (size_t)slot * NDF + k;
        const float s = __fadd_rn(__fmul_rn(a, *sp), __fmul_rn(one_m_a, mag));
        *sp = s;
        const float den = sqrtf(s + 1e-12f);
        float* ring = p.st.df_ring + ((size_t)slot * 3 + pos % 3) * 2 * NDF;
        ring[k] = kill ? 0.f : re / den;
        ring[NDF + k] = kill ? 0.f : im / den;
      }
    }
  }
  if (!p.d.hr48) {
    __syncthreads();
    for (int i = tid; i < nb * 32; i += NT) {
      const int bb = i >> 5, band = i & 31;
      const int slot = s_slot[bb], fl = s_flag[bb], pos = s_pos[bb];
      const int k0 = p.band_start[band], k1 = p.band_start[band + 1];
      const float iw = p.band_inv_w[band];
      float acc = 0.f;
      for (int k = k0; k < k1; ++k) acc = __fadd_rn(acc, __fmul_rn(pws[bb * F + k], iw));
      const float feat = 10.0f * log10f(acc + 1e-10f);
      float* mup = p.st.mu + (size_t)slot * 32 + band;
      const float mu = __fadd_rn(__fmul_rn(a, *mup), __fmul_rn(one_m_a, feat));
      *mup = mu;
      const bool kill = (fl & (DPDF_FLAG_WARMUP_ | DPDF_FLAG_ZERO_FEAT_)) != 0;
      p.st.erb_ring[((size_t)slot * 3 + pos % 3) * 32 + band] = kill ? 0.f : (feat - mu) / 40.0f;
    }
  }
}

// The DFT sum over the win samples is split over `ksplit` thread groups (group g = n range [g win / ksplit, ...)):
// the per-thread chain of win dependent load + FFMA2 rounds is what bounds this kernel at every batch size, and
// the partial spectra are added through shared memory in a fixed order (g = 0 first).
template <int MAXT>      // 512: one group (full register budget for the 16-deep basis prefetch), 1024: split sum
__global__ void __launch_bounds__(MAXT) k_analysis(AnaParams p, int ksplit) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  const int win = p.d.win, hop = p.d.hop, F = p.d.F;
  float2* xs2 = reinterpret_cast<float2*>(smem);                 // [win][ABT] {x,x}
  float* pws = smem + 2 * win * ABT;                              // [ABT][F]
  float2* part = reinterpret_cast<float2*>(pws + ABT * F);        // [ksplit - 1][ABT][F] partial spectra of groups 1..
  const IoDesc* io = p.io;
  const int b0 = blockIdx.x * ABT;
  const int nb = min(ABT, p.B - b0);
  const int NT = blockDim.x, NTg = NT / ksplit;
  const int grp = threadIdx.x / NTg;
  const int tid = threadIdx.x;
  const bool pcm_mode = io->mode == 0;

  __shared__ int s_slot[ABT], s_flag[ABT], s_pos[ABT];
  if (tid < ABT) {
    int b = b0 + tid;
    s_slot[tid] = b < p.B ? io_slot(io, b) : 0;
    s_flag[tid] = b < p.B ? io_flags(io, b) : 0;
    s_pos[tid] = b < p.B ? p.st.pos[s_slot[tid]] : 0;
  }
  __syncthreads();

  float2 X[ABT];
#pragma unroll
  for (int bb = 0; bb < ABT; ++bb) X[bb] = make_float2(0.f, 0.f);

  if (pcm_mode) {
    const long long toff = (long long)io->t_in * hop;
    for (int i = tid; i < ABT * win; i += NT) {
      int bb = i / win, n = i % win;
      float v = 0.f;
      if (bb < nb) {
        v = n < hop ? p.st.in_hist[(size_t)s_slot[bb] * hop + n]
                    : __ldg(io->in + (size_t)(b0 + bb) * io->in_stride + toff + (n - hop));
      }
      xs2[n * ABT + bb] = make_float2(v, v);
    }
    __syncthreads();
    for (int i = tid; i < nb * hop; i += NT) {      // history <- this hop (after all reads above)
      int bb = i / hop, n = i % hop;
      p.st.in_hist[(size_t)s_slot[bb] * hop + n] = xs2[(n + hop) * ABT + bb].x;
    }
    const int kb = tid - grp * NTg;                               // bin of this thread inside its group
    if (kb < F) {
      // the (cos, sin) basis streams from L2 (412 KB at 16 kHz, larger than L1): keep 16 loads in flight
      const float2* basis = reinterpret_cast<const float2*>(p.dft_fwd) + kb;
      const int nper = win / ksplit;
      for (int n0 = grp * nper; n0 < (grp + 1) * nper; n0 += 16) {
        float2 cs[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) cs[u] = __ldg(basis + (size_t)(n0 + u) * F);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float4* xr = reinterpret_cast<const float4*>(xs2 + (n0 + u) * ABT);
#pragma unroll
          for (int q = 0; q < ABT / 2; ++q) {
            float4 xx = xr[q];
            X[2 * q] = ffma2(cs[u], lo2(xx), X[2 * q]);
            X[2 * q + 1] = ffma2(cs[u], hi2(xx), X[2 * q + 1]);
          }
        }
      }
      if (grp > 0) {
#pragma unroll
        for (int bb = 0; bb < ABT; ++bb) part[((grp - 1) * ABT + bb) * F + kb] = X[bb];
      }
    }
    if (ksplit > 1) {
      __syncthreads();
      if (tid < F) {
        for (int g = 1; g < ksplit; ++g)
#pragma unroll
          for (int bb = 0; bb < ABT; ++bb) {
            const float2 v = part[((g - 1) * ABT + bb) * F + tid];
            X[bb].x += v.x; X[bb].y += v.y;
          }
      }
    }
  } else if (tid < F) {
    for (int bb = 0; bb < nb; ++bb) {
      float2 v = __ldg(reinterpret_cast<const float2*>(io->in) + (size_t)(b0 + bb) * F + tid);
      X[bb] = make_float2(v.x * p.d.wnorm, v.y * p.d.wnorm);
    }
  }

  const float a = 0.98f, one_m_a = 0.02f;     // float32(0.98), float32(1 - 0.98)
  if (tid < F) {
    const int k = tid;
#pragma unroll
    for (int bb = 0; bb < ABT; ++bb) {
      if (bb >= nb) break;
      const int slot = s_slot[bb], fl = s_flag[bb], pos = s_pos[bb];
      if (fl & DPDF_FLAG_ZERO_SPEC_) X[bb] = make_float2(0.f, 0.f);
      const float re = X[bb].x, im = X[bb].y;
      reinterpret_cast<float2*>(p.st.mask_ring)[((size_t)slot * 3 + pos % 3) * F + k] = X[bb];
      const float pw = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
      const bool kill = (fl & (DPDF_FLAG_WARMUP_ | DPDF_FLAG_ZERO_FEAT_)) != 0;
      if (p.d.hr48) {
        const float feat = 10.0f * log10f(sqrtf(pw) + 1e-10f);
        float* mup = p.st.mu + (size_t)slot * p.d.fe_feat + k;
        const float mu = __fadd_rn(__fmul_rn(a, *mup), __fmul_rn(one_m_a, feat));
        *mup = mu;
        p.st.erb_ring[((size_t)slot * 3 + pos % 3) * p.d.fe_feat + k] = kill ? 0.f : (feat - mu) / 40.0f;
      } else {
        pws[bb * F + k] = pw;
      }
      if (k < NDF) {
        const float mag = sqrtf(pw);
        float* sp = p.st.s + (size_t)slot * NDF + k;
        const float s = __fadd_rn(__fmul_rn(a, *sp), __fmul_rn(one_m_a, mag));
        *sp = s;
        const float den = sqrtf(s + 1e-12f);
        float* ring = p.st.df_ring + ((size_t)slot * 3 + pos % 3) * 2 * NDF;
        ring[k] = kill ? 0.f : re / den;
        ring[NDF + k] = kill ? 0.f : im / den;
      }
    }
  }
  if (!p.d.hr48) {
    __syncthreads();
    for (int i = tid; i < nb * 32; i += NT) {
      const int bb = i >> 5, band = i & 31;
      const int slot = s_slot[bb], fl = s_flag[bb], pos = s_pos[bb];
      const int k0 = p.band_start[band], k1 = p.band_start[band + 1];
      const float iw = p.band_inv_w[band];
      float acc = 0.f;
      for (int k = k0; k < k1; ++k) acc = __fadd_rn(acc, __fmul_rn(pws[bb * F + k], iw));
      const float feat = 10.0f * log10f(acc + 1e-10f);
      float* mup = p.st.mu + (size_t)slot * 32 + band;
      const float mu = __fadd_rn(__fmul_rn(a, *mup), __fmul_rn(one_m_a, feat));
      *mup = mu;
      const bool kill = (fl & (DPDF_FLAG_WARMUP_ | DPDF_FLAG_ZERO_FEAT_)) != 0;
      p.st.erb_ring[((size_t)slot * 3 + pos % 3) * 32 + band] = kill ? 0.f : (feat - mu) / 40.0f;
    }
  }
}

struct SynParams {
  IoDesc* io;
  Dims d;
  State st;
  const float* dft_inv;      // [F][win][2]
  const float* m;            // [B][fe0]
  const int* band_of_bin;    // [F]
  int B;
};

__global__ void __launch_bounds__(1024) k_synthesis(SynParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float2* Ys = reinterpret_cast<float2*>(smem);          // [F][SBT]
  const int win = p.d.win, hop = p.d.hop, F = p.d.F;
  const IoDesc* io = p.io;
  const int b0 = blockIdx.x * SBT;
  const int nb = min(SBT, p.B - b0);
  const int tid = threadIdx.x;
  const bool pcm_mode = io->mode == 0;

  __shared__ int s_slot[SBT], s_flag[SBT], s_pos[SBT];
  if (tid < SBT) {
    int b = b0 + tid;
    s_slot[tid] = b < p.B ? io_slot(io, b) : 0;
    s_flag[tid] = b < p.B ? io_flags(io, b) : 0;
    s_pos[tid] = b < p.B ? p.st.pos[s_slot[tid]] : 0;
  }
  __syncthreads();

  if (tid < F) {
    const int k = tid;
    for (int bb = 0; bb < SBT; ++bb) {
      float2 Y = make_float2(0.f, 0.f);
      if (bb < nb) {
        const int slot = s_slot[bb], pos = s_pos[bb];
        const int b = b0 + bb;
        float gain;
        if (s_flag[bb] & DPDF_FLAG_WARMUP_) gain = 0.f;
        else if (p.d.hr48) gain = p.m[(size_t)b * p.d.fe[0] + (k < p.d.fe[0] ? k : p.d.fe[0] - 2)];
        else gain = p.m[(size_t)b * 32 + p.band_of_bin[k]];
        const float2* mring = reinterpret_cast<const float2*>(p.st.mask_ring) + (size_t)slot * 3 * F;
        const float2 xd = mring[((pos + 1) % 3) * F + k];
        const float2 S = make_float2(xd.x * gain, xd.y * gain);
        float2* dring = reinterpret_cast<float2*>(p.st.dfspec_ring) + (size_t)slot * ORD * F;
        dring[(pos % ORD) * F + k] = S;
        if (k < NDF) {
          const float2* cf = reinterpret_cast<const float2*>(
              p.st.coef_ring + (((size_t)slot * 3 + (pos + 1) % 3) * NDF + k) * 2 * ORD);
          float rr = 0.f, ii = 0.f, ri = 0.f, ir = 0.f;
#pragma unroll
          for (int n = 0; n < ORD; ++n) {
            const float2 sn = (n == ORD - 1) ? S : dring[((pos + 1 + n) % ORD) * F + k];
            const float2 c = cf[n];
            rr = __fadd_rn(rr, __fmul_rn(sn.x, c.x));
            ii = __fadd_rn(ii, __fmul_rn(sn.y, c.y));
            ri = __fadd_rn(ri, __fmul_rn(sn.x, c.y));
            ir = __fadd_rn(ir, __fmul_rn(sn.y, c.x));
          }
          Y = make_float2(rr - ii, ri + ir);
        } else {
          Y = dring[((pos + 3) % ORD) * F + k];
        }
        if (!pcm_mode)
          reinterpret_cast<float2*>(io->out)[(size_t)b * F + k] = make_float2(Y.x * p.d.inv_wnorm, Y.y * p.d.inv_wnorm);
      }
      Ys[k * SBT + bb] = Y;
    }
  }
  float old[SBT];
#pragma unroll
  for (int bb = 0; bb < SBT; ++bb) old[bb] = 0.f;
  if (pcm_mode && tid < hop)
    for (int bb = 0; bb < nb; ++bb) old[bb] = p.st.ola[(size_t)s_slot[bb] * hop + tid];
  __syncthreads();

  if (pcm_mode && tid < win) {
    const int n = tid;
    float2 acc[SBT];
#pragma unroll
    for (int bb = 0; bb < SBT; ++bb) acc[bb] = make_float2(0.f, 0.f);
    const float2* basis = reinterpret_cast<const float2*>(p.dft_inv) + n;
    for (int k0 = 0; k0 < F; k0 += 16) {                     // 16 basis loads in flight (F = 161 / 481: one tail element)
      float2 cs[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) cs[u] = (k0 + u < F) ? __ldg(basis + (size_t)(k0 + u) * win) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (k0 + u >= F) break;
        const float4* yr = reinterpret_cast<const float4*>(Ys + (k0 + u) * SBT);
#pragma unroll
        for (int q = 0; q < SBT / 2; ++q) {
          float4 yy = yr[q];
          acc[2 * q] = ffma2(cs[u], lo2(yy), acc[2 * q]);
          acc[2 * q + 1] = ffma2(cs[u], hi2(yy), acc[2 * q + 1]);
        }
      }
    }
    const long long toff = (long long)io->t_out * hop;
    for (int bb = 0; bb < nb; ++bb) {
      const float fr = acc[bb].x + acc[bb].y;
      if (n < hop) io->out[(size_t)(b0 + bb) * io->out_stride + toff + n] = old[bb] + fr;
      else p.st.ola[(size_t)s_slot[bb] * hop + (n - hop)] = fr;
    }
  }
  if (tid < nb) p.st.pos[s_slot[tid]] = (s_pos[tid] + 1) % 15;      // only pos % 3 and pos % 5 are ever used: no int32 wrap after 248 days
}

__global__ void k_prime(State st, int hop, const float* pcm, long long stride, const int* slot_ids, int B) {
  const int b = blockIdx.x;
  const int slot = slot_ids ? slot_ids[b] : b;
  for (int n = threadIdx.x; n < hop; n += blockDim.x) st.in_hist[(size_t)slot * hop + n] = pcm[(size_t)b * stride + n];
}

struct ResetSeg { float* base; long long per_slot; const float* init; };
struct ResetParams { ResetSeg seg[16]; int nseg; int* pos; const int* slots; };

__global__ void k_reset(ResetParams p) {
  const int slot = p.slots ? p.slots[blockIdx.x] : blockIdx.x;
  for (int s = 0; s < p.nseg; ++s) {
    float* dst = p.seg[s].base + (size_t)slot * p.seg[s].per_slot;
    const float* init = p.seg[s].init;
    for (long long i = threadIdx.x; i < p.seg[s].per_slot; i += blockDim.x) dst[i] = init ? init[i] : 0.f;
  }
  if (threadIdx.x == 0) p.pos[slot] = 0;
}

// ---- launchers --------------------------------------------------------------------------------
void launch_prime(Engine& e, const float* pcm, long long stride, const int* slot_ids, int B, cudaStream_t st) {
  launch_k(e, k_prime, dim3(B), dim3(128), 0, st, e.st, e.d.hop, pcm, stride, slot_ids, B);
}

void launch_reset(Engine& e, const int* slots_dev, int n, cudaStream_t st) {
  const Dims& d = e.d;
  ResetParams p{};
  int i = 0;
  auto add = [&](float* base, long long per, const float* init) { p.seg[i++] = ResetSeg{base, per, init}; };
  add(e.st.mu, d.fe_feat, e.w.mu0);
  add(e.st.s, NDF, e.w.s0);
  add(e.st.erb_ring, 3LL * d.fe_feat, nullptr);
  add(e.st.df_ring, 3LL * 2 * NDF, nullptr);
  if (d.N > 0) {
    add(e.st.inter_erb, (long long)d.N * d.fe[3] * C, nullptr);
    add(e.st.inter_df, (long long)d.N * (NDF / 2) * C, nullptr);
  }
  add(e.st.h_enc, H, nullptr);
  add(e.st.h_erb, 2 * H, nullptr);
  add(e.st.h_df, 2 * H, nullptr);
  add(e.st.c0_ring, (long long)ORD * NDF * C, nullptr);
  add(e.st.dfp_acc, (long long)ORD * NDF * 10, nullptr);
  add(e.st.mask_ring, 3LL * d.F * 2, nullptr);
  add(e.st.coef_ring, 3LL * NDF * 2 * ORD, nullptr);
  add(e.st.dfspec_ring, (long long)ORD * d.F * 2, nullptr);
  add(e.st.in_hist, d.hop, nullptr);
  add(e.st.ola, d.hop, nullptr);
  p.nseg = i;
  p.pos = e.st.pos;
  p.slots = slots_dev;
  launch_k(e, k_reset, dim3(n), dim3(256), 0, st, p);
}

void launch_analysis(Engine& e, int B, cudaStream_t st) {
  AnaParams p{e.io_dev, e.d, e.st, e.w.dft_fwd, e.w.band_inv_w, e.w.band_start, B};
  const int ntg = (e.d.F + 31) / 32 * 32;
  // latency bound below ~2 CTAs per SM (split the sum: 5 groups at 16 kHz, 2 at 48 kHz), throughput bound above
  int ksplit = (std::max(B, e.total_B) + ABT - 1) / ABT <= 2 * e.num_sms ? 1024 / ntg : 1;   // total_B: all lanes of the step
  while (ksplit > 1 && ((e.d.win / 16) % ksplit != 0)) --ksplit; // every group walks whole 16-sample rounds
  const size_t smem = (size_t)(2 * e.d.win * ABT + ABT * e.d.F + 2 * (ksplit - 1) * ABT * e.d.F) * sizeof(float);
  if (ksplit > 1) launch_k(e, k_analysis<1024>, dim3((B + ABT - 1) / ABT), dim3(ntg * ksplit), smem, st, p, ksplit);
  else launch_k(e, k_analysis_one, dim3((B + ABT - 1) / ABT), dim3(ntg), smem, st, p);
}

void launch_synthesis(Engine& e, int B, cudaStream_t st) {
  SynParams p{e.io_dev, e.d, e.st, e.w.dft_inv, e.sc.m, e.w.band_of_bin, B};
  const int nt = (e.d.win + 31) / 32 * 32;
  const size_t smem = (size_t)(2 * e.d.F * SBT) * sizeof(float);
  launch_k(e, k_synthesis, dim3((B + SBT - 1) / SBT), dim3(nt), smem, st, p);
}

void init_frontend_kernels() {
  cudaFuncSetAttribute(k_analysis_one, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_analysis<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}

}  // namespace dpdf
