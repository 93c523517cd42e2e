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

struct AnaParams {
  const IoDesc* io;
  Dims d;
  State st;
  const float* dft_fwd;      // [win][F][2]
  const float* band_inv_w;   // [32]
  const int* band_start;     // [33]
  int B;
  const float* pre_spec;     // spectrum of this hop from k_dft_tc ([B][pre_ld] floats, interleaved re / im, wnorm applied), or nullptr
  int pre_ld;
};

// One CTA = NB streams, thread = frequency bin.  The windowed DFT is a [NB x win] x [win x F] product whose (cos, sin)
// basis (0.4 MB at 16 kHz, 3.7 MB at 48 kHz) streams from L2 once per CTA: NB sets the L2 traffic per stream (ncu, round 1:
// 8 streams per CTA made the 48 kHz kernel L2-bandwidth bound at 2.7 TB/s), so throughput-sized grids use 16 or 32 streams
// per CTA and small grids 8 with the sum over the win samples split over `ksplit` thread groups (the per-thread chain of
// win dependent load + FFMA2 rounds is what bounds a small grid; partial spectra are added through shared memory in a
// fixed order, g = 0 first).  Accumulators pair two STREAMS per FFMA2 - (re_b, re_b+1) += (cos, cos) * (x_b, x_b+1) - so
// the PCM tile sits in shared memory once, sample-major, and one LDS.128 feeds four FFMA2.
// PCM staging: float4 loads of the hop each stream contributes (16-byte aligned rows; scalar otherwise), streams fastest
// across the threads so that the transposed stores are conflict free.
template <int NB, int MAXT>
__global__ void __launch_bounds__(MAXT) k_analysis(AnaParams p, int ksplit) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  const int win = p.d.win, hop = p.d.hop, F = p.d.F;
  float* xs = smem;                                               // [win][NB]
  float* pws = smem + win * NB;                                   // [NB][F]   (16 kHz: band energies)
  float2* part = reinterpret_cast<float2*>(pws + (p.d.hr48 ? 0 : NB * F));   // [ksplit - 1][NB][F] partial spectra of groups 1..
  const IoDesc* io = p.io;
  const int b0 = blockIdx.x * NB;
  const int nb = min(NB, p.B - b0);
  const int NT = blockDim.x, NTg = NT / ksplit;
  const int grp = threadIdx.x / NTg;
  const int tid = threadIdx.x;
  const bool pcm_mode = io->mode == 0;

  __shared__ int s_slot[NB], s_flag[NB], s_pos[NB];
  if (tid < NB) {
    int b = b0 + tid;
    s_slot[tid] = b < p.B ? io_slot(io, b) : 0;
    s_flag[tid] = b < p.B ? io_flags(io, b) : 0;
    s_pos[tid] = b < p.B ? p.st.pos[s_slot[tid]] : 0;
  }
  __syncthreads();

  float2 Xr[NB / 2], Xi[NB / 2];                                 // real / imaginary parts of streams (2q, 2q + 1)
#pragma unroll
  for (int q = 0; q < NB / 2; ++q) Xr[q] = Xi[q] = make_float2(0.f, 0.f);

  if (pcm_mode && p.pre_spec) {
    // the DFT of this hop came from the tensor cores (k_dft_tc.cu): only the history update is left of the PCM side
    const long long toff = (long long)io->t_in * hop;
    const bool vec = ((io->in_stride | toff) & 3) == 0 && (reinterpret_cast<size_t>(io->in) & 15) == 0;
    for (int i = tid; i < NB * (hop / 4); i += NT) {
      const int bb = i / (hop / 4), n = (i % (hop / 4)) * 4;
      if (bb < nb) {
        const float* src = io->in + (size_t)(b0 + bb) * io->in_stride + toff + n;
        const float4 v = vec ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
        *reinterpret_cast<float4*>(p.st.in_hist + (size_t)s_slot[bb] * hop + n) = v;
      }
    }
    if (tid < F) {
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        if (bb >= nb) break;
        const float2 v = __ldg(reinterpret_cast<const float2*>(p.pre_spec + (size_t)(b0 + bb) * p.pre_ld) + tid);
        if (bb & 1) { Xr[bb / 2].y = v.x; Xi[bb / 2].y = v.y; }
        else { Xr[bb / 2].x = v.x; Xi[bb / 2].x = v.y; }
      }
    }
  } else if (pcm_mode) {
    const long long toff = (long long)io->t_in * hop;
    const bool vec = ((io->in_stride | toff) & 3) == 0 && (reinterpret_cast<size_t>(io->in) & 15) == 0;
    for (int i = tid; i < NB * (win / 4); i += NT) {              // item = (4 consecutive samples, stream); streams fastest
      const int bb = i % NB, n = (i / NB) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bb < nb) {
        if (n < hop) {
          v = *reinterpret_cast<const float4*>(p.st.in_hist + (size_t)s_slot[bb] * hop + n);
        } else {
          const float* src = io->in + (size_t)(b0 + bb) * io->in_stride + toff + (n - hop);
          if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
          else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
        }
      }
      xs[n * NB + bb] = v.x; xs[(n + 1) * NB + bb] = v.y; xs[(n + 2) * NB + bb] = v.z; xs[(n + 3) * NB + bb] = v.w;
    }
    __syncthreads();
    for (int i = tid; i < NB * (hop / 4); i += NT) {              // history <- this hop (after all reads above)
      const int bb = i % NB, n = (i / NB) * 4;
      if (bb < nb)
        *reinterpret_cast<float4*>(p.st.in_hist + (size_t)s_slot[bb] * hop + n) =
            make_float4(xs[(n + hop) * NB + bb], xs[(n + hop + 1) * NB + bb], xs[(n + hop + 2) * NB + bb], xs[(n + hop + 3) * NB + bb]);
    }
    const int kb = tid - grp * NTg;                               // bin of this thread inside its group
    if (kb < F) {
      // the (cos, sin) basis streams from L2: keep 16 loads in flight
      const float2* basis = reinterpret_cast<const float2*>(p.dft_fwd) + kb;
      const int nper = win / ksplit;
      for (int n0 = grp * nper; n0 < (grp + 1) * nper; n0 += 16) {
        float2 cs[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) cs[u] = __ldg(basis + (size_t)(n0 + u) * F);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float2 c2 = make_float2(cs[u].x, cs[u].x), s2 = make_float2(cs[u].y, cs[u].y);
          const float4* xr = reinterpret_cast<const float4*>(xs + (n0 + u) * NB);
#pragma unroll
          for (int q = 0; q < NB / 4; ++q) {
            const float4 xx = xr[q];
            Xr[2 * q] = ffma2(c2, lo2(xx), Xr[2 * q]);
            Xi[2 * q] = ffma2(s2, lo2(xx), Xi[2 * q]);
            Xr[2 * q + 1] = ffma2(c2, hi2(xx), Xr[2 * q + 1]);
            Xi[2 * q + 1] = ffma2(s2, hi2(xx), Xi[2 * q + 1]);
          }
        }
      }
      if (grp > 0) {
#pragma unroll
        for (int q = 0; q < NB / 2; ++q) {
          part[((grp - 1) * NB + q) * F + kb] = Xr[q];
          part[((grp - 1) * NB + NB / 2 + q) * F + kb] = Xi[q];
        }
      }
    }
    if (ksplit > 1) {
      __syncthreads();
      if (tid < F) {
        for (int g = 1; g < ksplit; ++g)
#pragma unroll
          for (int q = 0; q < NB / 2; ++q) {
            const float2 vr = part[((g - 1) * NB + q) * F + tid], vi = part[((g - 1) * NB + NB / 2 + q) * F + tid];
            Xr[q].x += vr.x; Xr[q].y += vr.y;
            Xi[q].x += vi.x; Xi[q].y += vi.y;
          }
      }
    }
  } else if (tid < F) {
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      if (bb >= nb) break;
      const float2 v = __ldg(reinterpret_cast<const float2*>(io->in) + (size_t)(b0 + bb) * F + tid);
      if (bb & 1) { Xr[bb / 2].y = v.x * p.d.wnorm; Xi[bb / 2].y = v.y * p.d.wnorm; }
      else { Xr[bb / 2].x = v.x * p.d.wnorm; Xi[bb / 2].x = v.y * p.d.wnorm; }
    }
  }

  const float a = 0.98f, one_m_a = 0.02f;     // float32(0.98), float32(1 - 0.98)
  if (tid < F) {
    const int k = tid;
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      if (bb >= nb) break;
      const int slot = s_slot[bb], fl = s_flag[bb], pos = s_pos[bb];
      float re = (bb & 1) ? Xr[bb / 2].y : Xr[bb / 2].x, im = (bb & 1) ? Xi[bb / 2].y : Xi[bb / 2].x;
      if (fl & DPDF_FLAG_ZERO_SPEC_) re = im = 0.f;
      reinterpret_cast<float2*>(p.st.mask_ring)[((size_t)slot * 3 + pos % 3) * F + k] = make_float2(re, im);
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
  float* yout;               // PCM mode: hand Y to k_dft_tc ([B][yld] floats, interleaved re / im) instead of the inverse DFT here, or nullptr
  int yld;
};

// One CTA = SB streams.  Prologue: thread = bin (mask, deep filter, spectral rings); inverse DFT: thread = output sample,
// the [F][win] (cos, sin) basis streams from L2 once per CTA (SB amortises it, as NB does in k_analysis) and the
// accumulators pair two streams per FFMA2: (y_b, y_b+1) += (cos, cos) * (re_b, re_b+1) + (sin, sin) * (im_b, im_b+1).
// DEPTH basis loads are kept in flight per thread (8 where 16 streams x 1024 threads leave 64 registers per thread).
template <int SB, int DEPTH, int MAXT>
__global__ void __launch_bounds__(MAXT) k_synthesis(SynParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* Ys = smem;                                      // [F][2][SB]: real parts of the SB streams, then imaginary parts
  const int win = p.d.win, hop = p.d.hop, F = p.d.F;
  const IoDesc* io = p.io;
  const int b0 = blockIdx.x * SB;
  const int nb = min(SB, p.B - b0);
  const int tid = threadIdx.x;
  const bool pcm_mode = io->mode == 0;

  __shared__ int s_slot[SB], s_flag[SB], s_pos[SB];
  if (tid < SB) {
    int b = b0 + tid;
    s_slot[tid] = b < p.B ? io_slot(io, b) : 0;
    s_flag[tid] = b < p.B ? io_flags(io, b) : 0;
    s_pos[tid] = b < p.B ? p.st.pos[s_slot[tid]] : 0;
  }
  __syncthreads();

  if (tid < F) {
    const int k = tid;
    for (int bb = 0; bb < SB; ++bb) {
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
      if (p.yout) {
        if (bb < nb && pcm_mode) reinterpret_cast<float2*>(p.yout + (size_t)(b0 + bb) * p.yld)[k] = Y;
      } else {
        Ys[(k * 2) * SB + bb] = Y.x;
        Ys[(k * 2 + 1) * SB + bb] = Y.y;
      }
    }
  }
  if (p.yout) {                                            // inverse DFT + overlap-add follow on the tensor cores (k_dft_tc.cu)
    if (tid < nb) p.st.pos[s_slot[tid]] = (s_pos[tid] + 1) % 15;
    return;
  }
  __syncthreads();

  float2 accr[SB / 2], acci[SB / 2];
#pragma unroll
  for (int q = 0; q < SB / 2; ++q) accr[q] = acci[q] = make_float2(0.f, 0.f);
  const int n = tid;
  if (pcm_mode && tid < win) {
    const float2* basis = reinterpret_cast<const float2*>(p.dft_inv) + n;
    for (int k0 = 0; k0 < F; k0 += DEPTH) {                  // DEPTH basis loads in flight (F = 161 / 481: one tail element)
      float2 cs[DEPTH];
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) cs[u] = (k0 + u < F) ? __ldg(basis + (size_t)(k0 + u) * win) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) {
        if (k0 + u >= F) break;
        const float2 c2 = make_float2(cs[u].x, cs[u].x), s2 = make_float2(cs[u].y, cs[u].y);
        const float4* yr = reinterpret_cast<const float4*>(Ys + (k0 + u) * 2 * SB);
#pragma unroll
        for (int q = 0; q < SB / 4; ++q) {
          const float4 re = yr[q], im = yr[SB / 4 + q];
          accr[2 * q] = ffma2(c2, lo2(re), accr[2 * q]);
          acci[2 * q] = ffma2(s2, lo2(im), acci[2 * q]);
          accr[2 * q + 1] = ffma2(c2, hi2(re), accr[2 * q + 1]);
          acci[2 * q + 1] = ffma2(s2, hi2(im), acci[2 * q + 1]);
        }
      }
    }
  }
  // overlap-add: the first half of the frame completes the previous hop's tail and leaves, the second half becomes the new
  // tail - read by thread n, written by thread n + hop, hence the barrier between the two
  const long long toff = (long long)io->t_out * hop;
  if (pcm_mode && n < hop) {
#pragma unroll
    for (int bb = 0; bb < SB; ++bb) {
      if (bb >= nb) break;
      const float fr = (bb & 1) ? accr[bb / 2].y + acci[bb / 2].y : accr[bb / 2].x + acci[bb / 2].x;
      io->out[(size_t)(b0 + bb) * io->out_stride + toff + n] = p.st.ola[(size_t)s_slot[bb] * hop + n] + fr;
    }
  }
  __syncthreads();
  if (pcm_mode && n >= hop && n < win) {
#pragma unroll
    for (int bb = 0; bb < SB; ++bb) {
      if (bb >= nb) break;
      const float fr = (bb & 1) ? accr[bb / 2].y + acci[bb / 2].y : accr[bb / 2].x + acci[bb / 2].x;
      p.st.ola[(size_t)s_slot[bb] * hop + (n - hop)] = fr;
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

template <int NB, int MAXT>
static void launch_analysis_t(Engine& e, const AnaParams& p, int B, int ntg, int ksplit, cudaStream_t st) {
  const size_t smem = (size_t)(e.d.win * NB + (e.d.hr48 ? 0 : NB * e.d.F) + 2 * (ksplit - 1) * NB * e.d.F) * sizeof(float);
  launch_k(e, k_analysis<NB, MAXT>, dim3((B + NB - 1) / NB), dim3(ntg * ksplit), smem, st, p, ksplit);
}

bool dft_on_tc(const Engine& e, int B) {
  return e.dft_tc == 1 || (e.dft_tc == 2 && std::max(B, e.total_B) >= e.dft_tc_min);
}

void launch_analysis(Engine& e, int B, cudaStream_t st) {
  AnaParams p{e.io_dev, e.d, e.st, e.w.dft_fwd, e.w.band_inv_w, e.w.band_start, B, nullptr, 0};
  const int ntg = (e.d.F + 31) / 32 * 32;
  const int Bt = std::max(B, e.total_B);          // total_B: all lanes of the step
  if (dft_on_tc(e, B)) {
    // spectrum from the tensor cores (k_dft_tc; a no-op in spectrum mode): features only, no reason to keep CTAs small
    p.pre_spec = e.sc.spec_tc;
    p.pre_ld = e.spec_tc_ld;
    if ((Bt + 7) / 8 <= 4 * e.num_sms) launch_analysis_t<8, 1024>(e, p, B, ntg, 1, st);
    else launch_analysis_t<16, 512>(e, p, B, ntg, 1, st);
    return;
  }
  // latency bound while 8-stream CTAs leave SMs idle: split the sum over thread groups (5 at 16 kHz, 2 at 48 kHz); above
  // that the basis traffic per stream decides.  Measured (profiles/r2i_sweep.log): 16 streams per CTA wins from 2048
  // streams up at both rates (48 kHz, 2048 streams: 0.33 -> 0.17 ms; 16 kHz, 16384: 0.32 -> 0.25 ms with the synthesis
  // kernel at 8); 32 per CTA is slower again (fewer, longer CTAs)
  const int force = e.ana_force;
  if (force ? force == 8 : (Bt + 7) / 8 <= e.num_sms + e.num_sms / 2) {
    int ksplit = (Bt + 7) / 8 <= 2 * e.num_sms ? 1024 / ntg : 1;
    while (ksplit > 1 && ((e.d.win / 16) % ksplit != 0)) --ksplit;   // every group walks whole 16-sample rounds
    launch_analysis_t<8, 1024>(e, p, B, ntg, ksplit, st);
  } else if (force ? force == 32 : e.ana_nb >= 64) {
    launch_analysis_t<32, 512>(e, p, B, ntg, 1, st);
  } else if (force ? force == 16 : e.ana_nb >= 16) {
    launch_analysis_t<16, 512>(e, p, B, ntg, 1, st);
  } else {
    launch_analysis_t<8, 1024>(e, p, B, ntg, 1, st);
  }
}

void launch_synthesis(Engine& e, int B, cudaStream_t st) {
  SynParams p{e.io_dev, e.d, e.st, e.w.dft_inv, e.sc.m, e.w.band_of_bin, B, nullptr, 0};
  const int nt = (e.d.win + 31) / 32 * 32;
  const int Bt = std::max(B, e.total_B);
  auto go = [&](auto kernel, int SB) {
    launch_k(e, kernel, dim3((B + SB - 1) / SB), dim3(nt), (size_t)(2 * e.d.F * SB) * sizeof(float), st, p);
  };
  if (dft_on_tc(e, B)) {
    // mask + deep filter only: Y goes to the scratch of k_dft_tc, which does the inverse DFT and the overlap-add
    p.yout = e.sc.yspec_tc;
    p.yld = e.yspec_tc_ld;
    const int ntf = (e.d.F + 31) / 32 * 32;
    if (ntf > 384) launch_k(e, k_synthesis<4, 16, 1024>, dim3((B + 3) / 4), dim3(ntf), 0, st, p);
    else launch_k(e, k_synthesis<4, 16, 384>, dim3((B + 3) / 4), dim3(ntf), 0, st, p);
    return;
  }
  // 16 kHz: 320 threads per CTA (register budget is not an issue); 48 kHz: 960 threads, 64 registers each
  const int sforce = e.syn_force;
  if (sforce ? sforce == 16 : e.syn_sb >= 64) {
    if (nt > 384) go(k_synthesis<8, 8, 1024>, 8);       // 16 streams x 960 threads do not fit 64 registers (measured slower: spills)
    else go(k_synthesis<16, 16, 384>, 16);
  } else if (sforce ? sforce == 8 : (Bt >= 2048 && e.syn_sb >= 8)) {
    if (nt > 384) go(k_synthesis<8, 8, 1024>, 8); else go(k_synthesis<8, 16, 384>, 8);
  } else {
    if (nt > 384) go(k_synthesis<4, 16, 1024>, 4); else go(k_synthesis<4, 16, 384>, 4);
  }
}

void init_frontend_kernels() {
  cudaFuncSetAttribute(k_analysis<8, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_analysis<16, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_analysis<32, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_synthesis<16, 16, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

}  // namespace dpdf
