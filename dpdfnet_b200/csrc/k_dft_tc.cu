// k_dft_tc: the framed real DFT of the analysis (a1; stream.py:119-126) and the inverse DFT + overlap-add of the synthesis
// (a13; stream.py:138-156) as tensor-core GEMMs over tiles of 128 streams.
//
//   analysis :  X[128 x 2F]   = frame[128 x win] * Bf[win x 2F]      frame = previous hop (state) | new hop (PCM)
//   synthesis:  y[128 x win]  = Y[128 x 2F]      * Bi[2F x win]      then out = ola + y[:hop], ola = y[hop:]
//
// The FFMA2 kernels (k_frontend.cu) stream the (cos, sin) basis from L2 once per 8-16 streams - 0.4 MB at 16 kHz, 3.7 MB at
// 48 kHz - and are bound by exactly that; here a CTA pulls the slabs of its column chunk once per 128 streams, as FP16
// hi | lo operand images (weights.py:dft_tc_images; the window and wnorm are folded in), and every product is the usual
// three tcgen05.mma.kind::f16 passes hi*hi + lo*hi + hi*lo with FP32 accumulation in tensor memory.
//
// Structure (as k_gru_tc): K is walked in 64-wide stages; eight converter warps load the FP32 activation chunk of the
// stage after next while they split and store the current one as K-major operand images (2-deep ring); the issuer warp
// streams the [NC x 64] basis slabs through a 3-deep ring of bulk copies and fires 12 MMAs per stage; tcgen05.commit
// frees both rings.  Epilogue: thread = (TMEM lane = stream, column half).
//   analysis : NC = 128 interleaved (re, im) columns per CTA -> spectrum scratch [B][ncol]; features, normalisers, ring
//              pushes and the history update stay in k_analysis (k_frontend.cu), which then skips its own DFT.
//   synthesis: a CTA owns 80 samples n of the first frame half AND the samples n + hop of the second (NC = 160), so the
//              same thread reads the old overlap-add tail, emits out[n] and writes the new tail ola[n]: no two CTAs ever
//              touch the same tail element.  The masked / deep-filtered spectrum Y comes from k_synthesis (which then skips
//              its own inverse DFT) through a zero-padded scratch [B][kpad].
#include "engine.h"
#include "tc_common.cuh"

namespace dpdf {

namespace {

using namespace tc;

constexpr int DT_CONV = 256;                 // converter / epilogue threads
constexpr int DT_NT = DT_CONV + 32;          // + issuer warp
constexpr int DT_AIMG = 128 * 64 * 2;        // one FP16 [128][64] image
constexpr int DT_OFF_W = 2 * 2 * DT_AIMG;    // A ring: 2 stages x (hi | lo)
constexpr int DT_NW = 3;                     // basis slab ring depth
constexpr int DT_MAXSLAB = 2 * 160 * 64 * 2; // [160][64] hi | lo
constexpr int DT_MAXSTAGE = 16;              // K <= 1024
constexpr int DT_OFF_MISC = DT_OFF_W + DT_NW * DT_MAXSLAB;
constexpr size_t DFT_TC_SMEM = DT_OFF_MISC + 128 * 4 + (DT_NW + DT_MAXSTAGE) * 8 + 16;

}  // namespace

struct DftTcParams {
  IoDesc* io;
  State st;
  const float* wimg;       // [chunk][stage][NC x 64] hi | lo operand images
  const float* scale;      // [2] powers of two: the activations are multiplied by [0] before the FP16 split (keeps `lo` out of the
                           // subnormals down to -120 dBFS), the accumulators by [1] = 1 / ([0] * basis scale) (weights.py:dft_tc_images)
  float* spec;             // analysis: spectrum scratch (written); synthesis: Y scratch (read)
  int ld;                  // floats per stream of `spec`
  int hop, nstage, B;
};

template <int NC, int SYN>
__global__ void __launch_bounds__(DT_NT, 1) k_dft_tc(DftTcParams p) {
  constexpr int SLAB = 2 * NC * 64 * 2;
  pdl_trigger();
  if (p.io->mode != 0) return;                                // spectrum in / out (the ONNX call shape): nothing to transform
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* Asm = smem_raw;
  unsigned char* Wsm = smem_raw + DT_OFF_W;
  int* s_slot = reinterpret_cast<int*>(smem_raw + DT_OFF_MISC);
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem_raw + DT_OFF_MISC + 128 * 4);   // [DT_NW]
  uint64_t* done = full_w + DT_NW;                                                    // [nstage], single use
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + DT_MAXSTAGE);

  const IoDesc* io = p.io;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * 128, chunk = blockIdx.y;
  const int valid = min(128, p.B - b0);
  const int NS = p.nstage;

  if (tid == 0) {
    for (int i = 0; i < DT_NW + NS; ++i) mbar_init(full_w + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<(NC > 128 ? 256 : 128)>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---- issuer warp: basis slab ring + MMAs --------------------------------------------------------------------
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wimg) + (size_t)chunk * NS * SLAB;
    auto load_w = [&](int s) {
      mbar_expect_tx(full_w + s % DT_NW, SLAB);
      bulk_g2s(Wsm + (s % DT_NW) * SLAB, wsrc + (size_t)s * SLAB, SLAB, full_w + s % DT_NW);
    };
    if (elect_one())                                         // elect.sync, not `lane == 0`: tc_common.cuh:elect_one
      for (int s = 0; s < DT_NW && s < NS; ++s) load_w(s);
    pdl_wait();
    for (int s = 0; s < NS; ++s) {
      if ((s & 1) == 0) asm volatile("bar.sync 1, %0;" ::"n"(DT_NT) : "memory");      // A images of stage s written
      else asm volatile("bar.sync 2, %0;" ::"n"(DT_NT) : "memory");
      if (elect_one()) {
        tc_fence_after();
        mbar_wait(full_w + s % DT_NW, (s / DT_NW) & 1);
        const uint32_t ah = smem_u32(Asm) + (s & 1) * 2 * DT_AIMG, al = ah + DT_AIMG;
        const uint32_t bh = smem_u32(Wsm) + (s % DT_NW) * SLAB, bl = bh + SLAB / 2;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t dah = umma_desc(ah + ks * 256, 1024), dal = umma_desc(al + ks * 256, 1024);
          const uint64_t dbh = umma_desc(bh + ks * 256, 1024), dbl = umma_desc(bl + ks * 256, 1024);
          umma_f16(tmem, dah, dbh, idesc_f16(128, NC), (s > 0 || ks > 0) ? 1u : 0u);
          umma_f16(tmem, dal, dbh, idesc_f16(128, NC), 1);
          umma_f16(tmem, dah, dbl, idesc_f16(128, NC), 1);
        }
        umma_commit(done + s);
        if (s >= 1 && s + DT_NW - 1 < NS) {                   // the slab slot of stage s - 1 is free once its MMAs are done
          mbar_wait(done + s - 1, 0);
          load_w(s + DT_NW - 1);
        }
      }
      __syncwarp();
    }
  } else {
    // ---- converter warps: activation chunk -> operand images ---------------------------------------------------
    // thread = (k quad g, row r & 7 ...): the 8 lanes of a k quad write 8 consecutive rows of one 16-byte chunk column,
    // so a warp's 8-byte stores fill two whole 128-byte core matrices
    const int g = (tid >> 3) & 15;
    const int rsub = (tid & 7) | ((tid >> 7) << 3);
    pdl_wait();                                               // PCM / history / Y come from the kernels before this one
    if (tid < 128) s_slot[tid] = tid < valid ? io_slot(io, b0 + tid) : 0;
    asm volatile("bar.sync 3, %0;" ::"n"(DT_CONV) : "memory");
    const long long toff_in = (long long)io->t_in * p.hop;
    const bool vec_in = ((io->in_stride | toff_in) & 3) == 0 && (reinterpret_cast<size_t>(io->in) & 15) == 0;
    auto load_stage = [&](int s, float4 (&v)[8]) {
      const int n0 = s * 64 + g * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= valid) continue;
        if constexpr (SYN) {
          v[i] = __ldg(reinterpret_cast<const float4*>(p.spec + (size_t)(b0 + r) * p.ld + n0));
        } else if (n0 < p.hop) {                              // first half of the frame: the previous hop (state)
          v[i] = *reinterpret_cast<const float4*>(p.st.in_hist + (size_t)s_slot[r] * p.hop + n0);
        } else {
          const float* src = io->in + (size_t)(b0 + r) * io->in_stride + toff_in + (n0 - p.hop);
          if (vec_in) v[i] = __ldg(reinterpret_cast<const float4*>(src));
          else v[i] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
        }
      }
    };
    uint32_t ovf = 0;                                         // FP16 range guard (tc_common.cuh:f16_nonfinite)
    const float s_in = __ldg(p.scale), s_out = __ldg(p.scale + 1);
    auto convert_stage = [&](int s, const float4 (&v)[8]) {
      if (s >= 2) mbar_wait(done + s - 2, 0);                 // MMAs that read this image pair are complete
      unsigned char* img = Asm + (s & 1) * 2 * DT_AIMG;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        uint2 h, l;
        split2_f16(v[i].x * s_in, v[i].y * s_in, h.x, l.x);
        split2_f16(v[i].z * s_in, v[i].w * s_in, h.y, l.y);
        ovf |= f16_nonfinite(h.x) | f16_nonfinite(h.y);
        unsigned char* dst = img + (r >> 3) * 1024 + (g >> 1) * 128 + (r & 7) * 16 + (g & 1) * 8;
        *reinterpret_cast<uint2*>(dst) = h;
        *reinterpret_cast<uint2*>(dst + DT_AIMG) = l;
      }
      fence_async_smem();
      if ((s & 1) == 0) asm volatile("bar.arrive 1, %0;" ::"n"(DT_NT) : "memory");
      else asm volatile("bar.arrive 2, %0;" ::"n"(DT_NT) : "memory");
    };
    float4 va[8], vb[8];
    load_stage(0, va);
#pragma unroll 1
    for (int s = 0; s < NS; s += 2) {
      if (s + 1 < NS) load_stage(s + 1, vb);
      convert_stage(s, va);
      if (s + 1 < NS) {
        if (s + 2 < NS) load_stage(s + 2, va);
        convert_stage(s + 1, vb);
      }
    }
    if (ovf) p.io->err[DPDF_ERRW_RANGE] = 1;
    // ---- epilogue: thread = (stream row = TMEM lane, column half) --------------------------------------------------
    mbar_wait(done + NS - 1, 0);
    tc_fence_after();
    const int qd = warp & 3, half = warp >> 2, row = qd * 32 + lane;
    const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16);
    if constexpr (!SYN) {
      float* dst = p.spec + (size_t)(b0 + row) * p.ld + chunk * NC + half * (NC / 2);
#pragma unroll 1
      for (int c = 0; c < NC / 16; ++c) {
        uint32_t v[8];
        tmem_ld8_nowait(ta + half * (NC / 2) + c * 8, v);
        tmem_ld_wait();
        if (row < valid) {
          *reinterpret_cast<float4*>(dst + c * 8) = make_float4(__uint_as_float(v[0]) * s_out, __uint_as_float(v[1]) * s_out, __uint_as_float(v[2]) * s_out, __uint_as_float(v[3]) * s_out);
          *reinterpret_cast<float4*>(dst + c * 8 + 4) = make_float4(__uint_as_float(v[4]) * s_out, __uint_as_float(v[5]) * s_out, __uint_as_float(v[6]) * s_out, __uint_as_float(v[7]) * s_out);
        }
      }
    } else {
      constexpr int W = NC / 2;                              // samples per frame half and CTA
      const long long toff = (long long)io->t_out * p.hop;
      const bool vec = ((io->out_stride | toff) & 3) == 0 && (reinterpret_cast<size_t>(io->out) & 15) == 0;
      const int n0 = chunk * W + half * (W / 2);
      float* outp = io->out + (size_t)(b0 + row) * io->out_stride + toff + n0;
      float* olap = p.st.ola + (size_t)s_slot[row] * p.hop + n0;
#pragma unroll 1
      for (int c = 0; c < W / 16; ++c) {
        uint32_t f0[8], f1[8];
        tmem_ld8_nowait(ta + half * (W / 2) + c * 8, f0);        // y[n]
        tmem_ld8_nowait(ta + W + half * (W / 2) + c * 8, f1);    // y[n + hop]
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
        if (row < valid) {
          t0 = *reinterpret_cast<const float4*>(olap + c * 8);
          t1 = *reinterpret_cast<const float4*>(olap + c * 8 + 4);
        }
        tmem_ld_wait();
        if (row < valid) {
          const float4 o0 = make_float4(t0.x + __uint_as_float(f0[0]) * s_out, t0.y + __uint_as_float(f0[1]) * s_out, t0.z + __uint_as_float(f0[2]) * s_out, t0.w + __uint_as_float(f0[3]) * s_out);
          const float4 o1 = make_float4(t1.x + __uint_as_float(f0[4]) * s_out, t1.y + __uint_as_float(f0[5]) * s_out, t1.z + __uint_as_float(f0[6]) * s_out, t1.w + __uint_as_float(f0[7]) * s_out);
          if (vec) {
            *reinterpret_cast<float4*>(outp + c * 8) = o0;
            *reinterpret_cast<float4*>(outp + c * 8 + 4) = o1;
          } else {
            outp[c * 8] = o0.x; outp[c * 8 + 1] = o0.y; outp[c * 8 + 2] = o0.z; outp[c * 8 + 3] = o0.w;
            outp[c * 8 + 4] = o1.x; outp[c * 8 + 5] = o1.y; outp[c * 8 + 6] = o1.z; outp[c * 8 + 7] = o1.w;
          }
          *reinterpret_cast<float4*>(olap + c * 8) = make_float4(__uint_as_float(f1[0]) * s_out, __uint_as_float(f1[1]) * s_out, __uint_as_float(f1[2]) * s_out, __uint_as_float(f1[3]) * s_out);
          *reinterpret_cast<float4*>(olap + c * 8 + 4) = make_float4(__uint_as_float(f1[4]) * s_out, __uint_as_float(f1[5]) * s_out, __uint_as_float(f1[6]) * s_out, __uint_as_float(f1[7]) * s_out);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<(NC > 128 ? 256 : 128)>(tmem);
}

// Framed DFT of the hop for B streams -> spectrum scratch [B][ncol] (interleaved re / im, wnorm applied)
void launch_dft_tc(Engine& e, int B, cudaStream_t st) {
  DftTcParams p{e.io_dev, e.st, e.w.dft_fwd_tc, e.w.dft_tc_scale, e.sc.spec_tc, e.spec_tc_ld, e.d.hop, e.d.win / 64, B};
  launch_k(e, k_dft_tc<128, 0>, dim3((B + 127) / 128, e.spec_tc_ld / 128), dim3(DT_NT), DFT_TC_SMEM, st, p);
}

// Inverse DFT of the Y scratch [B][kpad] + overlap-add -> PCM out, new tail
void launch_idft_tc(Engine& e, int B, cudaStream_t st) {
  DftTcParams p{e.io_dev, e.st, e.w.dft_inv_tc, e.w.dft_tc_scale + 2, e.sc.yspec_tc, e.yspec_tc_ld, e.d.hop, e.yspec_tc_ld / 64, B};
  launch_k(e, k_dft_tc<160, 1>, dim3((B + 127) / 128, e.d.hop / 80), dim3(DT_NT), DFT_TC_SMEM, st, p);
}

void init_dft_tc_kernels() {
  cudaFuncSetAttribute(k_dft_tc<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DFT_TC_SMEM);
  cudaFuncSetAttribute(k_dft_tc<160, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DFT_TC_SMEM);
}

}  // namespace dpdf
