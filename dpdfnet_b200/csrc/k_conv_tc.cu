// k_sepconv_tc: the separable convolutions (a5 / a9: depthwise 1x3 with stride / sub-pixel / pathway add, or the
// grouped 3x3 of df_conv0, followed by a 64x64 pointwise conv + BN + ReLU; layers.py:761-834, 895-973) with the
// pointwise GEMM on the tensor cores.
//
// One CTA = 128 consecutive (stream, f) rows of one problem of the launch.  The prologue is the FP32 depthwise /
// grouped stage of k_sepconv, but its result goes straight into FP16 hi/lo K-major operand images (32 KB); thread 0
// issues the 12 tcgen05.mma.kind::f16 instructions of the error-compensated product against the 16 KB weight slab
// (one bulk copy), and the epilogue is thread = (TMEM lane = row, 32 columns): bias + ReLU -> XOR-swizzled FP32
// staging (over the dead images) -> coalesced store.  ~50 KB of shared memory and 64 TMEM columns per CTA: three to
// four CTAs per SM hide the prologue's load latency.
#include <algorithm>

#include "engine.h"
#include "tc_common.cuh"

namespace dpdf {

namespace {

using namespace tc;

// Threads per CTA (template parameter NT): the prologue is a per-thread serial chain (address math, three dependent
// loads, split) per row, so rows per thread set the kernel's latency (tools/ubench/sepconv_tc_timeline.cu): 512
// threads x 4 rows when the grid is small (latency bound), 256 threads x 8 rows at 3 CTAs per SM when it is not.
constexpr int SCT_ROWS = 128;
constexpr int SCT_IMG = SCT_ROWS * 64 * 2;       // bytes of one FP16 [128][64] image
constexpr int SCT_OFF_W = 2 * SCT_IMG;           // weight slab: hi | lo, 16 KB
constexpr int SCT_OFF_B = SCT_OFF_W + 2 * 64 * 64 * 2;
constexpr int SCT_OFF_BAR = SCT_OFF_B + 256;
constexpr size_t SCT_SMEM = SCT_OFF_BAR + 64;
constexpr int SCT_MAXP = 4;

}  // namespace

struct SepTcParams {
  const IoDesc* io;
  State st;
  SepProblem prob[SCT_MAXP];
  int nprob, B;
#ifdef SCT_TIMELINE
  long long* tl;          // [8] SM-clock stamps of CTA 0 (tools/ubench/sepconv_tc_timeline.cu)
#endif
};
#ifdef SCT_TIMELINE
#define STL(slot) do { if (blockIdx.x == 0 && threadIdx.x == 64) p.tl[slot] = clock64(); } while (0)
#else
#define STL(slot) do { } while (0)
#endif

template <int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 3) k_sepconv_tc(SepTcParams p) {
  constexpr int SCT_NT = NT, SCT_RPT = NT / 16;               // row stride of a thread: rows rsub, rsub + SCT_RPT, ...
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* Aimg = smem_raw;                               // hi image | lo image; later the FP32 output staging
  unsigned char* Wsm = smem_raw + SCT_OFF_W;
  float* bs = reinterpret_cast<float*>(smem_raw + SCT_OFF_B);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + SCT_OFF_BAR);   // [0] weights landed, [1] accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  int pi = 0;
#pragma unroll
  for (int i = 1; i < SCT_MAXP; ++i)
    if (i < p.nprob && (int)blockIdx.x >= p.prob[i].tile0) pi = i;
  const SepProblem& q = p.prob[pi];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  STL(0);
  const long long row0 = (long long)(blockIdx.x - q.tile0) * SCT_ROWS;
  const long long nrows = (long long)p.B * q.Fout;

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bars, 2 * 64 * 64 * 2);
    bulk_g2s(Wsm, q.tc_pw, 2 * 64 * 64 * 2, bars);
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  if (tid < 64) bs[tid] = __ldg(q.bias + tid);
  STL(1);
  uint32_t ovf = 0;                                        // FP16 range guard of the operand converter (tc_common.cuh)

  // prologue: A[row][c] -> operand images.  Thread = (channel quad g, row r & 7 ...): it keeps the same 4 channels for
  // all of its 8 rows (taps and pathway affine loaded once); the 8 lanes of a channel quad write 8 consecutive rows of
  // one 16-byte chunk column, so a warp's 8-byte stores fill two whole 128-byte core matrices (conflict-free).
  {
    const int g = (tid >> 3) & 15, c = g * 4;
    const int rsub = (tid & 7) | ((tid >> 7) << 3);
    const int nw = q.mode == 0 ? 3 * q.up : 9;
    float4 wt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = t < nw ? __ldg(reinterpret_cast<const float4*>(q.dw + (size_t)t * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
    if (q.mode == 0 && q.in2) {
      pa = __ldg(reinterpret_cast<const float4*>(q.pa + c));
      pb = __ldg(reinterpret_cast<const float4*>(q.pb + c));
    }
#pragma unroll
    for (int i = 0; i < SCT_ROWS / SCT_RPT; ++i) {
      const int r = rsub + SCT_RPT * i;
      const int row = (int)row0 + r;                          // < 2^31 rows: 32-bit index math (64-bit division is ~100 instructions)
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < (int)nrows) {
        const int b = row / q.Fout, fo = row - b * q.Fout;
        if (q.mode == 0) {
          int fc, j;
          if (q.up > 1) { fc = fo / q.up; j = fo % q.up; } else { fc = fo * q.stride; j = 0; }
          const float* in1 = q.in1 + (size_t)b * q.Fin * C;
          const float* in2 = q.in2 ? q.in2 + (size_t)b * q.Fin * C : nullptr;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const int fi = fc + t - 1;
            if (fi < 0 || fi >= q.Fin) continue;
            float4 v = __ldg(reinterpret_cast<const float4*>(in1 + (size_t)fi * C + c));
            if (in2) {
              const float4 u = __ldg(reinterpret_cast<const float4*>(in2 + (size_t)fi * C + c));
              v.x += fmaxf(fmaf(u.x, pa.x, pb.x), 0.f);
              v.y += fmaxf(fmaf(u.y, pa.y, pb.y), 0.f);
              v.z += fmaxf(fmaf(u.z, pa.z, pb.z), 0.f);
              v.w += fmaxf(fmaf(u.w, pa.w, pb.w), 0.f);
            }
            const float4 w = j == 0 ? wt[t] : (j == 1 ? wt[3 + t] : wt[6 + t]);
            acc.x = fmaf(w.x, v.x, acc.x);
            acc.y = fmaf(w.y, v.y, acc.y);
            acc.z = fmaf(w.z, v.z, acc.z);
            acc.w = fmaf(w.w, v.w, acc.w);
          }
        } else {
          const int slot = io_slot(p.io, b);
          const int pos = p.st.pos[slot];
          const float* ring = p.st.df_ring + (size_t)slot * 3 * 2 * NDF;
          const int plane = c >= 32 ? 1 : 0;
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) {
            const float* rowp = ring + (((pos + 1 + kt) % 3) * 2 + plane) * NDF;
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
              const int fi = fo + kf - 1;
              if (fi < 0 || fi >= NDF) continue;
              const float x = rowp[fi];
              const float4 w = wt[kt * 3 + kf];
              acc.x = fmaf(w.x, x, acc.x);
              acc.y = fmaf(w.y, x, acc.y);
              acc.z = fmaf(w.z, x, acc.z);
              acc.w = fmaf(w.w, x, acc.w);
            }
          }
        }
      }
      uint2 h, l;
      split2_f16(acc.x, acc.y, h.x, l.x);
      split2_f16(acc.z, acc.w, h.y, l.y);
      ovf |= f16_nonfinite(h.x) | f16_nonfinite(h.y);
      unsigned char* dst = Aimg + (r >> 3) * 1024 + (g >> 1) * 128 + (r & 7) * 16 + (g & 1) * 8;
      *reinterpret_cast<uint2*>(dst) = h;
      *reinterpret_cast<uint2*>(dst + SCT_IMG) = l;
    }
  }
  STL(2);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  STL(3);

  if (warp == 0 && elect_one()) {                              // elect.sync, not `tid == 0`: tc_common.cuh:elect_one
    mbar_wait(bars, 0);
    const uint32_t ah = smem_u32(Aimg), al = ah + SCT_IMG, bh = smem_u32(Wsm), bl = bh + 64 * 64 * 2;
    constexpr uint32_t IDESC = idesc_f16(128, 64);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t dah = umma_desc(ah + ks * 256, 1024), dal = umma_desc(al + ks * 256, 1024);
      const uint64_t dbh = umma_desc(bh + ks * 256, 1024), dbl = umma_desc(bl + ks * 256, 1024);
      umma_f16(tmem, dah, dbh, IDESC, ks > 0);
      umma_f16(tmem, dal, dbh, IDESC, 1);
      umma_f16(tmem, dah, dbl, IDESC, 1);
    }
    umma_commit(bars + 1);
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  STL(4);

  // epilogue: thread = (row = TMEM lane, 16-column group); bias + ReLU into the swizzled FP32 staging tile
  {
    const int qd = warp & 3, ch = warp >> 2, row = qd * 32 + lane;
    constexpr int CPW = 64 / (NT / 128);                        // accumulator columns per warp: 16 or 32
    const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16) + ch * CPW;
#pragma unroll
    for (int c16 = 0; c16 < CPW / 16; ++c16) {
      float v[16];
      tmem_ld16(ta + c16 * 16, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = ch * CPW + c16 * 16 + j * 4;
        const float4 o = make_float4(fmaxf(v[j * 4] + bs[col], 0.f), fmaxf(v[j * 4 + 1] + bs[col + 1], 0.f),
                                     fmaxf(v[j * 4 + 2] + bs[col + 2], 0.f), fmaxf(v[j * 4 + 3] + bs[col + 3], 0.f));
        *reinterpret_cast<float4*>(Aimg + row * 256 + ((((col >> 2)) ^ (row & 15)) << 4)) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  STL(5);
  for (int it = tid; it < SCT_ROWS * 16; it += SCT_NT) {
    const int r = it >> 4, c4 = it & 15, c = c4 * 4;
    const long long row = row0 + r;
    if (row >= nrows) continue;
    float4 v = *reinterpret_cast<const float4*>(Aimg + r * 256 + ((c4 ^ (r & 15)) << 4));
    if (q.mode == 0) {
      *reinterpret_cast<float4*>(q.out + (size_t)row * C + c) = v;
    } else {
      const int b = (int)(row / q.Fout), fo = (int)(row % q.Fout);
      const int slot = io_slot(p.io, b);
      const int pos = p.st.pos[slot];
      *reinterpret_cast<float4*>(q.out + (size_t)row * C + c) = v;      // c0 for df_conv1 (this hop)
      if (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.st.c0_fp16)
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.st.c0_ring + (size_t)slot * ORD * NDF * C) + ((size_t)(pos % ORD) * NDF + fo) * C + c) = pack4_f16(v);
      else
        *reinterpret_cast<float4*>(p.st.c0_ring + (((size_t)slot * ORD + pos % ORD) * NDF + fo) * C + c) = v;
    }
  }
  STL(6);
  if (ovf) p.io->err[DPDF_ERRW_RANGE] = 1;
  if (warp == 0) tmem_dealloc<64>(tmem);
}

void launch_sepconv_tc(Engine& e, const SepProblem* probs, int nprob, int B, cudaStream_t st) {
  SepTcParams p{};
  p.io = e.io_dev;
  p.st = e.st;
  p.nprob = nprob;
  p.B = B;
  int tiles = 0;
  for (int i = 0; i < nprob; ++i) {
    p.prob[i] = probs[i];
    p.prob[i].tile0 = tiles;
    tiles += (int)(((long long)B * probs[i].Fout + SCT_ROWS - 1) / SCT_ROWS);
  }
  if (std::max(B, e.total_B) < 4096) launch_k(e, k_sepconv_tc<512>, dim3(tiles), dim3(512), SCT_SMEM, st, p);
  else launch_k(e, k_sepconv_tc<256>, dim3(tiles), dim3(256), SCT_SMEM, st, p);
}

void init_conv_tc_kernels() {
  cudaFuncSetAttribute(k_sepconv_tc<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCT_SMEM);
  cudaFuncSetAttribute(k_sepconv_tc<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCT_SMEM);
}

}  // namespace dpdf
