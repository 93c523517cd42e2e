// k_sepconv_tma: the separable convolutions of a hop (a5 / a9; layers.py:761-834, 895-973) as a PERSISTENT,
// TMA-fed tensor-core kernel.  Same math and the same FP16 hi/lo pointwise GEMM as k_sepconv_tc (k_conv_tc.cu), but the
// structure that bounded that kernel - a per-thread serial chain of dependent global loads per output row
// (profiles/r01Z_sepconv_tc_timeline.txt: 76 % of a tile; ncu at 16 384 streams: 43 % long-scoreboard stalls, 1.7 TB/s) -
// is gone:
//
//   * inputs arrive by TMA.  A unit of 32 consecutive output rows of a problem needs a CONTIGUOUS run of input rows
//     ((b, f) rows are row-major and Fin = stride * Fout or Fout = up * Fin), so one elected thread fetches it with
//     cp.async.bulk.tensor.2d (two 128-byte-wide boxes per source, SWIZZLE_128B, out-of-range rows zero-filled by the
//     hardware) into a ring of 2..8 stages (56 KB cut to the problem's unit size), that many units ahead of the math; the df_conv0 problem fetches its stream's 2.3 KB
//     feature ring with one 1-D bulk copy.  Completion is an mbarrier transaction count - no thread ever waits on a
//     global load.
//   * the depthwise / grouped-3x3 stage reads shared memory only (the 128-byte swizzle makes the "8 rows x one 16-byte
//     column" access of a quarter warp conflict free) and writes the FP16 hi/lo operand images as before.
//   * the result leaves by TMA too: bias + ReLU out of TMEM into a swizzled FP32 staging tile, then two
//     cp.async.bulk.tensor stores (rows past the end are clipped by the hardware); df_conv0 additionally stores its
//     32-row blocks into the c0 ring slot of their stream.
//   * CTAs are persistent (2 per SM, a contiguous range of tiles each): the 16 KB weight slab, the barriers and the
//     TMEM allocation are set up once, and the input ring keeps running across tile boundaries.
//   * warp specialised: warp 8 is the producer (one lane waits for a stage to be released - an mbarrier the eight compute
//     warps arrive on - and issues the next loads, including the dependent slot look-ups of the df_conv0 problem), so the
//     compute warps never execute a global load and synchronise among themselves (named barrier) once per tile only.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>

#include "engine.h"
#include "tc_common.cuh"

namespace dpdf {

namespace {

using namespace tc;

constexpr int ST_NT = 256;                        // compute threads (warps 0..7); warp 8 is the TMA producer
constexpr int ST_NTB = ST_NT + 32;
constexpr int ST_UNIT = 32;                       // output rows per input stage
constexpr int ST_TILE = 128;                      // output rows per MMA tile
constexpr int ST_IMG = ST_TILE * 64 * 2;          // one FP16 [128][64] operand image (16 KB)
constexpr int ST_RING = 57344;                    // bytes of the input ring: cut into as many stages as the problem's unit size allows
constexpr int ST_MAXSTAGE = 8;
constexpr int ST_RING16 = 24576;                  // df_conv0 uses 8 x 3 KB of the ring: the FP16 c0-ring staging tile (16 KB) sits behind them
constexpr int ST_OFF_A = 0;                       // hi | lo images (32 KB); later the swizzled FP32 output staging
constexpr int ST_OFF_W = ST_OFF_A + 2 * ST_IMG;   // weight slab hi | lo (16 KB)
constexpr int ST_OFF_STAGE = ST_OFF_W + 2 * 64 * 64 * 2;
constexpr int ST_OFF_ZERO = ST_OFF_STAGE + ST_RING;      // 4 KB of zeros (ring blocks of WARMUP streams)
constexpr int ST_OFF_B = ST_OFF_ZERO + 4096;
constexpr int ST_OFF_BAR = ST_OFF_B + 256;
constexpr size_t ST_SMEM = ST_OFF_BAR + 512;
constexpr int ST_MAXP = 2;

struct alignas(64) TmaProblem {
  CUtensorMap in1, in2, out, ring;   // [rows][64] FP32 views: inputs box {32, n_in}, out box {32, 128}, c0 ring box {32, 32}
  const float *pa, *pb, *dw, *tc_pw, *bias;
  int mode, Fin, Fout, stride, up, has_in2;
  int n_in;                          // rows of an input box
  int half_off;                      // byte offset of the second channel-half box of a source inside a stage
  int src1_off;                      // byte offset of the pathway source inside a stage
  int stage_bytes, n_stage;          // the ring of this problem: n_stage stages of stage_bytes
  int tile0;                         // first tile of this problem in the launch
  int nrows;                         // B * Fout (< 2^31: checked by the launcher)
  unsigned magic_fout, magic_up;     // ceil(2^32 / d): x / d == __umulhi(x, magic) for the small x used here (x < 2^16, d < 2^10)
};

struct SepTmaParams {
  TmaProblem prob[ST_MAXP];
  const IoDesc* io;
  State st;
  int nprob, B, tiles, per_cta;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(smem_u32(src))
               : "memory");
}

}  // namespace

#define CSYNC() asm volatile("bar.sync 1, %0;" ::"n"(ST_NT) : "memory")      /* the eight compute warps only */

__global__ void __launch_bounds__(ST_NTB, 2) k_sepconv_tma(const __grid_constant__ SepTmaParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* Aimg = smem_raw + ST_OFF_A;
  unsigned char* Wsm = smem_raw + ST_OFF_W;
  unsigned char* stage = smem_raw + ST_OFF_STAGE;
  unsigned char* zeros = smem_raw + ST_OFF_ZERO;
  float* bs = reinterpret_cast<float*>(smem_raw + ST_OFF_B);
  // [0] weights landed, [1] accumulator ready, [2 + 16 seg + s] stage s full, [2 + 16 seg + 8 + s] stage s released by the compute warps
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + ST_OFF_BAR);
  uint64_t* seg_done = bars + 2 + ST_MAXP * 2 * ST_MAXSTAGE;        // [seg] the compute warps have left segment seg (its ring is free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(seg_done + ST_MAXP);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t_begin = blockIdx.x * p.per_cta, t_end = min(p.tiles, t_begin + p.per_cta);
  if (t_begin >= t_end) return;

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    for (int i = 0; i < ST_MAXP * 2 * ST_MAXSTAGE; ++i) mbar_init(bars + 2 + i, (i & ST_MAXSTAGE) ? ST_NT / 32 : 1);
    for (int i = 0; i < ST_MAXP; ++i) mbar_init(seg_done + i, ST_NT / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  for (int i = tid; i < 4096 / 16; i += ST_NTB) reinterpret_cast<uint4*>(zeros)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                   // everything above overlaps with the tail of the previous kernel

  const int g = (tid >> 3) & 15, c = g * 4;                    // channel quad of this thread in the prologue
  const int rs = (tid & 7) | ((tid >> 7) << 3);                // its row inside a 16-row group
  const int half = g >> 3, chunk = g & 7;                      // channel half (TMA box) and 16-byte chunk inside a 128-byte row
  uint32_t ovf = 0;                                            // FP16 range guard of the operand converter (tc_common.cuh)
  uint32_t w_loads = 0, mma_phase = 0;                          // weight slabs loaded / tiles multiplied so far (barrier parities)
  float4 wt[9], pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;

  // A CTA's tile range touches each problem of the launch in at most one contiguous segment; every segment runs its own
  // input pipeline (own barriers, ring cut to the problem's unit size) and is drained before the next one starts.
  int prev_seg = -1;                                            // the segment this CTA ran before the current one
  for (int pi = 0; pi < p.nprob; ++pi) {
    const TmaProblem& q = p.prob[pi];
    const int s_begin = max(t_begin, q.tile0);
    const int s_end = min(t_end, pi + 1 < p.nprob ? p.prob[pi + 1].tile0 : p.tiles);
    if (s_begin >= s_end) continue;
    const int n_units = (s_end - s_begin) * (ST_TILE / ST_UNIT);
    const int NS = q.n_stage;
    uint64_t* full = bars + 2 + pi * 2 * ST_MAXSTAGE;
    uint64_t* empty = full + ST_MAXSTAGE;

    // ---- producer (thread 0): the inputs of unit `lu` of this segment into stage lu % NS -------------------------------
    auto issue = [&](int lu) {
      const int r0 = (s_begin - q.tile0 + lu / 4) * ST_TILE + (lu & 3) * ST_UNIT;
      const int sidx = lu % NS;
      unsigned char* dst = stage + sidx * q.stage_bytes;
      uint64_t* bar = full + sidx;
      if (q.mode == 0) {
        // output rows [r0, r0 + 32) read input rows [first, first + n_in): centre of row R is R * stride (up == 1) or R / up
        const int first = (q.up > 1 ? r0 / q.up : r0 * q.stride) - 1;
        mbar_expect_tx(bar, (uint32_t)q.n_in * 256u * (q.has_in2 ? 2u : 1u));
        tma_load_2d(dst, &q.in1, 0, first, bar);
        tma_load_2d(dst + q.half_off, &q.in1, 32, first, bar);
        if (q.has_in2) {
          tma_load_2d(dst + q.src1_off, &q.in2, 0, first, bar);
          tma_load_2d(dst + q.src1_off + q.half_off, &q.in2, 32, first, bar);
        }
      } else {
        // df_conv0: the three-frame feature ring of the unit's stream (a unit never straddles streams: 96 % 32 == 0);
        // past the end any valid source will do, those rows are zeroed below
        const int b = r0 < q.nrows ? r0 / NDF : 0;
        const int slot = io_slot(p.io, b);
        *reinterpret_cast<int*>(dst + 3 * 2 * NDF * 4) = p.st.pos[slot];     // ring head for the consumers (released by the arrive below)
        mbar_expect_tx(bar, 3 * 2 * NDF * 4);
        bulk_g2s(dst, p.st.df_ring + (size_t)slot * 3 * 2 * NDF, 3 * 2 * NDF * 4, bar);
      }
    };
    if (warp == ST_NT / 32) {                                    // ---- producer warp: runs ahead of the compute warps by up to NS units
      if (elect_one()) {
        if (prev_seg >= 0) mbar_wait(seg_done + prev_seg, 0);      // the ring is re-cut: wait until the previous segment has drained
        for (int lu = 0; lu < n_units; ++lu) {
          if (lu >= NS) mbar_wait(empty + lu % NS, (lu / NS - 1) & 1);
          issue(lu);
        }
      }
      prev_seg = pi;
      continue;
    }
    ++w_loads;
    if (tid == 0) {                                              // all MMAs that read the previous slab have completed
      mbar_expect_tx(bars, 2 * 64 * 64 * 2);
      bulk_g2s(Wsm, q.tc_pw, 2 * 64 * 64 * 2, bars);
    }
    {
      const int nw = q.mode == 0 ? 3 * q.up : 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) wt[t] = t < nw ? __ldg(reinterpret_cast<const float4*>(q.dw + (size_t)t * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      pa = pb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q.has_in2) {
        pa = __ldg(reinterpret_cast<const float4*>(q.pa + c));
        pb = __ldg(reinterpret_cast<const float4*>(q.pb + c));
      }
      if (tid < 64) bs[tid] = __ldg(q.bias + tid);               // read in the epilogue, after several barriers
    }

    for (int lu = 0; lu < n_units; ++lu) {
      const int tile = s_begin + lu / 4, uu = lu & 3;
      const int r0 = (tile - q.tile0) * ST_TILE + uu * ST_UNIT;
      const int sidx = lu % NS;
      const unsigned char* src = stage + sidx * q.stage_bytes;
      // per-unit index arithmetic (one 32-bit division each; the per-row part below is multiply-high only)
      const unsigned fo0 = (unsigned)r0 % (unsigned)q.Fout;      // frequency position of the unit's first row
      const unsigned a0 = q.up > 1 ? (unsigned)r0 / (unsigned)q.up : 0u;
      const unsigned rem0 = (unsigned)r0 - a0 * (unsigned)q.up;  // r0 % up
      mbar_wait(full + sidx, (lu / NS) & 1);

      // ---- depthwise / grouped stage of two rows per thread, out of shared memory -------------------------------------
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int j = rs + 16 * k;                               // row inside the unit
        const int R = r0 + j;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (R < q.nrows) {
          if (q.mode == 0) {
            const unsigned x = fo0 + (unsigned)j;
            const int fo = (int)(x - (unsigned)q.Fout * __umulhi(x, q.magic_fout));   // (fo0 + j) % Fout
            int fc, jj, lr;                                      // input position of the centre tap, sub-pixel phase, its row in the box
            if (q.up > 1) {
              fc = (int)__umulhi((unsigned)fo, q.magic_up); jj = fo - fc * q.up;
              lr = (int)__umulhi(rem0 + (unsigned)j, q.magic_up) + 1;                 // R / up - r0 / up + 1
            } else {
              fc = fo * q.stride; jj = 0;
              lr = j * q.stride + 1;
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int fi = fc + t - 1;
              if (fi < 0 || fi >= q.Fin) continue;               // taps beyond the stream's own rows (the box holds a neighbour there)
              const int l = lr + t - 1;
              const int off = half * q.half_off + l * 128 + ((chunk ^ (l & 7)) << 4);
              float4 v = *reinterpret_cast<const float4*>(src + off);
              if (q.has_in2) {
                const float4 u = *reinterpret_cast<const float4*>(src + q.src1_off + off);
                v.x += fmaxf(fmaf(u.x, pa.x, pb.x), 0.f);
                v.y += fmaxf(fmaf(u.y, pa.y, pb.y), 0.f);
                v.z += fmaxf(fmaf(u.z, pa.z, pb.z), 0.f);
                v.w += fmaxf(fmaf(u.w, pa.w, pb.w), 0.f);
              }
              const float4 w = jj == 0 ? wt[t] : (jj == 1 ? wt[3 + t] : wt[6 + t]);
              acc.x = fmaf(w.x, v.x, acc.x);
              acc.y = fmaf(w.y, v.y, acc.y);
              acc.z = fmaf(w.z, v.z, acc.z);
              acc.w = fmaf(w.w, v.w, acc.w);
            }
          } else {
            const int fo = (int)fo0 + j;                         // a unit never straddles streams (96 % 32 == 0)
            const float* ring = reinterpret_cast<const float*>(src);
            const int pos = *reinterpret_cast<const int*>(src + 3 * 2 * NDF * 4);
            const int plane = c >= 32 ? 1 : 0;
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              int ph = pos + 1 + kt;                             // (pos + 1 + kt) % 3 with pos in [0, 15)
              ph -= 3 * ((ph * 11) >> 5);
              const float* rowp = ring + (ph * 2 + plane) * NDF;
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
        uint2 h, l2;
        split2_f16(acc.x, acc.y, h.x, l2.x);
        split2_f16(acc.z, acc.w, h.y, l2.y);
        ovf |= f16_nonfinite(h.x) | f16_nonfinite(h.y);
        const int ar = uu * ST_UNIT + j;                         // row of the MMA tile
        unsigned char* dst = Aimg + (ar >> 3) * 1024 + (g >> 1) * 128 + (ar & 7) * 16 + (g & 1) * 8;
        *reinterpret_cast<uint2*>(dst) = h;
        *reinterpret_cast<uint2*>(dst + ST_IMG) = l2;
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + sidx)) : "memory");   // stage released
      if (uu != 3) continue;
      fence_async_smem();                                        // operand images -> async proxy (tensor core)
      tc_fence_before();
      CSYNC();                                                   // the images of the tile are complete

      // ---- pointwise GEMM of the tile: 12 tcgen05.mma (hi*hi + lo*hi + hi*lo per 16-wide k step) -------------------
      tc_fence_after();
      if (warp == 0 && elect_one()) {                            // elect.sync, not `tid == 0`: tc_common.cuh:elect_one
        mbar_wait(bars, (w_loads - 1) & 1);                      // the slab of this problem has landed (returns at once later on)
        const uint32_t ah = smem_u32(Aimg), al = ah + ST_IMG, bh = smem_u32(Wsm), bl = bh + 64 * 64 * 2;
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
      if (warp == 0) mbar_wait(bars + 1, mma_phase & 1);         // one warp polls, the others park on the named barrier
      CSYNC();
      ++mma_phase;
      tc_fence_after();

      // ---- epilogue: thread = (row = TMEM lane, 32 columns); bias + ReLU into the 128B-swizzled FP32 staging tile ------
      {
        const int qd = warp & 3, ch = warp >> 2, row = qd * 32 + lane;
        const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16) + ch * 32;
        unsigned char* hb = Aimg + ch * (ST_TILE * 128) + row * 128;      // channel half ch: [128 rows][128 B]
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
          float v[16];
          tmem_ld16(ta + c16 * 16, v);
#pragma unroll
          for (int jq = 0; jq < 4; ++jq) {
            const int col = ch * 32 + c16 * 16 + jq * 4;
            const float4 o = make_float4(fmaxf(v[jq * 4] + bs[col], 0.f), fmaxf(v[jq * 4 + 1] + bs[col + 1], 0.f),
                                         fmaxf(v[jq * 4 + 2] + bs[col + 2], 0.f), fmaxf(v[jq * 4 + 3] + bs[col + 3], 0.f));
            *reinterpret_cast<float4*>(hb + (((c16 * 4 + jq) ^ (row & 7)) << 4)) = o;
            if (q.mode == 1 && p.st.c0_fp16) {                   // FP16 copy of the tile for the c0 ring: [128 rows][64 halves], 128B swizzle
              const int hc = ch * 4 + c16 * 2 + (jq >> 1);         // 16-byte chunk of the 128-byte row
              *reinterpret_cast<uint2*>(stage + ST_RING16 + row * 128 + ((hc ^ (row & 7)) << 4) + (jq & 1) * 8) = pack4_f16(o);
            }
          }
        }
      }
      fence_async_smem();                                        // staging -> async proxy (TMA store)
      tc_fence_before();
      CSYNC();
      if (tid == 0) {
        const int row0 = (tile - q.tile0) * ST_TILE;
        tma_store_2d(&q.out, 0, row0, Aimg);
        tma_store_2d(&q.out, 32, row0, Aimg + ST_TILE * 128);
        if (q.mode == 1) {                                       // the same rows into the c0 ring slot of their stream, 32-row blocks
          for (int k = 0; k < 4; ++k) {
            const int R = row0 + 32 * k;
            if (R >= q.nrows) break;
            const int b = R / NDF, fo = R - b * NDF;
            const int slot = io_slot(p.io, b);
            const int rrow = (slot * ORD + p.st.pos[slot] % ORD) * NDF + fo;
            const bool warm = (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) != 0;
            if (p.st.c0_fp16) {                                  // ring rows are 128 B (64 halves): two per FP32-sized row of the region
              tma_store_2d(&q.ring, 0, (slot * 2 * ORD + p.st.pos[slot] % ORD) * NDF + fo, warm ? zeros : stage + ST_RING16 + k * 4096);
              continue;
            }
            tma_store_2d(&q.ring, 0, rrow, warm ? zeros : Aimg + k * 4096);
            tma_store_2d(&q.ring, 32, rrow, warm ? zeros : Aimg + ST_TILE * 128 + k * 4096);
          }
        }
        bulk_commit();
        bulk_wait_read<0>();                                     // the staging tile may be overwritten by the next tile's images
      }
      CSYNC();
    }
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(seg_done + pi)) : "memory");
  }
  if (warp == ST_NT / 32) return;                                // producer: everything it issued has been consumed
  if (tid == 0) bulk_wait<0>();
  if (ovf) p.io->err[DPDF_ERRW_RANGE] = 1;
  tc_fence_before();
  CSYNC();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

// [rows][64] FP32 row-major view, box {32 floats, box_rows}, 128-byte swizzle, zero fill outside
bool make_map_f16(CUtensorMap* m, const void* base, long long rows, int box_rows) {     // [rows][64] FP16, box {64, box_rows}
  const cuuint64_t dims[2] = {64, (cuuint64_t)std::max<long long>(rows, 1)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map(CUtensorMap* m, const void* base, long long rows, int box_rows) {
  const cuuint64_t dims[2] = {64, (cuuint64_t)std::max<long long>(rows, 1)};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool sepconv_tma_available() { return g_encode != nullptr; }

void launch_sepconv_tma(Engine& e, const SepProblem* probs, int nprob, int B, cudaStream_t st) {
  SepTmaParams p{};
  p.io = e.io_dev;
  p.st = e.st;
  p.nprob = nprob;
  p.B = B;
  int tiles = 0;
  bool ok = nprob <= ST_MAXP;
  for (int i = 0; i < nprob && ok; ++i) {
    const SepProblem& s = probs[i];
    TmaProblem& q = p.prob[i];
    q.pa = s.pa; q.pb = s.pb; q.dw = s.dw; q.tc_pw = s.tc_pw; q.bias = s.bias;
    q.mode = s.mode; q.Fin = s.Fin; q.Fout = s.Fout; q.stride = s.stride; q.up = s.up; q.has_in2 = s.in2 ? 1 : 0;
    ok = ok && (long long)B * std::max(s.Fout, s.Fin) < (1ll << 31) - 4 * ST_TILE;
    q.nrows = (int)((long long)B * s.Fout);
    q.magic_fout = (unsigned)(((1ull << 32) + s.Fout - 1) / s.Fout);
    q.magic_up = (unsigned)(((1ull << 32) + s.up - 1) / std::max(s.up, 1));
    q.tile0 = tiles;
    tiles += (q.nrows + ST_TILE - 1) / ST_TILE;
    ok = ok && make_map(&q.out, s.out, q.nrows, ST_TILE);
    if (s.mode == 0) {
      q.n_in = s.up > 1 ? (ST_UNIT - 1) / s.up + 4 : (ST_UNIT - 1) * s.stride + 3;
      q.half_off = (q.n_in * 128 + 1023) / 1024 * 1024;
      q.src1_off = 2 * q.half_off;
      q.stage_bytes = (q.has_in2 ? 4 : 2) * q.half_off;
      q.n_stage = std::min(ST_MAXSTAGE, ST_RING / q.stage_bytes);
      ok = ok && q.n_stage >= 2 && q.n_in <= 256;
      ok = ok && make_map(&q.in1, s.in1, (long long)B * s.Fin, q.n_in);
      if (s.in2) ok = ok && make_map(&q.in2, s.in2, (long long)B * s.Fin, q.n_in);
    } else {
      q.stage_bytes = 3072;                      // 3 x 2 x 96 floats of ring, rounded up
      q.n_stage = ST_MAXSTAGE;
      ok = ok && s.Fout == NDF &&
           (e.st.c0_fp16 ? make_map_f16(&q.ring, e.st.c0_ring, (long long)e.max_streams * 2 * ORD * NDF, 32)
                         : make_map(&q.ring, e.st.c0_ring, (long long)e.max_streams * ORD * NDF, 32));
    }
  }
  if (!ok) {                                     // geometry this kernel does not cover: the per-tile kernel handles everything
    launch_sepconv_tc(e, probs, nprob, B, st);
    return;
  }
  p.tiles = tiles;
  const int grid = std::min(tiles, 2 * e.num_sms);
  p.per_cta = (tiles + grid - 1) / grid;
  launch_k(e, k_sepconv_tma, dim3((tiles + p.per_cta - 1) / p.per_cta), dim3(ST_NTB), ST_SMEM, st, p);
}

void init_conv_tma_kernels() {
  cudaFuncSetAttribute(k_sepconv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (getenv("DPDF_DEBUG")) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sepconv_tma, ST_NTB, ST_SMEM);
    fprintf(stderr, "[dpdf] k_sepconv_tma: %zu B smem, %d CTAs/SM, TMA encode %s\n", ST_SMEM, n, g_encode ? "ok" : "missing");
  }
}

}  // namespace dpdf
