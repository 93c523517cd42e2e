// k_dprnn_intra_tc: the bidirectional intra-frame GRU of a DPRNN block (layers.py:126-132, 176-177) on the
// 5th-generation tensor cores.
//
// One CTA = (branch, direction, tile of 128 streams) and walks the F' frequency positions sequentially.  Per step
//     P[128 x 256] = x_t * W_ih^T  (N = 192: r | z | in)   +   h_{t-1} * W_hh^T  (r, z on top; hn in its own 64 columns)
// is a chain of tcgen05.mma.kind::f16 instructions with the FP32 accumulator in tensor memory.  FP32 accuracy is
// kept by the error-compensated split x = hi + lo (both FP16, 11-bit significands; three products hi*hi + lo*hi
// + hi*lo; weights.py:fp16_split): the same accuracy as 3xTF32 at half the shared-memory footprint and twice the
// tensor rate, which is what lets both 192x64 weight matrices of a direction stay resident in shared memory
// (96 KB) next to the operand images of x_t and x_{t+1}.  h_t never touches shared memory: the gate warps write
// its FP16 hi/lo halves with tcgen05.st into tensor memory, from where the recurrent product reads its A operand
// (SS-mode MMAs re-read the 4 KB A tile from shared memory for every instruction, which made the N = 64 / 128
// recurrent MMAs shared-memory-bandwidth bound: tools/ubench/intra_tc_timeline.cu).
//
// Pipeline per step t (the r|z|in accumulators are double buffered in TMEM: 2 x 192 + 64 (hn) + 64 (h operand) columns):
//   warp 16      : h-part MMAs of step t (critical path) -> commit -> x-part MMAs of step t+1 into the other buffer
//   warps 0..15  : wait commit, tcgen05.ld the four gate pre-activations of (row = TMEM lane, 16 units), gate
//                  math in registers (h_{t-1} never leaves registers), write h_t as FP16 hi/lo operand rows for
//                  the next step and as FP32 to hcat[b][f][dir*64 + u]; convert the prefetched x_{t+2} tile.
// The x-part MMAs and the global loads of x are hidden behind the gate math of the previous step, and the recurrent
// MMAs of step t+1 are issued per K slice (16 units) as soon as the gate warps have produced that slice of h_t.
//
// Row duplication (template parameter D = 1, 2, 4).  A step of the sweep is bound by the gate math of its 128 x 64
// (stream, unit) pairs, not by the tensor core, and at latency batch sizes most SMs are idle (32 CTAs per 1024
// streams).  With D > 1 a CTA owns only 128 / D streams and every stream occupies D rows of the M = 128 tile: the
// MMAs are unchanged (an M = 64 instruction would cost the same tensor time, B300_MICROARCH "tcgen05 floor"), rows
// (s, 0..D-1) carry identical operands and therefore identical accumulators, and the thread of row (s, part) does the
// gate math of only 4 / D of the cg-group's units per K slice.  The D rows of a stream sit in the same warp (lanes
// l + part * 32 / D), so the partners exchange their h units with shuffles before every thread writes the complete
// operand columns of its own row with tcgen05.st.  Per-step gate work per CTA drops by D and the sweep spreads over
// D times as many SMs (128 CTAs at 1024 streams with D = 4).
#include "engine.h"
#include "tc_common.cuh"

#ifndef ITC_POLY
#define ITC_POLY 0          // 1: r, z exponentials on the FMA pipe instead of MUFU (measured: no gain, the gate phase is issue/latency bound)
#endif
#ifndef ITC_XEARLY
#define ITC_XEARLY 0        // convert x_{t+2} inside the gate loop instead of after it
#endif
namespace dpdf {

namespace {

using namespace tc;

constexpr int ITC_GATE = 512;               // 16 gate warps: warp w -> TMEM lane quadrant w & 3, 16-unit group w >> 2
constexpr int ITC_NT = ITC_GATE + 32;       // + warp 16: the MMA issuer (does nothing else, so its descriptors stay in uniform registers)
constexpr int W_IMG = 192 * 64 * 2;         // bytes of one FP16 [192][64] weight image
constexpr int A_IMG = 128 * 64 * 2;         // bytes of one FP16 [128][64] activation image
constexpr int OFF_X = 4 * W_IMG;            // x images: [2 buffers][hi | lo]
constexpr int OFF_ST = OFF_X + 4 * A_IMG;   // h_t staging for the coalesced write-out: [2 buffers][128 rows][16 chunks of 16 B],
constexpr int ST_BUF = 128 * 256;           // chunk index XOR-swizzled with the row so that both the per-row writes of the gate
constexpr int OFF_BIAS = OFF_ST + 2 * ST_BUF;  // warps and the row-contiguous reads of the write-out are conflict-free
constexpr int OFF_BAR = OFF_BIAS + 1024;    // two mbarriers + TMEM base slot
constexpr size_t INTRA_TC_SMEM = OFF_BAR + 64;
// tensor-memory columns: P[2] = (r | z | in) double buffered, hn single (written and drained inside one step),
// h_{t} as the FP16 hi / lo A operand of the recurrent product (two K halves per 32-bit column)
constexpr uint32_t TM_P = 0, TM_HN = 384, TM_HHI = 448, TM_HLO = 480;

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#if ITC_POLY
// 2^x for a pair of values on the FMA pipe (the gate math is MUFU bound: 16 lanes/clk/SM): round-to-nearest split
// x = n + f with the 1.5*2^23 trick, degree-5 minimax polynomial for 2^f on [-0.5, 0.5] (2.3e-7 max relative error
// in FP32 Horner form, the same as ex2.approx), 2^n by adding n to the exponent field.  Needs -125 <= x <= 126.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
  const float2 y = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(y, nmagic);
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 q = __ffma2_rn(make_float2(0.0013276409590616822f, 0.0013276409590616822f), f, make_float2(0.00967552699148655f, 0.00967552699148655f));
  q = __ffma2_rn(q, f, make_float2(0.05550713464617729f, 0.05550713464617729f));
  q = __ffma2_rn(q, f, make_float2(0.24022120237350464f, 0.24022120237350464f));
  q = __ffma2_rn(q, f, make_float2(0.6931469440460205f, 0.6931469440460205f));
  q = __ffma2_rn(q, f, make_float2(1.0000001192092896f, 1.0000001192092896f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(y.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(y.y) << 23)));
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
#endif
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace

struct IntraTcParams {
  const float* x[2];      // [B][Fp][64]    (index 0 = df branch, 1 = erb branch)
  float* hcat[2];         // [B][Fp][128]
  int Fp[2];
  const float* wimg[2];   // [2 dirs][W_ih hi | W_ih lo | W_hh hi | W_hh lo] FP16 operand images
  const float* wimg_f[2]; // fragment form: the same with the K axis of W_hh in the order of intra_sweep_f (weights.py: tc.intra_f)
  const float* bias[2];   // [2][4][64], exponent scales folded in (weights.py: tc.intra_bias)
  int tiles[2];           // stream tiles of the sweep per branch: ceil(B / (128 / D)) with the branch's row duplication D
  int B;
  int* progress;          // [df: 2 dirs x tiles[0] | erb: 2 dirs x tiles[1]] completed steps of each CTA, or nullptr (overlapped post kernel)
  int* err;               // engine error words (IoDesc::err), may be nullptr
#ifdef ITC_TIMELINE
  long long* tl;          // [steps][12] SM-clock stamps of CTA 0 (tools/ubench/intra_tc_timeline.cu)
#endif
};

#ifdef ITC_TIMELINE
// stamp AFTER everything issued so far has completed as far as this warp can tell: the clock read depends on a
// volatile shared-memory load, which cannot issue before a pending (deferred-blocking) barrier has released
#define TL(slot) do { if (blockIdx.x == 0 && lane == 0 && (warp == 16 || warp == 5)) { \
    unsigned v_; long long c_; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v_) : "r"(smem_u32(tmem_slot))); \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) : "r"(v_)); p.tl[t * 12 + (slot)] = c_; } } while (0)
#else
#define TL(slot) do { } while (0)
#endif

// The sweep of one CTA: branch br, direction dir, stream tile `tile` of 128 / D streams.
template <int D, bool SR = false>
__device__ __forceinline__ void intra_sweep(const IntraTcParams& p, const int br, const int dir, const int tile) {
  static_assert(D == 1 || D == 2 || D == 4, "row duplication factor");
  static_assert(!SR || D > 1, "split rows need at least two rows per stream");
  constexpr int SPC = 128 / D;        // streams per CTA
  constexpr int LPQ = 32 / D;         // lanes of a warp (TMEM lane quadrant) that hold distinct streams
  constexpr int UPS = 4 / D;          // units per gate thread and K slice
  constexpr int NX = D == 1 ? 2 : 1;  // x staging items per thread
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* Wsm = smem_raw;
  unsigned char* Xsm = smem_raw + OFF_X;
  unsigned char* Ssm = smem_raw + OFF_ST;
  float* sb = reinterpret_cast<float*>(smem_raw + OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + OFF_BAR);       // [0] weights landed, [1] step accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3, cg = warp >> 2;
  const int part = lane / LPQ;                               // which of the D rows of its stream this thread is
  const int srow = qd * LPQ + (lane % LPQ);                  // the stream's row in the CTA's staging tiles
  const bool lo_row = SR && part >= D / 2;                   // split rows: the upper half of a stream's D rows carries the lo halves
  const int T = br ? p.Fp[1] : p.Fp[0];
  const int b0 = tile * SPC;
  const float* __restrict__ xg = br ? p.x[1] : p.x[0];
  float* __restrict__ hg = br ? p.hcat[1] : p.hcat[0];

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid < 256) sb[tid] = __ldg((br ? p.bias[1] : p.bias[0]) + dir * 4 * C + tid);
  for (int i = tid; i < ST_BUF / 16; i += ITC_NT) reinterpret_cast<uint4*>(Ssm + ST_BUF)[i] = make_uint4(0u, 0u, 0u, 0u);   // h_{-1} = 0 (read back as h_prev of step 0)
  // progress counter of this CTA (recomputed where it is used: it must not cost the gate warps a live register)
#ifdef ITC_NO_PROGRESS
  auto progress_ptr = [&]() -> int* { return nullptr; };
#else
  auto progress_ptr = [&]() -> int* { return p.progress ? p.progress + (br ? 2 * p.tiles[0] + dir * p.tiles[1] : dir * p.tiles[0]) + tile : nullptr; };
#endif
  __syncthreads();                                           // barriers initialised
  if (tid == 0) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(br ? p.wimg[1] : p.wimg[0]) + (size_t)dir * 4 * W_IMG;
    mbar_expect_tx(bars, 4 * W_IMG);
#pragma unroll
    for (int i = 0; i < 4; ++i) bulk_g2s(Wsm + i * W_IMG, src + (size_t)i * W_IMG, W_IMG, bars);
  }

  // ---- x tile staging: SPC streams x 8 chunks of 8 floats, NX (stream, chunk) items per thread ---------------
  // warp item j = warp + 16 i covers stream group j >> 1, K half j & 1: lanes = 8 streams x 4 chunks, so the 16-byte
  // operand rows a warp writes are four contiguous 128-byte core matrices (conflict-free); every item is stored to
  // the D operand rows of its stream (D = 4: 256 items, the upper eight warps have none)
  const bool x_active = D < 4 || warp < 8;
  const int xr_[2] = {((warp) >> 1) * 8 + (lane & 7), ((warp + 16) >> 1) * 8 + (lane & 7)};
  const int xkc = (warp & 1) * 4 + (lane >> 3);
  auto load_x = [&](int t, float (&v)[NX][8]) {
    const int f = dir ? T - 1 - t : t;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int bb = b0 + xr_[i];
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
      if (bb < p.B && x_active) {
        const float4* src = reinterpret_cast<const float4*>(xg + ((size_t)bb * T + f) * C + xkc * 8);
        a = __ldg(src);
        c = __ldg(src + 1);
      }
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
      v[i][4] = c.x; v[i][5] = c.y; v[i][6] = c.z; v[i][7] = c.w;
    }
  };
  auto store_x1 = [&](int buf, const float (&v)[NX][8], int i) {
    if (!x_active) return;
    uint4 hi, lo;
    split8_f16(v[i], hi, lo);
    // FP16 range guard (tc_common.cuh:f16_nonfinite); stored at once: a flag word would be a live register of the gate warps
    if ((f16_nonfinite(hi.x) | f16_nonfinite(hi.y) | f16_nonfinite(hi.z) | f16_nonfinite(hi.w)) && p.err) p.err[DPDF_ERRW_RANGE] = 1;
#pragma unroll
    for (int pt = 0; pt < D; ++pt) {                         // operand row of (stream, part): quadrant, part block, lane
      const int r = (xr_[i] / LPQ) * 32 + pt * LPQ + (xr_[i] % LPQ);
      unsigned char* dst = Xsm + buf * 2 * A_IMG + img16_off(r, xkc);
      if constexpr (SR) {                                    // one image: hi in the lower half of the stream's rows, lo in the upper
        *reinterpret_cast<uint4*>(dst) = pt < D / 2 ? hi : lo;
      } else {
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + A_IMG) = lo;
      }
    }
  };
  auto store_x = [&](int buf, const float (&v)[NX][8]) {
#pragma unroll
    for (int i = 0; i < NX; ++i) store_x1(buf, v, i);
  };
  // Everything above (barriers, TMEM, the 96 KB of weight images) overlaps with the tail of the previous kernel.
  pdl_wait();
  if (tid == 0 && progress_ptr()) {
    *reinterpret_cast<volatile int*>(progress_ptr()) = 0;    // the previous consumer of these counters has completed
    __threadfence();
  }
  __syncthreads();
  pdl_trigger();                                             // every CTA of this grid is resident (or done): dependents may launch
  float xv[NX][8];
  if (warp < 16) {
    load_x(0, xv);
    store_x(0, xv);
    if (T > 1) {
      load_x(1, xv);
      store_x(1, xv);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 16) {                                           // h_0 = 0 in the TMEM operand columns
    const uint32_t a = tmem + ((uint32_t)(qd * 32) << 16) + cg * 8;
    tmem_st4(a + TM_HHI, 0u, 0u, 0u, 0u); tmem_st4(a + TM_HHI + 4, 0u, 0u, 0u, 0u);
    tmem_st4(a + TM_HLO, 0u, 0u, 0u, 0u); tmem_st4(a + TM_HLO + 4, 0u, 0u, 0u, 0u);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- MMA issue (warp 16, one lane elected with elect.sync: with `lane == 0` every tcgen05.mma sat in an elect /
  // broadcast loop of its own, ~100 cycles per issued instruction on the critical path of a step) -----------------
  // The recurrent product of step t+1 is issued K-slice by K-slice: slice ks only needs units [16 ks, 16 ks + 16) of
  // h_t, and the gate warps produce the units in exactly that order (every thread owns 4 units of each slice), so the
  // tensor core works on slice ks while the gate math of slices ks+1.. is still running.  Only the last slice's six
  // MMAs and the commit remain on the critical path of a step.
  if (warp == 16) {
    int* my_progress = progress_ptr();
    const uint32_t w_base = smem_u32(Wsm), x_base = smem_u32(Xsm);
    // The TMEM base comes out of shared memory, i.e. out of a per-thread register: unless the compiler can see that it is
    // warp-uniform it wraps EVERY tcgen05.mma in an elect / broadcast loop (~100 cycles per issued MMA on the critical
    // path of the step).  A shuffle from lane 0 is uniform by construction.
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    constexpr uint64_t DESC0 = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46);
    auto x_mma = [&](int t) {                                // P[t & 1][0, 192) = x_t * W_ih^T
      const uint32_t xa = x_base + (t & 1) * 2 * A_IMG, d = tmem + TM_P + (t & 1) * 192;
      const uint64_t dah = DESC0 | (xa >> 4), dal = DESC0 | ((xa + A_IMG) >> 4);
      const uint64_t dbh = DESC0 | (w_base >> 4), dbl = DESC0 | ((w_base + W_IMG) >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                       // K = 64 in steps of 16 halves = two core matrices = 256 B
        umma_f16(d, dah + ks * 16, dbh + ks * 16, idesc_f16(128, 192), ks > 0);
        if constexpr (!SR) umma_f16(d, dal + ks * 16, dbh + ks * 16, idesc_f16(128, 192), 1);
        umma_f16(d, dah + ks * 16, dbl + ks * 16, idesc_f16(128, 192), 1);
      }
    };
    // K slice ks of the recurrent part: A = h (hi, lo) straight from tensor memory (16 halves = 8 columns), so the
    // only shared-memory traffic is W_hh
    auto h_mma_slice = [&](int t, int ks) {
      const uint32_t whi = w_base + 2 * W_IMG + ks * 256, wlo = w_base + 3 * W_IMG + ks * 256;
      const uint64_t rz_h = DESC0 | (whi >> 4), rz_l = DESC0 | (wlo >> 4);
      const uint64_t n_h = DESC0 | ((whi + 16 * 1024) >> 4), n_l = DESC0 | ((wlo + 16 * 1024) >> 4);
      const uint32_t ah = tmem + TM_HHI + ks * 8, al = tmem + TM_HLO + ks * 8;
      const uint32_t drz = tmem + TM_P + (t & 1) * 192, dn = tmem + TM_HN;
      umma_f16_ts(drz, ah, rz_h, idesc_f16(128, 128), 1);    // r, z += h * W_hh[r,z]^T (on top of the x part)
      if constexpr (!SR) umma_f16_ts(drz, al, rz_h, idesc_f16(128, 128), 1);
      umma_f16_ts(drz, ah, rz_l, idesc_f16(128, 128), 1);
      umma_f16_ts(dn, ah, n_h, idesc_f16(128, 64), ks > 0);  // hn = h * W_hh[n]^T
      if constexpr (!SR) umma_f16_ts(dn, al, n_h, idesc_f16(128, 64), 1);
      umma_f16_ts(dn, ah, n_l, idesc_f16(128, 64), 1);
    };
    if (elect_one()) {
      mbar_wait(bars, 0);                                    // weight images landed (async proxy -> async proxy)
      x_mma(0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) h_mma_slice(0, ks);
      umma_commit(bars + 1);
      if (T > 1) x_mma(1);
    }
    for (int t = 0; t + 1 < T; ++t) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        // units of slice ks of h_t are in TMEM (ks == 0: every thread has also drained hn of step t;
        // ks == 3: P[t & 1] drained and x_{t+2} staged)
        if (ks == 0) asm volatile("bar.sync 1, %0;" ::"n"(ITC_NT) : "memory");
        if (ks == 1) asm volatile("bar.sync 2, %0;" ::"n"(ITC_NT) : "memory");
        if (ks == 2) asm volatile("bar.sync 3, %0;" ::"n"(ITC_NT) : "memory");
        if (ks == 3) asm volatile("bar.sync 4, %0;" ::"n"(ITC_NT) : "memory");
        if (ks == 3) TL(5);
        if (elect_one()) {
          tc_fence_after();
          h_mma_slice(t + 1, ks);
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(bars + 1);
        TL(6);
#ifndef ITC_NO_X
        if (t + 2 < T) x_mma(t + 2);
#endif
        TL(7);
        if (my_progress) {
          // every gate warp issued the write-out of h_{t-1} before it handed over slice 1 of step t (named barriers are
          // CTA-scope synchronisation; the fence below makes those stores visible device-wide before the count)
          __threadfence();
          *reinterpret_cast<volatile int*>(my_progress) = t;
        }
      }
      __syncwarp();
    }
  } else {
    // ---- the sweep (gate warps) -----------------------------------------------------------------------------
    // thread (row = TMEM lane, g = warp >> 2) owns units 16 ks + 4 g + j  (ks, j = 0..3): four of every K slice
    const uint32_t lane_base = tmem + ((uint32_t)(qd * 32) << 16);
    // Write-out of h_s: staging[s & 1] -> hcat[b][f][dir*64 ..], two full 256-byte rows per warp instruction.  It runs
    // inside step s + 1, under the MUFU-bound gate math: passing the accumulator barrier of step s + 1 proves that every
    // warp has finished writing staging[s & 1], and the buffer is not rewritten before step s + 2, which no warp can
    // reach before all warps have handed over the last slice of step s + 1 (after this copy).
    const int nvalid = min(SPC, p.B - b0);
    auto write_out = [&](int s_) {
      const int f = dir ? T - 1 - s_ : s_;
      const unsigned char* sbuf = Ssm + (s_ & 1) * ST_BUF;
      float* gdst = hg + ((size_t)b0 * T + f) * 2 * C + dir * C + (tid & 15) * 4;
      float4 v[UPS];
#pragma unroll
      for (int i = 0; i < UPS; ++i) {
        const int r = (tid >> 4) + 32 * i;
        v[i] = *reinterpret_cast<const float4*>(sbuf + r * 256 + (((tid & 15) ^ (r & 15)) << 4));
      }
#pragma unroll
      for (int i = 0; i < UPS; ++i) {
        const int r = (tid >> 4) + 32 * i;
        if (r < nvalid) *reinterpret_cast<float4*>(gdst + (size_t)r * T * 2 * C) = v[i];
      }
    };
    // the thread's own UPS units out of the four columns of its cg group (a tcgen05.ld address is warp-uniform, so the
    // D parts of a warp load the same four columns and select)
    auto own_raw = [&](const uint32_t (&g)[4], int j) -> float {
      if constexpr (D == 1) return __uint_as_float(g[j]);
      else if constexpr (D == 2) return __uint_as_float(part ? g[2 + j] : g[j]);
      else return __uint_as_float((part & 2) ? ((part & 1) ? g[3] : g[2]) : ((part & 1) ? g[1] : g[0]));
    };
    // Split rows: this row's accumulators hold only the (hi | lo) * W part of the sum; the other part sits in the row of
    // the partner thread 16 lanes away (part ^ D/2), which owns other units - each sends the column the other one owns.
    auto own = [&](const uint32_t (&g)[4], int j) -> float {
      if constexpr (!SR) return own_raw(g, j);
      else {
        float theirs;
        if constexpr (D == 2) theirs = __uint_as_float(part ? g[j] : g[2 + j]);
        else theirs = __uint_as_float((part & 2) ? ((part & 1) ? g[1] : g[0]) : ((part & 1) ? g[3] : g[2]));
        return own_raw(g, j) + __shfl_xor_sync(0xffffffffu, theirs, 16);
      }
    };
    const int uoff = UPS * part;                             // first own unit inside the cg group's four
    for (int t = 0; t < T; ++t) {
      TL(0);
      if (t + 2 < T) load_x(t + 2, xv);                      // in flight during the wait
      mbar_wait(bars + 1, t & 1);
      tc_fence_after();
      TL(1);
      // hn is single buffered and slice 0 of the next step overwrites all of it: drain it first
      uint32_t ghn[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tmem_ld4_nowait(lane_base + TM_HN + 16 * ks + 4 * cg, ghn[ks]);
      // r, z pre-activations run one slice ahead of in / hn: the sigmoid stage of slice ks+1 is software pipelined
      // with the tanh stage of slice ks (four independent MUFU chains per thread instead of two)
      uint32_t grz[2][2][4], gi[4];
      const uint32_t pa = lane_base + TM_P + (t & 1) * 192 + 4 * cg;
      tmem_ld4_nowait(pa, grz[0][0]);
      tmem_ld4_nowait(pa + 64, grz[0][1]);
      tmem_ld4_nowait(pa + 16, grz[1][0]);
      tmem_ld4_nowait(pa + 64 + 16, grz[1][1]);
      tmem_ld4_nowait(pa + 128, gi);
      tmem_ld_wait();
      TL(2);
      unsigned char* srow_w = Ssm + (t & 1) * ST_BUF + srow * 256 + uoff * 4;
      const unsigned char* prow = Ssm + ((t + 1) & 1) * ST_BUF + srow * 256 + uoff * 4;      // h_{t-1} of this thread's units (FP32)
      const float2 one = make_float2(1.0f, 1.0f);
      // sigmoid stage: r = 1 / (1 + 2^a_r), z = 1 / (1 + 2^a_z) with one shared reciprocal.  The operand images and
      // biases carry the exponent scales (weights.py: r, z rows x -log2(e); n rows x 2 log2(e)).
      auto stage_a = [&](int ks, const uint32_t (&g_r)[4], const uint32_t (&g_z)[4], float (&r)[UPS], float (&z)[UPS]) {
        const int u0 = 16 * ks + 4 * cg + uoff;
        if constexpr (UPS >= 2) {
#pragma unroll
          for (int e = 0; e < UPS / 2; ++e) {
            const float2 b_r = *reinterpret_cast<const float2*>(sb + u0 + 2 * e);
            const float2 b_z = *reinterpret_cast<const float2*>(sb + C + u0 + 2 * e);
            const float2 ar = __fadd2_rn(make_float2(own(g_r, 2 * e), own(g_r, 2 * e + 1)), b_r);
            const float2 az = __fadd2_rn(make_float2(own(g_z, 2 * e), own(g_z, 2 * e + 1)), b_z);
#if ITC_POLY
            const float2 pr = __fadd2_rn(ex2_poly2(make_float2(clampf(ar.x, -125.f, 60.f), clampf(ar.y, -125.f, 60.f))), one);
            const float2 pz = __fadd2_rn(ex2_poly2(make_float2(clampf(az.x, -125.f, 60.f), clampf(az.y, -125.f, 60.f))), one);
#else
            // 2^60 * 2^60 stays finite in the shared reciprocal; sigmoid(-41) = 1e-18 is already 0 in FP32 terms
            const float2 pr = __fadd2_rn(make_float2(ex2_ftz(fminf(ar.x, 60.f)), ex2_ftz(fminf(ar.y, 60.f))), one);
            const float2 pz = __fadd2_rn(make_float2(ex2_ftz(fminf(az.x, 60.f)), ex2_ftz(fminf(az.y, 60.f))), one);
#endif
            const float2 pp = __fmul2_rn(pr, pz);
            const float2 ip = make_float2(rcp_ftz(pp.x), rcp_ftz(pp.y));
            const float2 rr = __fmul2_rn(ip, pz), zz = __fmul2_rn(ip, pr);
            r[2 * e] = rr.x; r[2 * e + 1] = rr.y;
            z[2 * e] = zz.x; z[2 * e + 1] = zz.y;
          }
        } else {                                             // one unit: r and z share the packed lanes instead
          const float2 a = __fadd2_rn(make_float2(own(g_r, 0), own(g_z, 0)), make_float2(sb[u0], sb[C + u0]));
          const float2 pq = __fadd2_rn(make_float2(ex2_ftz(fminf(a.x, 60.f)), ex2_ftz(fminf(a.y, 60.f))), one);
          const float ip = rcp_ftz(pq.x * pq.y);
          r[0] = ip * pq.y;
          z[0] = ip * pq.x;
        }
      };
      float rc[UPS], zc[UPS];
      stage_a(0, grz[0][0], grz[0][1], rc, zc);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int u0 = 16 * ks + 4 * cg + uoff;
        float rn[UPS], zn[UPS];
        if (ks < 3) stage_a(ks + 1, grz[(ks + 1) & 1][0], grz[(ks + 1) & 1][1], rn, zn);
        // tanh stage: n = tanh(c) = 1 - 2 / (1 + 2^c'),  h' = (1 - z) n + z h
        const int hchunk = ((4 * ks + cg) ^ (srow & 15)) << 4;
        float hp[UPS], hn[UPS];
        if constexpr (UPS == 4) {
          const float4 q = *reinterpret_cast<const float4*>(prow + hchunk);
          hp[0] = q.x; hp[1] = q.y; hp[2] = q.z; hp[3] = q.w;
        } else if constexpr (UPS == 2) {
          const float2 q = *reinterpret_cast<const float2*>(prow + hchunk);
          hp[0] = q.x; hp[1] = q.y;
        } else {
          hp[0] = *reinterpret_cast<const float*>(prow + hchunk);
        }
#ifdef ITC_NO_MATH
#pragma unroll
        for (int j = 0; j < UPS; ++j) hn[j] = (rc[j] + zc[j] + own(gi, j) + own(ghn[ks], j)) * 1e-30f + hp[j];
#else
        if constexpr (UPS >= 2) {
#pragma unroll
          for (int e = 0; e < UPS / 2; ++e) {
            const float2 b_i = *reinterpret_cast<const float2*>(sb + 2 * C + u0 + 2 * e);
            const float2 b_h = *reinterpret_cast<const float2*>(sb + 3 * C + u0 + 2 * e);
            const float2 vhn = __fadd2_rn(make_float2(own(ghn[ks], 2 * e), own(ghn[ks], 2 * e + 1)), b_h);
            const float2 vin = __fadd2_rn(make_float2(own(gi, 2 * e), own(gi, 2 * e + 1)), b_i);
            const float2 c = __ffma2_rn(make_float2(rc[2 * e], rc[2 * e + 1]), vhn, vin);
            const float2 pc = __fadd2_rn(make_float2(ex2_ftz(c.x), ex2_ftz(c.y)), one);
            const float2 q = make_float2(rcp_ftz(pc.x), rcp_ftz(pc.y));
            const float2 n = __ffma2_rn(make_float2(-2.0f, -2.0f), q, one);
            const float2 hv = __ffma2_rn(make_float2(zc[2 * e], zc[2 * e + 1]), __fadd2_rn(make_float2(hp[2 * e], hp[2 * e + 1]), make_float2(-n.x, -n.y)), n);
            hn[2 * e] = hv.x; hn[2 * e + 1] = hv.y;
          }
        } else {
          const float c = fmaf(rc[0], own(ghn[ks], 0) + sb[3 * C + u0], own(gi, 0) + sb[2 * C + u0]);
          const float n = fmaf(-2.0f, rcp_ftz(ex2_ftz(c) + 1.0f), 1.0f);
          hn[0] = fmaf(zc[0], hp[0] - n, n);
        }
#endif
        if (ks < 3) {                                        // pre-activations of the coming slices: in flight under the stores below
          tmem_ld4_nowait(pa + 128 + 16 * (ks + 1), gi);
          if (ks < 2) {
            tmem_ld4_nowait(pa + 16 * (ks + 2), grz[ks & 1][0]);
            tmem_ld4_nowait(pa + 64 + 16 * (ks + 2), grz[ks & 1][1]);
          }
#pragma unroll
          for (int j = 0; j < UPS; ++j) { rc[j] = rn[j]; zc[j] = zn[j]; }
        }
        // the cg group's four units of this slice as two packed FP16 columns (units 2c, 2c+1 -> column c), hi and lo:
        // every thread writes the complete pair of columns of its own row, the partner rows' units arrive by shuffle
        uint32_t hi0, lo0, hi1, lo1;
        if constexpr (D == 1) {
          split2_f16(hn[0], hn[1], hi0, lo0);
          split2_f16(hn[2], hn[3], hi1, lo1);
          *reinterpret_cast<float4*>(srow_w + hchunk) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        } else if constexpr (SR) {
          // split rows: a hi row needs the partner's hi halves of the other column, a lo row the partner's lo halves -
          // and the partner (16 lanes away) is always a row of the other kind, so each sends what it does not store
          uint32_t hi, lo;
          if constexpr (D == 2) {
            split2_f16(hn[0], hn[1], hi, lo);
            *reinterpret_cast<float2*>(srow_w + hchunk) = make_float2(hn[0], hn[1]);
          } else {
            const float o = __shfl_xor_sync(0xffffffffu, hn[0], 8);                // the other unit of this thread's column
            split2_f16((part & 1) ? o : hn[0], (part & 1) ? hn[0] : o, hi, lo);
            *reinterpret_cast<float*>(srow_w + hchunk) = hn[0];
          }
          const uint32_t recv = __shfl_xor_sync(0xffffffffu, lo_row ? hi : lo, 16);
          hi0 = lo_row ? recv : hi;                              // (column 0, column 1) of this row's kind
          hi1 = lo_row ? lo : recv;
          lo0 = lo1 = 0u;
        } else if constexpr (D == 2) {
          uint32_t hi, lo;
          split2_f16(hn[0], hn[1], hi, lo);
          const uint32_t ohi = __shfl_xor_sync(0xffffffffu, hi, 16), olo = __shfl_xor_sync(0xffffffffu, lo, 16);
          hi0 = part ? ohi : hi; hi1 = part ? hi : ohi;
          lo0 = part ? olo : lo; lo1 = part ? lo : olo;
          *reinterpret_cast<float2*>(srow_w + hchunk) = make_float2(hn[0], hn[1]);
        } else {
          const float o = __shfl_xor_sync(0xffffffffu, hn[0], 8);                  // the other unit of this thread's column
          uint32_t hi, lo;
          split2_f16((part & 1) ? o : hn[0], (part & 1) ? hn[0] : o, hi, lo);
          const uint32_t ohi = __shfl_xor_sync(0xffffffffu, hi, 16), olo = __shfl_xor_sync(0xffffffffu, lo, 16);
          hi0 = (part & 2) ? ohi : hi; hi1 = (part & 2) ? hi : ohi;
          lo0 = (part & 2) ? olo : lo; lo1 = (part & 2) ? lo : olo;
          *reinterpret_cast<float*>(srow_w + hchunk) = hn[0];
        }
        if constexpr (SR) {
          tmem_st2(lane_base + TM_HHI + 8 * ks + 2 * cg, hi0, hi1);
        } else {
          tmem_st2(lane_base + TM_HHI + 8 * ks + 2 * cg, hi0, hi1);
          tmem_st2(lane_base + TM_HLO + 8 * ks + 2 * cg, lo0, lo1);
        }
        if (ks == 1 && t > 0) write_out(t - 1);
        if (ks == 2 && t + 2 < T) store_x(t & 1, xv);        // x_mma(t) (reader of this buffer) completed with the commit
        if (ks == 3) {
          TL(3);
          TL(4);
          fence_async_smem();                                // generic-proxy smem writes -> visible to the tensor core
        }
        tmem_st_wait();
        if (ks < 3) tmem_ld_wait();
        if (t + 1 < T) {                                     // hand slice ks over to the issuer, do not wait
          tc_fence_before();
          if (ks == 0) asm volatile("bar.arrive 1, %0;" ::"n"(ITC_NT) : "memory");
          if (ks == 1) asm volatile("bar.arrive 2, %0;" ::"n"(ITC_NT) : "memory");
          if (ks == 2) asm volatile("bar.arrive 3, %0;" ::"n"(ITC_NT) : "memory");
          if (ks == 3) asm volatile("bar.arrive 4, %0;" ::"n"(ITC_NT) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 16) {                                           // write-out of the last step (all staging rows visible after the barrier)
    const int s_ = T - 1, f = dir ? 0 : T - 1;
    const unsigned char* sbuf = Ssm + (s_ & 1) * ST_BUF;
    const int nvalid = min(SPC, p.B - b0);
    float* gdst = hg + ((size_t)b0 * T + f) * 2 * C + dir * C + (tid & 15) * 4;
#pragma unroll
    for (int i = 0; i < UPS; ++i) {
      const int r = (tid >> 4) + 32 * i;
      if (r < nvalid)
        *reinterpret_cast<float4*>(gdst + (size_t)r * T * 2 * C) = *reinterpret_cast<const float4*>(sbuf + r * 256 + (((tid & 15) ^ (r & 15)) << 4));
    }
  }
  if (progress_ptr()) {
    __threadfence();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile int*>(progress_ptr()) = T;
  }
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// tcgen05.ld / tcgen05.st .16x128b.x1: thread T of the warp <-> (TMEM lane T / 4, column T % 4) in the first register and
// (lane 8 + T / 4, column T % 4) in the second (measured: tools/ubench/tmem_layout_probe.cu, profiles/r3d_*)
__device__ __forceinline__ void tmem_ld_16x128b_nowait(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_16x128b(uint32_t taddr, uint32_t r0, uint32_t r1) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}

// The sweep of one df-branch CTA in FRAGMENT form: 32 streams per CTA like D = 4, but a stream owns just TWO rows of the
// M = 128 tile - quadrant q, lanes i (hi halves of the operands) and 8 + i (lo halves), i = 0..7, lanes 16..31 idle - and
// its four gate threads are the four ADJACENT lanes 4 i + j of warp (q, cg): a .16x128b fragment load hands thread
// (i, j) column j of both rows at once, so a pre-activation is one tcgen05.ld and one add (hi * W + lo * W) instead of a
// 4-column load, six selects and a shuffle, and h goes back with one .16x128b fragment store per PAIR of K slices and no
// shuffle at all: thread (i, j) packs its units of slices 2p and 2p + 1 (16 (2p + e) + 4 cg + j, e = 0, 1) into ONE operand
// column 16 p + 4 cg + j (two K halves), the hi word to row i and the lo word to row 8 + i.  The recurrent matrix is
// packed with its K axis in that order (weights.py: tc.intra_f).  The gate math of a pair runs on packed f32x2 lanes.
template <int BR>      // branch: 0 = df, 1 = erb (worth it for the 48 kHz models only, whose erb sweep has 40 positions against 8)
__device__ __forceinline__ void intra_sweep_f(const IntraTcParams& p, const int dir, const int tile) {
  constexpr int SPC = 32;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* Wsm = smem_raw;
  unsigned char* Xsm = smem_raw + OFF_X;
  unsigned char* Ssm = smem_raw + OFF_ST;
  float* sb = reinterpret_cast<float*>(smem_raw + OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + OFF_BAR);       // [0] weights landed, [1] step accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3, cg = warp >> 2;
  const int si = lane >> 2, j = lane & 3;                    // stream of the quadrant, unit of the cg group
  const int srow = qd * 8 + si;                              // the stream's row in the staging tiles
  const int T = p.Fp[BR];
  const int b0 = tile * SPC;
  const float* __restrict__ xg = p.x[BR];
  float* __restrict__ hg = p.hcat[BR];

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid < 256) {                                           // biases of a pair's two units side by side: [gate][p][cg][j][e]
    const int u = tid & 63, ks = u >> 4;
    sb[(tid >> 6) * C + (((ks >> 1) * 4 + ((u >> 2) & 3)) * 4 + (u & 3)) * 2 + (ks & 1)] = __ldg(p.bias[BR] + dir * 4 * C + tid);
  }
  for (int i = tid; i < ST_BUF / 16; i += ITC_NT) reinterpret_cast<uint4*>(Ssm + ST_BUF)[i] = make_uint4(0u, 0u, 0u, 0u);   // h_{-1} = 0
  for (int i = tid; i < 4 * A_IMG / 16; i += ITC_NT) reinterpret_cast<uint4*>(Xsm)[i] = make_uint4(0u, 0u, 0u, 0u);        // idle operand rows stay 0
  auto progress_ptr = [&]() -> int* { return p.progress ? p.progress + (BR ? 2 * p.tiles[0] + dir * p.tiles[1] : dir * p.tiles[0]) + tile : nullptr; };
  __syncthreads();                                           // barriers initialised
  if (tid == 0) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wimg_f[BR]) + (size_t)dir * 4 * W_IMG;
    mbar_expect_tx(bars, 4 * W_IMG);
#pragma unroll
    for (int i = 0; i < 4; ++i) bulk_g2s(Wsm + i * W_IMG, src + (size_t)i * W_IMG, W_IMG, bars);
  }

  // ---- x tile staging: 32 streams x 8 chunks of 8 floats = 256 items, one per thread of warps 0..7 ----------------
  const bool x_active = warp < 8;
  const int xr = (warp >> 1) * 8 + (lane & 7), xkc = (warp & 1) * 4 + (lane >> 3);
  const int xrow = (xr >> 3) * 32 + (xr & 7);                // operand row of the stream's hi halves (lo: + 8)
  auto load_x = [&](int t, float (&v)[8]) {
    const int f = dir ? T - 1 - t : t;
    const int bb = b0 + xr;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (bb < p.B && x_active) {
      const float4* src = reinterpret_cast<const float4*>(xg + ((size_t)bb * T + f) * C + xkc * 8);
      a = __ldg(src);
      c = __ldg(src + 1);
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  };
  auto store_x = [&](int buf, const float (&v)[8]) {
    if (!x_active) return;
    uint4 hi, lo;
    split8_f16(v, hi, lo);
    if ((f16_nonfinite(hi.x) | f16_nonfinite(hi.y) | f16_nonfinite(hi.z) | f16_nonfinite(hi.w)) && p.err) p.err[DPDF_ERRW_RANGE] = 1;
    unsigned char* dst = Xsm + buf * 2 * A_IMG + img16_off(xrow, xkc);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + 1024) = lo;              // row + 8 = the next 8-row group of the image
  };
  pdl_wait();
  if (tid == 0 && progress_ptr()) {
    *reinterpret_cast<volatile int*>(progress_ptr()) = 0;    // the previous consumer of these counters has completed
    __threadfence();
  }
  __syncthreads();
  pdl_trigger();
  float xv[8];
  if (warp < 16) {
    load_x(0, xv);
    store_x(0, xv);
    if (T > 1) {
      load_x(1, xv);
      store_x(1, xv);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 16) {                                           // h_0 = 0 in the TMEM operand columns (all 32 lanes of the quadrant)
    const uint32_t a = *tmem_slot + ((uint32_t)(qd * 32) << 16) + TM_HHI + cg * 8;
    tmem_st4(a, 0u, 0u, 0u, 0u); tmem_st4(a + 4, 0u, 0u, 0u, 0u);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 16) {
    // ---- MMA issue: two passes per product (A rows carry hi | lo, B = W hi then W lo) ------------------------------
    int* my_progress = progress_ptr();
    const uint32_t w_base = smem_u32(Wsm), x_base = smem_u32(Xsm);
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    constexpr uint64_t DESC0 = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46);
    auto x_mma = [&](int t) {                                // P[t & 1][0, 192) = x_t * W_ih^T
      const uint32_t xa = x_base + (t & 1) * 2 * A_IMG, d = tmem + TM_P + (t & 1) * 192;
      const uint64_t da = DESC0 | (xa >> 4);
      const uint64_t dbh = DESC0 | (w_base >> 4), dbl = DESC0 | ((w_base + W_IMG) >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        umma_f16(d, da + ks * 16, dbh + ks * 16, idesc_f16(128, 192), ks > 0);
        umma_f16(d, da + ks * 16, dbl + ks * 16, idesc_f16(128, 192), 1);
      }
    };
    auto h_mma_slice = [&](int t, int kk, bool first) {      // K slice kk = operand columns [8 kk, 8 kk + 8); first of the step: hn is overwritten
      const uint32_t whi = w_base + 2 * W_IMG + kk * 256, wlo = w_base + 3 * W_IMG + kk * 256;
      const uint64_t rz_h = DESC0 | (whi >> 4), rz_l = DESC0 | (wlo >> 4);
      const uint64_t n_h = DESC0 | ((whi + 16 * 1024) >> 4), n_l = DESC0 | ((wlo + 16 * 1024) >> 4);
      const uint32_t a = tmem + TM_HHI + kk * 8;
      const uint32_t drz = tmem + TM_P + (t & 1) * 192, dn = tmem + TM_HN;
      umma_f16_ts(drz, a, rz_h, idesc_f16(128, 128), 1);
      umma_f16_ts(drz, a, rz_l, idesc_f16(128, 128), 1);
      umma_f16_ts(dn, a, n_h, idesc_f16(128, 64), first ? 0u : 1u);
      umma_f16_ts(dn, a, n_l, idesc_f16(128, 64), 1);
    };
    if (elect_one()) {
      mbar_wait(bars, 0);
      x_mma(0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) h_mma_slice(0, kk, kk == 0);
      umma_commit(bars + 1);
      if (T > 1) x_mma(1);
    }
    for (int t = 0; t + 1 < T; ++t) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // K slice kk = pair kk >> 1 of the eight gate warps with cg >> 1 == (kk & 1): they and this warp meet on barrier 1 + kk.
        // Of a pair's two slices the one of warps 8..15 is taken FIRST: those warps do not stage x, arrive earlier, and their
        // four MMAs then run while warps 0..7 finish - only one slice's MMAs stay behind the last arrive of a step.
        const int kk = i ^ 1;
        if (kk == 0) asm volatile("bar.sync 1, 288;" ::: "memory");
        if (kk == 1) asm volatile("bar.sync 2, 288;" ::: "memory");
        if (kk == 2) asm volatile("bar.sync 3, 288;" ::: "memory");
        if (kk == 3) asm volatile("bar.sync 4, 288;" ::: "memory");
        if (i == 3) TL(5);
        if (elect_one()) {
          tc_fence_after();
          h_mma_slice(t + 1, kk, i == 0);
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(bars + 1);
        TL(6);
        if (t + 2 < T) x_mma(t + 2);
        TL(7);
        if (my_progress) {
          __threadfence();
          *reinterpret_cast<volatile int*>(my_progress) = t;
        }
      }
      __syncwarp();
    }
  } else {
    // ---- the sweep (gate warps) -----------------------------------------------------------------------------------
    // TMEM addresses are warp-uniform operands: built from shuffled (provably uniform) values they live in uniform
    // registers, otherwise every tcgen05.ld / st pays a register -> uniform-register move first
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int cgu = warp_u >> 2;
    const uint32_t lane_base = __shfl_sync(0xffffffffu, *tmem_slot, 0) + ((uint32_t)((warp_u & 3) * 32) << 16);
    const int nvalid = min(SPC, p.B - b0);
    auto write_out = [&](int s_) {
      const int f = dir ? T - 1 - s_ : s_;
      const int r = tid >> 4;
      const float4 v = *reinterpret_cast<const float4*>(Ssm + (s_ & 1) * ST_BUF + r * 256 + (((tid & 15) ^ (r & 15)) << 4));
      if (r < nvalid) *reinterpret_cast<float4*>(hg + ((size_t)(b0 + r) * T + f) * 2 * C + dir * C + (tid & 15) * 4) = v;
    };
    const float2 one = make_float2(1.0f, 1.0f);
    float2 bias2[2][4];                                      // [pair][gate] of this thread's units: loop invariant
    {
      const float2* bp = reinterpret_cast<const float2*>(sb) + cg * 4 + j;        // + 16 p (+ 32 per gate)
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) bias2[pp][gt] = bp[16 * pp + 32 * gt];
    }
    // hi-row + lo-row accumulators of the pair's two units: scalar adds whose results the compiler places as an aligned
    // register pair (a packed add of (hi, hi) + (lo, lo) needs four moves to form its operands first)
    auto sum2 = [](const uint32_t (&a)[2], const uint32_t (&b)[2]) {
      return make_float2(__uint_as_float(a[0]) + __uint_as_float(a[1]), __uint_as_float(b[0]) + __uint_as_float(b[1]));
    };
    for (int t = 0; t < T; ++t) {
      TL(0);
      if (t + 2 < T) load_x(t + 2, xv);                      // in flight during the wait
      mbar_wait(bars + 1, t & 1);
      tc_fence_after();
      TL(1);
      const uint32_t pa = lane_base + TM_P + (t & 1) * 192 + 4 * cgu;
      uint32_t ghn[4][2], gr[2][2][2], gz[2][2][2], gi[2][2][2];                  // [pair][e][hi row | lo row]
      // the sigmoid stage of the first pair needs its r, z pre-activations only: wait for those four loads, then put
      // everything else in flight under that math (hn is single buffered and is drained before the pair's hand-over)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        tmem_ld_16x128b_nowait(pa + 16 * e, gr[0][e][0], gr[0][e][1]);
        tmem_ld_16x128b_nowait(pa + 64 + 16 * e, gz[0][e][0], gz[0][e][1]);
      }
      tmem_ld_wait();
      TL(2);
#pragma unroll
      for (int e = 0; e < 2; ++e) tmem_ld_16x128b_nowait(pa + 128 + 16 * e, gi[0][e][0], gi[0][e][1]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tmem_ld_16x128b_nowait(lane_base + TM_HN + 16 * ks + 4 * cgu, ghn[ks][0], ghn[ks][1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        tmem_ld_16x128b_nowait(pa + 32 + 16 * e, gr[1][e][0], gr[1][e][1]);
        tmem_ld_16x128b_nowait(pa + 64 + 32 + 16 * e, gz[1][e][0], gz[1][e][1]);
        tmem_ld_16x128b_nowait(pa + 128 + 32 + 16 * e, gi[1][e][0], gi[1][e][1]);
      }
      unsigned char* srow_w = Ssm + (t & 1) * ST_BUF + srow * 256 + j * 4;
      const unsigned char* prow = Ssm + ((t + 1) & 1) * ST_BUF + srow * 256 + j * 4;      // h_{t-1} of this thread's units (FP32)
      // pre-activation = hi-row accumulator + lo-row accumulator (+ bias); sigmoid / tanh with the exponent scales folded in.
      // Software pipeline: the sigmoid stage of the SECOND pair runs beside the tanh stage of the first (independent MUFU
      // chains in one basic block), so after the first hand-over only the second pair's tanh stage is left - the math of a
      // step is a dependent chain (ex2 -> rcp -> ex2 -> rcp), not an issue problem (profiles/r3k_*_rejected.txt).
      auto sig_stage = [&](int pp, float2& rr, float2& zz) {
        const float2 ar = __fadd2_rn(sum2(gr[pp][0], gr[pp][1]), bias2[pp][0]);
        const float2 az = __fadd2_rn(sum2(gz[pp][0], gz[pp][1]), bias2[pp][1]);
        const float2 pr = __fadd2_rn(make_float2(ex2_ftz(fminf(ar.x, 60.f)), ex2_ftz(fminf(ar.y, 60.f))), one);
        const float2 pz = __fadd2_rn(make_float2(ex2_ftz(fminf(az.x, 60.f)), ex2_ftz(fminf(az.y, 60.f))), one);
        const float2 pq = __fmul2_rn(pr, pz);
        const float2 ip = make_float2(rcp_ftz(pq.x), rcp_ftz(pq.y));
        rr = __fmul2_rn(ip, pz);
        zz = __fmul2_rn(ip, pr);
      };
      auto tanh_stage = [&](int pp, const float2 rr, const float2 zz) {
        const int hc0 = ((8 * pp + cg) ^ (srow & 15)) << 4, hc1 = ((8 * pp + 4 + cg) ^ (srow & 15)) << 4;      // staging chunks of the two units
        const float2 hp = make_float2(*reinterpret_cast<const float*>(prow + hc0), *reinterpret_cast<const float*>(prow + hc1));
        const float2 vin = __fadd2_rn(sum2(gi[pp][0], gi[pp][1]), bias2[pp][2]);
        const float2 vhn = __fadd2_rn(sum2(ghn[2 * pp], ghn[2 * pp + 1]), bias2[pp][3]);
        const float2 c = __ffma2_rn(rr, vhn, vin);
        const float2 pc = __fadd2_rn(make_float2(ex2_ftz(c.x), ex2_ftz(c.y)), one);
        const float2 q = make_float2(rcp_ftz(pc.x), rcp_ftz(pc.y));
        const float2 n = __ffma2_rn(make_float2(-2.0f, -2.0f), q, one);
        const float2 hv = __ffma2_rn(zz, __fadd2_rn(hp, make_float2(-n.x, -n.y)), n);
        uint32_t hi, lo;
        split2_f16(hv.x, hv.y, hi, lo);
        tmem_st_16x128b(lane_base + TM_HHI + 16 * pp + 4 * cgu, hi, lo);
        *reinterpret_cast<float*>(srow_w + hc0) = hv.x;
        *reinterpret_cast<float*>(srow_w + hc1) = hv.y;
      };
      float2 rr0, zz0, rr1, zz1;
      sig_stage(0, rr0, zz0);
      tmem_ld_wait();
      sig_stage(1, rr1, zz1);
      tanh_stage(0, rr0, zz0);
      tmem_st_wait();
      if (t + 1 < T) {                                       // hand the first pair's K slices over to the issuer, do not wait
        tc_fence_before();
        if (cg < 2) asm volatile("bar.arrive 1, 288;" ::: "memory");
        else asm volatile("bar.arrive 2, 288;" ::: "memory");
      }
      // the chores sit between the hand-over and the second pair's tanh stage: their loads, conversions and stores fill the
      // issue slots that chain leaves (after it they were a serial tail in front of the last arrive)
      if (t > 0) write_out(t - 1);
      if (t + 2 < T) store_x(t & 1, xv);                     // x_mma(t) (reader of this buffer) completed with the commit
      tanh_stage(1, rr1, zz1);
      TL(3);
      TL(4);
      fence_async_smem();                                    // generic-proxy smem writes -> visible to the tensor core
      tmem_st_wait();
      if (t + 1 < T) {
        tc_fence_before();
        if (cg < 2) asm volatile("bar.arrive 3, 288;" ::: "memory");
        else asm volatile("bar.arrive 4, 288;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 16) {                                           // write-out of the last step (all staging rows visible after the barrier)
    const int s_ = T - 1, f = dir ? 0 : T - 1;
    const int nvalid = min(SPC, p.B - b0), r = tid >> 4;
    if (r < nvalid)
      *reinterpret_cast<float4*>(hg + ((size_t)(b0 + r) * T + f) * 2 * C + dir * C + (tid & 15) * 4) =
          *reinterpret_cast<const float4*>(Ssm + (s_ & 1) * ST_BUF + r * 256 + (((tid & 15) ^ (r & 15)) << 4));
  }
  if (progress_ptr()) {
    __threadfence();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile int*>(progress_ptr()) = T;
  }
  if (warp == 0) tmem_dealloc<512>(*tmem_slot);
}

// Grid = df CTAs (2 directions x tiles[0]) followed by erb CTAs.  The erb sweep has F'e = 8 positions against the df
// sweep's 48, so it is never the critical path: it keeps full 128-stream tiles (DERB = 1) while the df branch is
// split DDF ways, which leaves more SMs to the overlapped post kernel than duplicating both branches.
template <int DDF, int DERB, int FORM = 0, int FORM_E = 0>      // FORM of the df / erb sweep: 0 = D copies of a row, 1 = split rows, 2 = fragment form (32 streams)
#ifdef ITC_MAXNREG
__global__ void __maxnreg__(ITC_MAXNREG) k_dprnn_intra_tc(const __grid_constant__ IntraTcParams p) {
#else
__global__ void __launch_bounds__(ITC_NT, 1) k_dprnn_intra_tc(const __grid_constant__ IntraTcParams p) {
#endif
  const int item = blockIdx.x, ndf = 2 * p.tiles[0];
  // two inlined copies of the sweep even for DDF == DERB: the branch index is then a compile-time constant in each
  // (one generic copy costs the gate warps live registers: 56 instead of 16 bytes of spills)
  if (item < ndf) {
    if constexpr (FORM == 2) intra_sweep_f<0>(p, item / p.tiles[0], item % p.tiles[0]);
    else intra_sweep<DDF, FORM == 1>(p, 0, item / p.tiles[0], item % p.tiles[0]);
  } else {
    if constexpr (FORM_E == 2) intra_sweep_f<1>(p, (item - ndf) / p.tiles[1], (item - ndf) % p.tiles[1]);
    else intra_sweep<DERB>(p, 1, (item - ndf) / p.tiles[1], (item - ndf) % p.tiles[1]);
  }
}

// Row duplication factor of the df-branch sweep for a step of B streams (Engine::intra_dup = 0: auto); the erb branch
// always runs D = 1.  Measured on dpdfnet4 (profiles/r2u_sweep_intra_dup.log, r2u_intra_timeline.txt): one sweep launch
// 168 / 113 / 99 us for D = 1 / 2 / 4 at 1024 streams; a sweep that needs more than one wave of its 225 KB CTAs loses
// outright, and one that holds most SMs squeezes the overlapped post kernel.  So: D = 4 while the sweep's CTAs take at
// most two thirds of the SMs, D = 2 while they fit one wave, else D = 1.
int intra_tc_dup(const Engine& e, int B) {
  if (e.intra_dup == 1 || e.intra_dup == 2 || e.intra_dup == 4) return e.intra_dup;
  const int Bt = std::max(B, e.total_B);                     // lanes run their sweeps side by side
  auto ctas = [&](int D) { return 2 * ((Bt * D + 127) / 128) + 2 * ((Bt + 127) / 128); };
  // The fragment form's 32-stream CTAs (intra_frag) win well beyond one wave of them, because the step then runs as lanes
  // (api.cu:lanes_for) whose sweeps are separate launches that rarely coincide (profiles/r3n_*, r3v_*: 2048 / 3072 / 4096
  // streams 1.01 / 1.44 / 1.83 ms per hop against 1.12 / 1.52 / 1.91 with D = 2 / 1 and the overlapped post kernel, r3w_*: 5120 /
  // 6144 streams 2.24 / 2.64 against 2.40 / 2.75; at 8192 streams the 128-stream tiles are back in front, 3.36 against 3.49 ms).
  // (48 kHz models: their separable convs keep the SMs busier, the cross-over is lower - StreamGroup ticks of dpdfnet8_48khz_hr
  // at 4096 / 4608 / 5120 streams 7.4 / 8.9 / 10.4 ms in fragment form against 7.8 / 8.6 / 9.2, profiles/r3X_bench.json)
  if (e.intra_frag && Bt <= (e.d.hr48 ? std::min(e.frag_max, 4096) : e.frag_max)) return 4;
  if (ctas(4) <= e.num_sms * 2 / 3) return 4;
  if (ctas(2) <= e.num_sms) return 2;
  return 1;
}

// Row count per stream of the ERB-branch sweep: 1, or 4 = fragment form (32 streams per CTA) for the 48 kHz models, whose erb
// sweep has 40 positions - with 128-stream tiles (5 700 cycles per step) it outlasts the df sweep's 48 fragment-form steps -
// while both branches' CTAs still fit one wave.
int intra_tc_dup_erb(const Engine& e, int B) {
  if (!e.intra_frag || !e.intra_frag_erb || e.d.fe[3] < 16 || intra_tc_dup(e, B) != 4) return 1;
  if (e.d.N == 0 || !e.w.dprnn_df[0].tc_intra_f || !e.w.dprnn_erb[0].tc_intra_f) return 1;      // blob packed before the form existed
  return std::max(B, e.total_B) <= e.frag_max / 3 ? 4 : 1;    // profiles/r3w_*: dpdfnet8_48khz_hr 2048 streams 3.24 -> 3.02 ms, 4096 streams 5.71 -> 5.79
}

void launch_dprnn_intra_tc(Engine& e, int blk, int B, cudaStream_t st) {
  IntraTcParams p{};
  p.x[0] = e.sc.c1;
  p.x[1] = blk == 0 ? e.sc.e3 : e.sc.xe;
  p.hcat[0] = e.sc.hcat_d;
  p.hcat[1] = e.sc.hcat_e;
  p.Fp[0] = NDF / 2;
  p.Fp[1] = e.d.fe[3];
  p.wimg[0] = e.w.dprnn_df[blk].tc_intra;  p.bias[0] = e.w.dprnn_df[blk].tc_intra_bias;
  p.wimg[1] = e.w.dprnn_erb[blk].tc_intra; p.bias[1] = e.w.dprnn_erb[blk].tc_intra_bias;
  p.B = B;
  const int D = intra_tc_dup(e, B);
  const int DE = intra_tc_dup_erb(e, B);
  p.tiles[0] = (B * D + 127) / 128;
  p.tiles[1] = (B * DE + 127) / 128;
  p.progress = e.overlap_now ? e.progress_dev + (size_t)e.cur_lane * 4 * e.progress_tiles : nullptr;
  p.err = e.err_dev;
  const dim3 grid(2 * (p.tiles[0] + p.tiles[1]));
  const bool sr = e.intra_sr == 1 ? D > 1 : (e.intra_sr == 2 && D == 4);    // auto: where the step is tensor bound (profiles/r3b_*)
  e.prio_now = e.sweep_prio;
  p.wimg_f[0] = e.w.dprnn_df[blk].tc_intra_f;
  p.wimg_f[1] = e.w.dprnn_erb[blk].tc_intra_f;
  if (D == 4 && e.intra_frag && p.wimg_f[0] && DE == 4 && p.wimg_f[1]) launch_k(e, k_dprnn_intra_tc<4, 4, 2, 2>, grid, dim3(ITC_NT), INTRA_TC_SMEM, st, p);
  else if (D == 4 && e.intra_frag && p.wimg_f[0]) launch_k(e, k_dprnn_intra_tc<4, 1, 2>, grid, dim3(ITC_NT), INTRA_TC_SMEM, st, p);
  else if (D == 4 && sr) launch_k(e, k_dprnn_intra_tc<4, 1, 1>, grid, dim3(ITC_NT), INTRA_TC_SMEM, st, p);
  else if (D == 2 && sr) launch_k(e, k_dprnn_intra_tc<2, 1, 1>, grid, dim3(ITC_NT), INTRA_TC_SMEM, st, p);
  else if (D == 4) launch_k(e, k_dprnn_intra_tc<4, 1>, grid, dim3(ITC_NT), INTRA_TC_SMEM, st, p);
  else if (D == 2) launch_k(e, k_dprnn_intra_tc<2, 1>, grid, dim3(ITC_NT), INTRA_TC_SMEM, st, p);
  else launch_k(e, k_dprnn_intra_tc<1, 1>, grid, dim3(ITC_NT), INTRA_TC_SMEM, st, p);
  e.prio_now = 0;
}

void init_dprnn_intra_tc_kernels() {
  cudaFuncSetAttribute(k_dprnn_intra_tc<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
  cudaFuncSetAttribute(k_dprnn_intra_tc<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
  cudaFuncSetAttribute(k_dprnn_intra_tc<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
  cudaFuncSetAttribute(k_dprnn_intra_tc<2, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
  cudaFuncSetAttribute(k_dprnn_intra_tc<4, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
  cudaFuncSetAttribute(k_dprnn_intra_tc<4, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
  cudaFuncSetAttribute(k_dprnn_intra_tc<4, 4, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
}

}  // namespace dpdf
