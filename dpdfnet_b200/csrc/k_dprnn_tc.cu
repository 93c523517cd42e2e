// k_dprnn_post_tc: the position-parallel half of a DPRNN block on the 5th-generation tensor cores.
//
// Same math as k_dprnn_post (fc_intra + LayerNorm + residual, inter-frame GRUCell, fc_inter + LayerNorm +
// residual; layers.py:178-196) for a 128-row tile, but every matrix product is a chain of
// tcgen05.mma.kind::f16 instructions with the FP32 accumulators in tensor memory.  FP32 accuracy is kept
// with the error-compensated split scheme: every operand is x = hi + lo with hi, lo in FP16 (11-bit significands,
// like TF32, at half the bytes and twice the tensor rate), and D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (the dropped
// lo*lo term and the rounding of lo are ~2^-22 relative).  Weights are split and laid out on the host
// (weights.py:umma_operand16); activations are split on the fly while they are staged into shared memory.
//
// Operand layout: K-major, SWIZZLE_NONE ("interleave"): 8 rows x 16 B core matrices (128 B contiguous), core
// matrices adjacent in K are LBO = 128 B apart, 8-row groups are SBO = (K/8)*128 B = 1 KB apart.
// One CTA = 512 threads, thread (row = TMEM lane, 16-column group): LayerNorm and the GRU gate math are register
// code straight out of tcgen05.ld.  104 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM, so the
// load -> MMA -> epilogue chain of one tile overlaps with the other's.
//
// PAIR = true (template parameter): the kernel is launched as clusters of two CTAs on the two SMs of a TPC and every
// matrix product is ONE tcgen05.mma.cta_group::2 of M = 256 over the pair's two tiles.  The B operand of such an MMA is
// split across the pair - each CTA holds 32 of a slab's 64 rows - so a weight slab costs 8 KB per CTA instead of 16:
// the same 32 KB ring is FOUR slabs deep instead of two and every slab is pulled from L2 once per 256 rows.  What bounded
// the single-CTA form was exactly that ring (profiles/r2z_post_timeline.txt: 11-17 k of a tile's 37-44 k cycles are
// phase 2, 72 MMAs that need 3 k, because each pair of slabs is requested only when its buffer drains and a bulk copy
// from L2 takes 1-2 us).  Only the leader (cluster rank 0) issues MMAs; the peer tells it with remote mbarrier arrives
// when its half of a slab has landed and when its activation images of a phase are staged; commits are multicast to
// the barriers of both CTAs.
#include "engine.h"
#include "tc_common.cuh"

namespace dpdf {

namespace {

using namespace tc;

constexpr int TC_NT = 512;        // 16 warps: warp w -> TMEM lane quadrant w & 3, 16-column group w >> 2
constexpr int IMG = 128 * 64 * 2; // bytes of one FP16 [128][64] activation image (16 KB)
constexpr int WSLAB = 2 * 64 * 64 * 2;   // bytes of one [64][64] weight slab, hi image | lo image (16 KB)

__device__ __forceinline__ float2 h2f(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
// the 8 values of a 16-byte operand row chunk, reconstructed as hi + lo
__device__ __forceinline__ void unsplit8(const uint4& hi, const uint4& lo, float (&v)[8]) {
  const float2 a0 = h2f(hi.x), a1 = h2f(hi.y), a2 = h2f(hi.z), a3 = h2f(hi.w);
  const float2 b0 = h2f(lo.x), b1 = h2f(lo.y), b2 = h2f(lo.z), b3 = h2f(lo.w);
  v[0] = a0.x + b0.x; v[1] = a0.y + b0.y; v[2] = a1.x + b1.x; v[3] = a1.y + b1.y;
  v[4] = a2.x + b2.x; v[5] = a2.y + b2.y; v[6] = a3.x + b3.x; v[7] = a3.y + b3.y;
}

}  // namespace

struct PostTcBranch {
  const float* hcat;      // [rows][128]
  const float* xin;       // [rows][64]
  float* xout;            // [rows][64]
  float* hstate;          // inter-GRU state of this block: + slot*per_slot + f*64
  long long per_slot;
  int Fp;
  const float *tc_fc_w, *tc_gates, *tc_fc2_w;         // [64x64] weight slabs, each hi image | lo image (32 KB)
  const float *fc_b, *ln_g, *ln_b, *bias, *fc2_b, *ln2_g, *ln2_b;
};
struct PostTcParams {
  const IoDesc* io;
  PostTcBranch br[2];
  int tiles0, B;
  // overlapped mode: tile = (position f, 128-stream tile), erb tiles first, positions middle-out (the
  // order in which both sweep directions complete them); the CTA waits for the two intra CTAs it depends on
  int tiles1;             // tile counts per branch as launched (PAIR: rounded up to even so that a CTA pair never straddles
                          // the branches, which have different weights; the padding tile has no valid rows)
  const int* progress;    // sweep counters [df: 2 dirs x ptiles | erb: 2 dirs x stiles], nullptr = row-major tiles of a finished sweep
  int stiles;             // ceil(B / 128)
  int dup, ptiles;        // the df sweep's CTAs own 128 / dup streams each: ptiles = ceil(B * dup / 128) counters per direction
  int dup_e, ptiles_e;    // the same for the erb sweep (1, or 4 in fragment form: intra_tc_dup_erb)
  int* ctr;               // k_dprnn_post_res: [0] next df tile, [1] next erb tile, [2] CTAs that have finished (the last one zeroes all three)
  int ctas_erb;           // k_dprnn_post_res: CTAs [0, ctas_erb) serve the erb branch, the others the df branch
  int pf_dist;            // row-major mode: warm L2 with the inputs of tile blockIdx.x + pf_dist (0 = off); CTAs are dispatched in
                          // index order, so with pf_dist = resident CTAs that tile starts about when this one ends
#ifdef PT_TIMELINE
  long long* tl;          // [16] SM-clock stamps of CTA 0 (tools/ubench/post_tc_timeline.cu)
#endif
};

#ifdef PT_TIMELINE
#define PTL(slot) do { if (blockIdx.x == 0 && tid == 64) p.tl[slot] = clock64(); } while (0)
#else
#define PTL(slot) do { } while (0)
#endif

// Shared-memory layout per kernel form (MODE 0: one tile, 2-slab ring; 1: CTA pair, 4 half slabs; 2: two tiles per CTA,
// 5-slab ring shared by both): operand images | weight ring | small parameters | LayerNorm partials | state offsets |
// commit flags | mbarriers + TMEM base slot
template <int MODE>
struct PostLayout {
  static constexpr int NT = MODE == 2 ? 2 : 1;                          // tiles per CTA
  static constexpr int NBUF = MODE == 1 ? 4 : (MODE == 2 ? 5 : 2);      // ring depth in slabs
  static constexpr int SLAB_B = MODE == 1 ? WSLAB / 2 : WSLAB;          // bytes of a slab in this CTA's ring
  static constexpr int OFF_W = NT * 4 * IMG;
  static constexpr int OFF_SP = OFF_W + NBUF * SLAB_B;                  // 640 floats of small parameters
  static constexpr int OFF_RED = OFF_SP + 640 * 4;                      // [NT][2][4][128] LayerNorm partials
  static constexpr int OFF_HOFF = OFF_RED + NT * 1024 * 4;              // [NT][128] int64 state offsets
  static constexpr int OFF_COMMIT = OFF_HOFF + NT * 128 * 8;            // [NT][128] int
  static constexpr int OFF_BAR = OFF_COMMIT + NT * 128 * 4;             // 16 mbarriers + TMEM base slot
  static constexpr size_t SMEM = OFF_BAR + 144;
};
constexpr size_t POST_TC_SMEM = PostLayout<0>::SMEM;
// mbarriers: slab landed [0,5), slab consumed [5,10), peer's half of the slab landed [10,14) (leader of a pair only),
// phase result ready [14], peer's activation images staged [15] (leader of a pair only)
constexpr int BAR_FULL = 0, BAR_CONS = 5, BAR_PFULL = 10, BAR_PHASE = 14, BAR_ACT = 15, NBARS = 16;

// One CTA = one 128-row tile.  Thread 0 is the single MMA issuer (PAIR: thread 0 of the leader CTA, for both tiles) and
// streams the nine weight slabs of the block (fc_intra K-halves, six GRU gate slabs, fc_inter) through a ring of NBUF
// buffers with 1-D bulk copies that complete on mbarriers, NBUF slabs ahead of the tensor core; the other 511 threads
// never wait for weights.
// TMEM columns: [0,64) fc_intra, then the gate pre-activations r [0,64) z [64,128) in [128,192) hn [192,256),
// then fc_inter in [0,64) again (each phase is drained by all threads before the next one is issued).
//
// MODE 2 (DUAL): one CTA of 1024 threads per SM works on TWO consecutive tiles in lock step - thread group g = tid >> 9 is the
// 512-thread tile team of MODE 0 with its own operand images, LayerNorm scratch and 256 TMEM columns - and the two teams
// share ONE weight ring: every slab is pulled once per 256 rows and, with 128 KB instead of 2 x 64 KB of images, five
// slabs fit where two co-resident CTAs hold 2 x 2.  Unlike the CTA pair (MODE 1) the teams meet on CTA barriers, not on
// remote mbarriers, and unlike the persistent form (k_dprnn_post_res) two tiles are in flight per SM.
template <int MODE>
__global__ void __launch_bounds__(MODE == 2 ? 2 * TC_NT : TC_NT, MODE == 2 ? 1 : 2) k_dprnn_post_tc(PostTcParams p) {
  constexpr bool PAIR = MODE == 1, DUAL = MODE == 2;
  using L = PostLayout<MODE>;
  constexpr int NBUF = L::NBUF, SLAB_B = L::SLAB_B;
  pdl_trigger();
  if (!p.progress) pdl_wait();      // overlapped mode synchronises with the sweep through its progress counters instead
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid_all = threadIdx.x;
  const int grp = DUAL ? tid_all >> 9 : 0;                 // tile team
  const int tid = DUAL ? (tid_all & 511) : tid_all;        // thread inside its team
  unsigned char* RA = smem_raw + grp * 4 * IMG;     // four activation operand images (per team)
  unsigned char* RW = smem_raw + L::OFF_W;
  float* sp = reinterpret_cast<float*>(smem_raw + L::OFF_SP);
  float* red = reinterpret_cast<float*>(smem_raw + L::OFF_RED) + grp * 1024;
  long long* s_hoff = reinterpret_cast<long long*>(smem_raw + L::OFF_HOFF) + grp * 128;
  int* s_commit = reinterpret_cast<int*>(smem_raw + L::OFF_COMMIT) + grp * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::OFF_BAR);   // BAR_* above
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;                    // 0 = leader (issues the pair's MMAs)

  const int warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3, cg = warp >> 2, row = qd * 32 + lane;
  const int tile = DUAL ? (int)blockIdx.x * 2 + grp : (int)blockIdx.x;
  int bi, valid, fpos = 0;
  long long rbase, rstride;
  uint32_t ovf = 0;                                        // FP16 range guard of the y converter (tc_common.cuh:f16_nonfinite)
  __shared__ int s_abort_[2];
  int& s_abort = s_abort_[grp];
  if (tid == 0) s_abort = 0;
  if (p.progress) {
    const int tiles_e = p.tiles1;                          // erb tiles first (PAIR / DUAL: padded to even)
    bi = tile < tiles_e ? 1 : 0;
    const int local = tile - (bi ? 0 : tiles_e);
    const int k = local / p.stiles, stile = local % p.stiles, T = p.br[bi].Fp;
    fpos = (T - 1) / 2 + ((k & 1) ? (k + 1) / 2 : -(k / 2));
    rbase = (long long)stile * 128 * T + fpos;
    rstride = T;
    valid = k < T ? min(128, p.B - stile * 128) : 0;       // k == T: the padding tile of an odd tile count
  } else {
    bi = tile >= p.tiles0 ? 1 : 0;
    rbase = (long long)(tile - (bi ? p.tiles0 : 0)) * 128;
    rstride = 1;
    valid = (int)max(0ll, min((long long)128, (long long)p.B * p.br[bi].Fp - rbase));   // 0: padding tile (PAIR)
  }
  const PostTcBranch& q = p.br[bi];

  // ---- weight slab ring (thread 0 only) --------------------------------------------------------------------
  auto slab_src = [&](int i) -> const unsigned char* {
    if (i < 2) return reinterpret_cast<const unsigned char*>(q.tc_fc_w) + (size_t)i * WSLAB;
    if (i == 8) return reinterpret_cast<const unsigned char*>(q.tc_fc2_w);
    const int pidx = i - 2;                                // processing order r(y),r(h),z(y),z(h),n(y),n(h)
    return reinterpret_cast<const unsigned char*>(q.tc_gates) + (size_t)((pidx & 1) ? 3 + (pidx >> 1) : (pidx >> 1)) * WSLAB;
  };
  auto load_slab = [&](int i) {
    const int buf = i % NBUF;
    if (i >= NBUF) mbar_wait(bars + BAR_CONS + buf, (i / NBUF - 1) & 1);      // MMAs of the previous tenant have completed
    mbar_expect_tx(bars + BAR_FULL + buf, SLAB_B);
    if constexpr (PAIR) {                                  // rows [32 rank, 32 rank + 32) of the hi and of the lo image: 4 KB each
      bulk_g2s(RW + buf * SLAB_B, slab_src(i) + rank * (WSLAB / 4), WSLAB / 4, bars + BAR_FULL + buf);
      bulk_g2s(RW + buf * SLAB_B + WSLAB / 4, slab_src(i) + WSLAB / 2 + rank * (WSLAB / 4), WSLAB / 4, bars + BAR_FULL + buf);
    } else {
      bulk_g2s(RW + buf * SLAB_B, slab_src(i), WSLAB, bars + BAR_FULL + buf);
    }
  };
  if (tid_all == 0) {
#pragma unroll
    for (int i = 0; i < NBARS; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first slabs do not depend on the sweep: in overlapped mode they land while this CTA waits for its rows
#pragma unroll
    for (int i = 0; i < NBUF; ++i) load_slab(i);
  }
  if (p.progress) {
    const int stile = (tile - (bi ? 0 : p.tiles1)) % p.stiles, T = q.Fp;
    if (tid == 0 && valid > 0) {
      // the sweep CTAs this tile's 128 streams come from: dup per direction (branch index as in the intra kernel: 0 = df, 1 = erb)
      const int dup = bi ? p.dup_e : p.dup, pt = bi ? p.ptiles_e : p.ptiles;
      const volatile int* fw = p.progress + (bi ? 2 * p.ptiles : 0) + stile * dup;
      const volatile int* bw = fw + pt;
      const int nsrc = min(dup, pt - stile * dup);
      // a forward CTA has finished position f after f + 1 steps, a backward one after T - f; bounded wait (~2 s):
      // a broken or preempted producer must show up as an error, not as a hung GPU - and never as stale data: on
      // time-out the tile is SKIPPED (nothing read, nothing written) and the engine's error word is raised, which
      // the host turns into DPDF_ERR_CUDA for this hop (api.cu:check_device_errors)
      auto ready = [&]() {
        bool ok = true;
        for (int k = 0; k < nsrc; ++k) ok = ok && fw[k] >= fpos + 1 && bw[k] >= T - fpos;
        return ok;
      };
      long long spin = 0;
      for (; !ready() && spin < (1ll << 23); ++spin) __nanosleep(256);
      if (!ready()) {
        s_abort = 1;
        p.io->err[DPDF_ERRW_OVERLAP] = 1;
        __threadfence_system();
      }
      __threadfence();                                     // acquire: the rows those counts cover are visible (loads below bypass L1)
    }
  }

  // Warm L2 for the CTA that will take this CTA's place: its hcat tile (64 KB) and block input (32 KB) are contiguous,
  // its 128 inter-GRU state rows are scattered over the slot arena.  At throughput batch sizes none of them is L2
  // resident any more (hcat alone is 0.4 GB at 16 384 streams) and the tile's dependent phases would each eat a DRAM
  // round trip: ncu showed 38 % long-scoreboard stalls, hcat staging 9 k and the first epilogue 12 k of 44 k cycles.
  if (p.pf_dist > 0 && !p.progress) {
    const long long nt = (long long)tile + p.pf_dist;
    if (nt < (long long)p.tiles0 + p.tiles1) {
      const int nb = nt >= p.tiles0 ? 1 : 0;
      const PostTcBranch& nq = p.br[nb];
      const long long nbase = (nt - (nb ? p.tiles0 : 0)) * 128;
      const int nvalid = (int)max(0ll, min((long long)128, (long long)p.B * nq.Fp - nbase));
      if (nvalid == 0) {
      } else if (tid == 0) {
        bulk_prefetch_l2(nq.hcat + nbase * 2 * C, (uint32_t)nvalid * 2 * C * 4);
        bulk_prefetch_l2(nq.xin + nbase * C, (uint32_t)nvalid * C * 4);
      } else if (tid >= 128 && tid < 128 + nvalid) {
        const long long r = nbase + (tid - 128);
        const int b = (int)(r / nq.Fp), f = (int)(r % nq.Fp);
        const float* hp = nq.hstate + (long long)io_slot(p.io, b) * nq.per_slot + (long long)f * C;
        prefetch_l2(hp);
        prefetch_l2(hp + 32);
      }
    }
  }

  if (tid < 128) {
    long long off = 0;
    int commit = 0;
    if (tid < valid) {
      int b, f;
      if (p.progress) { b = (int)(rbase / q.Fp) + tid; f = fpos; }
      else { const long long r = rbase + tid; b = (int)(r / q.Fp); f = (int)(r % q.Fp); }
      off = (long long)io_slot(p.io, b) * q.per_slot + (long long)f * C;
      commit = (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) ? 0 : 1;
    }
    s_hoff[tid] = off;
    s_commit[tid] = commit;
  }
  if (tid < 64) {
    sp[tid] = q.fc_b[tid]; sp[64 + tid] = q.ln_g[tid]; sp[128 + tid] = q.ln_b[tid];
    sp[448 + tid] = q.fc2_b[tid]; sp[512 + tid] = q.ln2_g[tid]; sp[576 + tid] = q.ln2_b[tid];
  }
  if (tid < 256) sp[192 + tid] = q.bias[tid];
  if constexpr (PAIR) {
    cluster_sync_all();                                  // both CTAs' barriers exist before either signals across; s_hoff visible
    if (warp == 0) tmem_alloc_2cta<256>(tmem_slot);
    __syncthreads();
    if (s_abort) valid = 0;                              // the sweep never delivered this tile's rows: the pair protocol still runs,
                                                         // but nothing is read or written for this tile (see above)
  } else if constexpr (DUAL) {
    if (tid_all < 32) tmem_alloc<512>(tmem_slot);
    __syncthreads();
    if (s_abort) valid = 0;                              // as for the pair: the other team's tile goes on
  } else {
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    __syncthreads();                                     // barriers initialised, s_hoff visible
    if (s_abort) {                                       // the sweep never delivered this tile's rows: skip it (see above)
      tc_fence_after();
      if (tid == 0) {                                      // ... once the slab copies already in flight have landed
#pragma unroll
        for (int i = 0; i < NBUF; ++i) mbar_wait(bars + BAR_FULL + i, 0);
      }
      __syncthreads();
      if (warp == 0) tmem_dealloc<256>(*tmem_slot);
      return;
    }
  }

  PTL(0);

  // ---- stage the hcat tile as two K=64 operand image pairs (split on the fly) --------------------------------
  // a warp instruction covers 8 rows x 16 floats; thread = (row rr of the group, float4 cc): its four halves are
  // 8 bytes of a 16-byte operand row chunk, the 32 lanes write two full 128-byte core matrices
  const int rr = lane >> 2, cc = lane & 3;
  const int sub_off = (cc >> 1) * 128 + rr * 16 + (cc & 1) * 8;      // inside a [8 rows][16 k] block of an image
  {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                          // 128 blocks of 8 rows x 16 floats, 8 per warp, all in flight
      const int blk = warp + 16 * i, rg = blk >> 3, kb = blk & 7;
      const int r = rg * 8 + rr;
      v[i] = r < valid ? __ldcg(reinterpret_cast<const float4*>(q.hcat + (rbase + r * rstride) * 2 * C + kb * 16 + cc * 4))   // L2: written by a concurrent kernel
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int blk = warp + 16 * i, rg = blk >> 3, kb = blk & 7;
      uint2 h, l;
      split2_f16(v[i].x, v[i].y, h.x, l.x);
      split2_f16(v[i].z, v[i].w, h.y, l.y);
      unsigned char* dst = RA + (kb >> 2) * 2 * IMG + rg * 1024 + (kb & 3) * 256 + sub_off;
      *reinterpret_cast<uint2*>(dst) = h;
      *reinterpret_cast<uint2*>(dst + IMG) = l;
    }
  }
  // prefetch what the first epilogue needs while the tensor core works: the residual input
  float4 xv[4];
  {
    const float* xr = q.xin + (rbase + row * rstride) * C + cg * 16;
#pragma unroll
    for (int c = 0; c < 4; ++c) xv[c] = row < valid ? __ldg(reinterpret_cast<const float4*>(xr + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  PTL(1);
  const uint32_t lane_base = tmem + ((uint32_t)(qd * 32) << 16) + grp * 256 + cg * 16;   // a team's 256 TMEM columns
  constexpr uint32_t IDESC = idesc_f16(PAIR ? 256 : 128, 64);
  const uint32_t a_base = smem_u32(smem_raw), w_base = smem_u32(RW);      // MMAs are issued by thread 0 for every tile of the CTA

  // slab i: D[col .. col+64) (+)= A[128 or 256][64] * W_i[64][64]^T, three FP16 passes per 16-wide k-step
  auto run_slab = [&](int i, int a_img, uint32_t col, uint32_t accumulate) {
    const int buf = i % NBUF;
    mbar_wait(bars + BAR_FULL + buf, (i / NBUF) & 1);      // slab landed (async proxy write -> async proxy read)
    if constexpr (PAIR) {
      if (rank != 0) {                                     // peer: relay "my half has landed" to the leader, which issues for both
        mbar_arrive_remote(bars + BAR_PFULL + buf, 0);
        return;
      }
      mbar_wait_cluster(bars + BAR_PFULL + buf, (i / NBUF) & 1);
    }
    constexpr uint64_t DESC0 = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t ah = a_base + a_img * IMG, bh = w_base + buf * SLAB_B;
    const uint64_t dah = DESC0 | (ah >> 4), dal = dah + (IMG >> 4), dbh = DESC0 | (bh >> 4), dbl = dbh + (SLAB_B / 2 >> 4);
#pragma unroll
    for (int g = 0; g < L::NT; ++g) {                      // DUAL: the same slab serves both teams' tiles
      const uint64_t ga = (uint64_t)(g * 4 * IMG >> 4);
      const uint32_t d = tmem + g * 256 + col;
      uint32_t acc = accumulate;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                     // 16 halves of K per step = two core matrices = 256 B
        if constexpr (PAIR) {
          umma_f16_2cta(d, dah + ks * 16, dbh + ks * 16, IDESC, acc);
          umma_f16_2cta(d, dal + ks * 16, dbh + ks * 16, IDESC, 1);
          umma_f16_2cta(d, dah + ks * 16, dbl + ks * 16, IDESC, 1);
        } else {
          umma_f16(d, dah + ga + ks * 16, dbh + ks * 16, IDESC, acc);
          umma_f16(d, dal + ga + ks * 16, dbh + ks * 16, IDESC, 1);
          umma_f16(d, dah + ga + ks * 16, dbl + ks * 16, IDESC, 1);
        }
        acc = 1;
      }
    }
    if constexpr (PAIR) umma_commit_2cta(bars + BAR_CONS + buf);
    else umma_commit(bars + BAR_CONS + buf);
  };
  // All operands of a phase are staged in this CTA (the __syncthreads before); PAIR: the leader also needs the peer's
  auto phase_begin = [&](unsigned use) {                   // thread 0 only; use = 0, 1, 2
    if constexpr (PAIR) {
      if (rank != 0) mbar_arrive_remote(bars + BAR_ACT, 0);
      else { mbar_wait_cluster(bars + BAR_ACT, use & 1); tc_fence_after(); }
    }
  };
  auto phase_commit = [&]() {                              // thread 0 only: phase result ready, in both CTAs of a pair
    if constexpr (PAIR) { if (rank == 0) umma_commit_2cta(bars + BAR_PHASE); }
    else umma_commit(bars + BAR_PHASE);
  };

  // ---- phase 1: acc[0,64) = hcat * fc_intra^T ---------------------------------------------------------------
  // The issuing lane is chosen with elect.sync in the converged warp 0 (tc_common.cuh:elect_one): under `tid == 0` the
  // compiler wrapped every tcgen05.mma and bulk copy in an elect / broadcast loop of its own (626 of them in this file).
  if (tid_all < 32 && elect_one()) {
    phase_begin(0);
    run_slab(0, 0, 0, 0);
    run_slab(1, 2, 0, 1);
    phase_commit();
    load_slab(NBUF);
    load_slab(NBUF + 1);
  }
  float4 hv[4];                                            // h_prev tile: its latency hides behind the phase-1 MMAs
#pragma unroll
  for (int i = 0; i < 4; ++i) {                            // 64 blocks of 8 rows x 16 floats, 4 per warp
    const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
    const int r = rg * 8 + rr;
    hv[i] = r < valid ? __ldg(reinterpret_cast<const float4*>(q.hstate + s_hoff[r] + kb * 16 + cc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // one warp polls the mbarrier, the other fifteen park on the hardware barrier: 512 spinning threads cost a quarter of
  // the kernel's issued instructions (ncu, r2c: SYNCS + BRA + YIELD = 25 %) and stole issue slots from the co-resident CTA
  if (warp == 0) mbar_wait(bars + BAR_PHASE, 0);
  __syncthreads();
  tc_fence_after();
  PTL(2);

  unsigned char* y_hi = RA;                                // images 0, 1: y = block input of the inter-frame half
  unsigned char* h_hi = RA + 2 * IMG;                      // images 2, 3: h_prev, then h_new
  // row statistics over 64 columns held by the four column-group warps of a lane quadrant: only those four warps
  // exchange partials (rows of different quadrants are disjoint), so they meet on a named barrier of their own
  auto quad_sync = [&]() {                                 // immediate barrier ids: a register id makes ptxas reserve all sixteen
    const int id = qd + 4 * grp;
    if (id == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else if (id == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
    else if (id == 2) asm volatile("bar.sync 3, 128;" ::: "memory");
    else if (id == 3) asm volatile("bar.sync 4, 128;" ::: "memory");
    else if (id == 4) asm volatile("bar.sync 5, 128;" ::: "memory");
    else if (id == 5) asm volatile("bar.sync 6, 128;" ::: "memory");
    else if (id == 6) asm volatile("bar.sync 7, 128;" ::: "memory");
    else asm volatile("bar.sync 8, 128;" ::: "memory");
  };
  auto layernorm16 = [&](float (&v)[16], const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    red[cg * 128 + row] = s;
    quad_sync();
    const float mean = (red[row] + red[128 + row] + red[256 + row] + red[384 + row]) * (1.0f / 64.0f);
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] -= mean; qq += v[i] * v[i]; }
    red[512 + cg * 128 + row] = qq;
    quad_sync();
    const float rstd = rsqrtf((red[512 + row] + red[640 + row] + red[768 + row] + red[896 + row]) * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] * rstd * g[cg * 16 + i] + b[cg * 16 + i];
  };
  {
    float v[16];
    tmem_ld16(lane_base, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += sp[cg * 16 + i];
    layernorm16(v, sp + 64, sp + 128);                     // its barriers also order the hcat image reads (MMAs done) before the writes below
#pragma unroll
    for (int c = 0; c < 2; ++c) {                           // y = LN(...) + x, staged as the hi / lo operand of the gate GEMMs
      float y[8];
      y[0] = v[c * 8 + 0] + xv[2 * c].x; y[1] = v[c * 8 + 1] + xv[2 * c].y; y[2] = v[c * 8 + 2] + xv[2 * c].z; y[3] = v[c * 8 + 3] + xv[2 * c].w;
      y[4] = v[c * 8 + 4] + xv[2 * c + 1].x; y[5] = v[c * 8 + 5] + xv[2 * c + 1].y; y[6] = v[c * 8 + 6] + xv[2 * c + 1].z; y[7] = v[c * 8 + 7] + xv[2 * c + 1].w;
      uint4 h, l;
      split8_f16(y, h, l);
      ovf |= f16_nonfinite(h.x) | f16_nonfinite(h.y) | f16_nonfinite(h.z) | f16_nonfinite(h.w);
      const int off = img16_off(row, cg * 2 + c);
      *reinterpret_cast<uint4*>(y_hi + off) = h;
      *reinterpret_cast<uint4*>(y_hi + IMG + off) = l;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                           // h_prev tile (loaded before the wait) -> hi / lo images
      const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
      uint2 h, l;
      split2_f16(hv[i].x, hv[i].y, h.x, l.x);
      split2_f16(hv[i].z, hv[i].w, h.y, l.y);
      unsigned char* dst = h_hi + rg * 1024 + kb * 256 + sub_off;
      *reinterpret_cast<uint2*>(dst) = h;
      *reinterpret_cast<uint2*>(dst + IMG) = l;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  PTL(3);
  // ---- phase 2: GRU gate pre-activations in TMEM: r [0,64) z [64,128) in [128,192) hn [192,256) ---------------
  if (tid_all < 32 && elect_one()) {
    phase_begin(1);
    run_slab(2, 0, 0, 0);     // Wih_r * y
    run_slab(3, 2, 0, 1);     // Whh_r * h
    if constexpr (MODE == 0) { load_slab(4); load_slab(5); }   // PAIR / DUAL: slabs 4, 5 were requested before
    run_slab(4, 0, 64, 0);    // Wih_z * y
    run_slab(5, 2, 64, 1);    // Whh_z * h
    if constexpr (DUAL) { load_slab(7); load_slab(8); }        // slab 6 came in under the first epilogue
    else { load_slab(6); load_slab(7); }
    run_slab(6, 0, 128, 0);   // Wih_n * y
    run_slab(7, 2, 192, 0);   // Whh_n * h
    phase_commit();
    if constexpr (!DUAL) load_slab(8);
  }
  if (warp == 0) mbar_wait(bars + BAR_PHASE, 1);
  __syncthreads();
  tc_fence_after();
  PTL(4);
#pragma unroll
  for (int c = 0; c < 2; ++c) {                             // two chunks of 8 units: bounded register footprint (2 CTAs / SM)
    uint32_t gr[8], gz[8], gi[8], gh[8];
    tmem_ld8_nowait(lane_base + c * 8, gr);
    tmem_ld8_nowait(lane_base + 64 + c * 8, gz);
    tmem_ld8_nowait(lane_base + 128 + c * 8, gi);
    tmem_ld8_nowait(lane_base + 192 + c * 8, gh);
    const int off = img16_off(row, cg * 2 + c);
    float hp[8];
    unsplit8(*reinterpret_cast<const uint4*>(h_hi + off), *reinterpret_cast<const uint4*>(h_hi + IMG + off), hp);
    tmem_ld_wait();
    float hn[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int u = cg * 16 + c * 8 + e;
      const float r = sigmoidf_(__uint_as_float(gr[e]) + sp[192 + u]);
      const float z = sigmoidf_(__uint_as_float(gz[e]) + sp[256 + u]);
      const float n = tanhf_(__uint_as_float(gi[e]) + sp[320 + u] + r * (__uint_as_float(gh[e]) + sp[384 + u]));
      hn[e] = (1.0f - z) * n + z * hp[e];
    }
    uint4 h, l;
    split8_f16(hn, h, l);
    *reinterpret_cast<uint4*>(h_hi + off) = h;              // in place: this thread is the only reader of these 16 bytes
    *reinterpret_cast<uint4*>(h_hi + IMG + off) = l;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  PTL(5);
  // ---- phase 3: acc[0,64) = h_new * fc_inter^T; meanwhile commit the new state to the slot arena --------------
  if (tid_all < 32 && elect_one()) {
    phase_begin(2);
    run_slab(8, 2, 0, 0);
    phase_commit();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
    const int r = rg * 8 + rr;
    if (r < valid && s_commit[r]) {
      const unsigned char* src = h_hi + rg * 1024 + kb * 256 + sub_off;
      const uint2 a = *reinterpret_cast<const uint2*>(src), b = *reinterpret_cast<const uint2*>(src + IMG);
      const float2 a0 = h2f(a.x), a1 = h2f(a.y), b0 = h2f(b.x), b1 = h2f(b.y);
      *reinterpret_cast<float4*>(q.hstate + s_hoff[r] + kb * 16 + cc * 4) = make_float4(a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y);
    }
  }
  float yv[16];                                             // y back from its operand images, before they become the output staging
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int off = img16_off(row, cg * 2 + c);
    float t8[8];
    unsplit8(*reinterpret_cast<const uint4*>(y_hi + off), *reinterpret_cast<const uint4*>(y_hi + IMG + off), t8);
#pragma unroll
    for (int e = 0; e < 8; ++e) yv[c * 8 + e] = t8[e];
  }
  if (warp == 0) mbar_wait(bars + BAR_PHASE, 0);
  __syncthreads();
  tc_fence_after();
  PTL(6);
  {
    float v[16];
    tmem_ld16(lane_base, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += sp[448 + cg * 16 + i];
    layernorm16(v, sp + 512, sp + 576);                    // its barriers: every thread has read its y chunks
    // output tile as FP32 [128][16 chunks of 16 B] over images 0/1 (32 KB), chunk XOR-swizzled with the row
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(RA + row * 256 + (((cg * 4 + c) ^ (row & 15)) << 4)) =
          make_float4(v[c * 4] + yv[c * 4], v[c * 4 + 1] + yv[c * 4 + 1], v[c * 4 + 2] + yv[c * 4 + 2], v[c * 4 + 3] + yv[c * 4 + 3]);
  }
  tc_fence_before();
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {                             // coalesced write-out of the block output: two rows per warp instruction
    const int r = (tid >> 4) + 32 * i, ch = tid & 15;
    if (r < valid)
      *reinterpret_cast<float4*>(q.xout + (rbase + r * rstride) * C + ch * 4) = *reinterpret_cast<const float4*>(RA + r * 256 + ((ch ^ (r & 15)) << 4));
  }
  PTL(7);
  if (ovf) p.io->err[DPDF_ERRW_RANGE] = 1;
  if constexpr (PAIR) {
    tc_fence_before();
    cluster_sync_all();                                    // both CTAs have drained their accumulators; no signal is in flight
    if (warp == 0) tmem_dealloc_2cta<256>(tmem);
  } else if constexpr (DUAL) {
    tc_fence_before();
    __syncthreads();                                       // both teams have drained their accumulators
    if (tid_all < 32) tmem_dealloc<512>(tmem);
  } else {
    if (warp == 0) tmem_dealloc<256>(tmem);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// k_dprnn_post_res: the same tile math as k_dprnn_post_tc, as a PERSISTENT kernel with RESIDENT weights (throughput form,
// used when the kernel runs after its sweep).  The timeline of the streaming form (profiles/r2z_post_timeline.txt) shows
// 11-17 k of a tile's 37-44 k cycles in phase 2 - 72 MMAs that need 3 k - and 2-4 k in phases 1 and 3, all of it waiting
// for 16 KB weight slabs that are requested only when a ring buffer drains (1-2 us per bulk copy from L2, 144 KB per
// tile).  Here one CTA per SM serves one branch, pulls that branch's nine slabs ONCE (144 KB next to the 64 KB of
// operand images), takes tiles from an atomic counter and issues every phase back to back.  What two co-resident CTAs
// used to hide - the DRAM latency of a tile's inputs - is hidden explicitly: the hcat tile of the NEXT tile is loaded
// into registers (32 per thread; one CTA per SM leaves 128) right after the current one is staged, and its block input
// and state rows are warmed in L2 a whole tile ahead.
constexpr int PR_OFF_W = 4 * IMG;                        // nine resident weight slabs (processing order)
constexpr int PR_OFF_SP = PR_OFF_W + 9 * WSLAB;
constexpr int PR_OFF_RED = PR_OFF_SP + 640 * 4;
constexpr int PR_OFF_HOFF = PR_OFF_RED + 1024 * 4;
constexpr int PR_OFF_COMMIT = PR_OFF_HOFF + 128 * 8;
constexpr int PR_OFF_BAR = PR_OFF_COMMIT + 128 * 4;      // weights landed, phase result ready, TMEM slot, tile ids [2]
constexpr size_t POST_RES_SMEM = PR_OFF_BAR + 64;

__global__ void __launch_bounds__(TC_NT, 1) k_dprnn_post_res(PostTcParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* RA = smem_raw;
  unsigned char* RW = smem_raw + PR_OFF_W;
  float* sp = reinterpret_cast<float*>(smem_raw + PR_OFF_SP);
  float* red = reinterpret_cast<float*>(smem_raw + PR_OFF_RED);
  long long* s_hoff = reinterpret_cast<long long*>(smem_raw + PR_OFF_HOFF);
  int* s_commit = reinterpret_cast<int*>(smem_raw + PR_OFF_COMMIT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + PR_OFF_BAR);   // [0] weights landed, [1] phase result ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  volatile int* s_tile = reinterpret_cast<volatile int*>(tmem_slot + 1);  // [2] tile id of the current / next iteration

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3, cg = warp >> 2, row = qd * 32 + lane;
  const int bi = (int)blockIdx.x < p.ctas_erb ? 1 : 0;
  const PostTcBranch& q = p.br[bi];
  const int ntiles = bi ? p.tiles1 : p.tiles0;
  const long long nrows = (long long)p.B * q.Fp;
  uint32_t ovf = 0;

  if (tid < 64) {
    sp[tid] = q.fc_b[tid]; sp[64 + tid] = q.ln_g[tid]; sp[128 + tid] = q.ln_b[tid];
    sp[448 + tid] = q.fc2_b[tid]; sp[512 + tid] = q.ln2_g[tid]; sp[576 + tid] = q.ln2_b[tid];
  }
  if (tid < 256) sp[192 + tid] = q.bias[tid];
  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_tile[0] = atomicAdd(p.ctr + bi, 1);
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {                                          // the branch's nine slabs, once, in processing order
    mbar_expect_tx(bars, 9 * WSLAB);
    const unsigned char* fc = reinterpret_cast<const unsigned char*>(q.tc_fc_w);
    const unsigned char* gt = reinterpret_cast<const unsigned char*>(q.tc_gates);
    bulk_g2s(RW, fc, 2 * WSLAB, bars);                                       // fc_intra K halves
    const int order[6] = {0, 3, 1, 4, 2, 5};                                 // r(y), r(h), z(y), z(h), n(y), n(h)
#pragma unroll
    for (int i = 0; i < 6; ++i) bulk_g2s(RW + (2 + i) * WSLAB, gt + (size_t)order[i] * WSLAB, WSLAB, bars);
    bulk_g2s(RW + 8 * WSLAB, q.tc_fc2_w, WSLAB, bars);
  }

  const int rr = lane >> 2, cc = lane & 3;
  const int sub_off = (cc >> 1) * 128 + rr * 16 + (cc & 1) * 8;
  const uint32_t lane_base = tmem + ((uint32_t)(qd * 32) << 16) + cg * 16;
  constexpr uint32_t IDESC = idesc_f16(128, 64);
  const uint32_t a_base = smem_u32(RA), w_base = smem_u32(RW);
  auto run_slab = [&](int i, int a_img, uint32_t col, uint32_t accumulate) {
    constexpr uint64_t DESC0 = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t ah = a_base + a_img * IMG, bh = w_base + i * WSLAB;
    const uint64_t dah = DESC0 | (ah >> 4), dal = dah + (IMG >> 4), dbh = DESC0 | (bh >> 4), dbl = dbh + (WSLAB / 2 >> 4);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      umma_f16(tmem + col, dah + ks * 16, dbh + ks * 16, IDESC, accumulate);
      umma_f16(tmem + col, dal + ks * 16, dbh + ks * 16, IDESC, 1);
      umma_f16(tmem + col, dah + ks * 16, dbl + ks * 16, IDESC, 1);
      accumulate = 1;
    }
  };
  auto quad_sync = [&]() {
    if (qd == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else if (qd == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
    else if (qd == 2) asm volatile("bar.sync 3, 128;" ::: "memory");
    else asm volatile("bar.sync 4, 128;" ::: "memory");
  };
  auto layernorm16 = [&](float (&v)[16], const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    red[cg * 128 + row] = s;
    quad_sync();
    const float mean = (red[row] + red[128 + row] + red[256 + row] + red[384 + row]) * (1.0f / 64.0f);
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] -= mean; qq += v[i] * v[i]; }
    red[512 + cg * 128 + row] = qq;
    quad_sync();
    const float rstd = rsqrtf((red[512 + row] + red[640 + row] + red[768 + row] + red[896 + row]) * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] * rstd * g[cg * 16 + i] + b[cg * 16 + i];
  };
  // hcat tile (rows [base, base + 128)) -> registers: 128 blocks of 8 rows x 16 floats, 8 per warp
  auto load_hcat = [&](long long base, float4 (&v)[8]) {
    const int nvalid = (int)max(0ll, min(128ll, nrows - base));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int blk = warp + 16 * i, rg = blk >> 3, kb = blk & 7;
      const int r = rg * 8 + rr;
      v[i] = r < nvalid ? __ldg(reinterpret_cast<const float4*>(q.hcat + (base + r) * 2 * C + kb * 16 + cc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  __syncthreads();                                         // s_tile[0]
  int cur = s_tile[0];
  float4 pv[8];
  if (cur < ntiles) load_hcat((long long)cur * 128, pv);
  uint32_t ph = 0;                                         // completions of the phase barrier so far
  bool wait_w = true;
#ifdef PT_TIMELINE
#define PRL(slot) do { if (blockIdx.x == gridDim.x - 1 && tid == 64 && it == 2) p.tl[slot] = clock64(); } while (0)
#else
#define PRL(slot) do { } while (0)
#endif
  for (int it = 0; cur < ntiles; ++it) {
    PRL(0);
    const long long rbase = (long long)cur * 128;
    const int valid = (int)min(128ll, nrows - rbase);
    if (tid == 0) s_tile[(it + 1) & 1] = atomicAdd(p.ctr + bi, 1);
    if (tid < 128) {
      long long off = 0;
      int commit = 0;
      if (tid < valid) {
        const long long r = rbase + tid;
        const int b = (int)(r / q.Fp), f = (int)(r % q.Fp);
        off = (long long)io_slot(p.io, b) * q.per_slot + (long long)f * C;
        commit = (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) ? 0 : 1;
      }
      s_hoff[tid] = off;
      s_commit[tid] = commit;
    }
    // ---- stage the prefetched hcat tile as two K=64 operand image pairs (split on the fly) --------------------------
    // (the barrier that ended the previous tile's write-out ordered its reads of the output staging before these writes)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int blk = warp + 16 * i, rg = blk >> 3, kb = blk & 7;
      uint2 h, l;
      split2_f16(pv[i].x, pv[i].y, h.x, l.x);
      split2_f16(pv[i].z, pv[i].w, h.y, l.y);
      unsigned char* dst = RA + (kb >> 2) * 2 * IMG + rg * 1024 + (kb & 3) * 256 + sub_off;
      *reinterpret_cast<uint2*>(dst) = h;
      *reinterpret_cast<uint2*>(dst + IMG) = l;
    }
    float4 xv[4];
    {
      const float* xr = q.xin + (rbase + row) * C + cg * 16;
#pragma unroll
      for (int c = 0; c < 4; ++c) xv[c] = row < valid ? __ldg(reinterpret_cast<const float4*>(xr + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();                                       // images, s_hoff and the next tile id are visible
    tc_fence_after();
    const int nxt = s_tile[(it + 1) & 1];
    PRL(1);

    // ---- phase 1: acc[0,64) = hcat * fc_intra^T ---------------------------------------------------------------
    if (warp == 0 && elect_one()) {
      if (wait_w) mbar_wait(bars, 0);                      // the resident weights have landed (first tile only)
      run_slab(0, 0, 0, 0);
      run_slab(1, 2, 0, 1);
      umma_commit(bars + 1);
    }
    wait_w = false;
    // the next tile: its hcat rows into registers, its block input and state rows into L2 (a whole tile ahead)
    if (nxt < ntiles) {
      const long long nbase = (long long)nxt * 128;
      load_hcat(nbase, pv);
      const int nvalid = (int)min(128ll, nrows - nbase);
      if (tid == 0) {
        bulk_prefetch_l2(q.xin + nbase * C, (uint32_t)nvalid * C * 4);
      } else if (tid >= 128 && tid < 128 + nvalid) {
        const long long r = nbase + (tid - 128);
        const int b = (int)(r / q.Fp), f = (int)(r % q.Fp);
        const float* hp = q.hstate + (long long)io_slot(p.io, b) * q.per_slot + (long long)f * C;
        prefetch_l2(hp);
        prefetch_l2(hp + 32);
      }
    }
    float4 hv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
      const int r = rg * 8 + rr;
      hv[i] = r < valid ? __ldg(reinterpret_cast<const float4*>(q.hstate + s_hoff[r] + kb * 16 + cc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (warp == 0) mbar_wait(bars + 1, ph & 1);
    ++ph;
    __syncthreads();
    tc_fence_after();
    PRL(2);

    unsigned char* y_hi = RA;
    unsigned char* h_hi = RA + 2 * IMG;
    {
      float v[16];
      tmem_ld16(lane_base, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += sp[cg * 16 + i];
      layernorm16(v, sp + 64, sp + 128);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float y[8];
        y[0] = v[c * 8 + 0] + xv[2 * c].x; y[1] = v[c * 8 + 1] + xv[2 * c].y; y[2] = v[c * 8 + 2] + xv[2 * c].z; y[3] = v[c * 8 + 3] + xv[2 * c].w;
        y[4] = v[c * 8 + 4] + xv[2 * c + 1].x; y[5] = v[c * 8 + 5] + xv[2 * c + 1].y; y[6] = v[c * 8 + 6] + xv[2 * c + 1].z; y[7] = v[c * 8 + 7] + xv[2 * c + 1].w;
        uint4 h, l;
        split8_f16(y, h, l);
        ovf |= f16_nonfinite(h.x) | f16_nonfinite(h.y) | f16_nonfinite(h.z) | f16_nonfinite(h.w);
        const int off = img16_off(row, cg * 2 + c);
        *reinterpret_cast<uint4*>(y_hi + off) = h;
        *reinterpret_cast<uint4*>(y_hi + IMG + off) = l;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
        uint2 h, l;
        split2_f16(hv[i].x, hv[i].y, h.x, l.x);
        split2_f16(hv[i].z, hv[i].w, h.y, l.y);
        unsigned char* dst = h_hi + rg * 1024 + kb * 256 + sub_off;
        *reinterpret_cast<uint2*>(dst) = h;
        *reinterpret_cast<uint2*>(dst + IMG) = l;
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    PRL(3);
    // ---- phase 2: GRU gate pre-activations: r [0,64) z [64,128) in [128,192) hn [192,256) -----------------------
    if (warp == 0 && elect_one()) {
      run_slab(2, 0, 0, 0);
      run_slab(3, 2, 0, 1);
      run_slab(4, 0, 64, 0);
      run_slab(5, 2, 64, 1);
      run_slab(6, 0, 128, 0);
      run_slab(7, 2, 192, 0);
      umma_commit(bars + 1);
    }
    if (warp == 0) mbar_wait(bars + 1, ph & 1);
    ++ph;
    __syncthreads();
    tc_fence_after();
    PRL(4);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t gr[8], gz[8], gi[8], gh[8];
      tmem_ld8_nowait(lane_base + c * 8, gr);
      tmem_ld8_nowait(lane_base + 64 + c * 8, gz);
      tmem_ld8_nowait(lane_base + 128 + c * 8, gi);
      tmem_ld8_nowait(lane_base + 192 + c * 8, gh);
      const int off = img16_off(row, cg * 2 + c);
      float hp[8];
      unsplit8(*reinterpret_cast<const uint4*>(h_hi + off), *reinterpret_cast<const uint4*>(h_hi + IMG + off), hp);
      tmem_ld_wait();
      float hn[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int u = cg * 16 + c * 8 + e;
        const float r = sigmoidf_(__uint_as_float(gr[e]) + sp[192 + u]);
        const float z = sigmoidf_(__uint_as_float(gz[e]) + sp[256 + u]);
        const float n = tanhf_(__uint_as_float(gi[e]) + sp[320 + u] + r * (__uint_as_float(gh[e]) + sp[384 + u]));
        hn[e] = (1.0f - z) * n + z * hp[e];
      }
      uint4 h, l;
      split8_f16(hn, h, l);
      *reinterpret_cast<uint4*>(h_hi + off) = h;
      *reinterpret_cast<uint4*>(h_hi + IMG + off) = l;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    PRL(5);

    // ---- phase 3: acc[0,64) = h_new * fc_inter^T; meanwhile commit the new state to the slot arena ----------------
    if (warp == 0 && elect_one()) {
      run_slab(8, 2, 0, 0);
      umma_commit(bars + 1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
      const int r = rg * 8 + rr;
      if (r < valid && s_commit[r]) {
        const unsigned char* src = h_hi + rg * 1024 + kb * 256 + sub_off;
        const uint2 a = *reinterpret_cast<const uint2*>(src), b = *reinterpret_cast<const uint2*>(src + IMG);
        const float2 a0 = h2f(a.x), a1 = h2f(a.y), b0 = h2f(b.x), b1 = h2f(b.y);
        *reinterpret_cast<float4*>(q.hstate + s_hoff[r] + kb * 16 + cc * 4) = make_float4(a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y);
      }
    }
    float yv[16];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int off = img16_off(row, cg * 2 + c);
      float t8[8];
      unsplit8(*reinterpret_cast<const uint4*>(y_hi + off), *reinterpret_cast<const uint4*>(y_hi + IMG + off), t8);
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[c * 8 + e] = t8[e];
    }
    if (warp == 0) mbar_wait(bars + 1, ph & 1);
    ++ph;
    __syncthreads();
    tc_fence_after();
    PRL(6);
    {
      float v[16];
      tmem_ld16(lane_base, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += sp[448 + cg * 16 + i];
      layernorm16(v, sp + 512, sp + 576);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(RA + row * 256 + (((cg * 4 + c) ^ (row & 15)) << 4)) =
            make_float4(v[c * 4] + yv[c * 4], v[c * 4 + 1] + yv[c * 4 + 1], v[c * 4 + 2] + yv[c * 4 + 2], v[c * 4 + 3] + yv[c * 4 + 3]);
    }
    tc_fence_before();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (tid >> 4) + 32 * i, ch = tid & 15;
      if (r < valid)
        *reinterpret_cast<float4*>(q.xout + (rbase + r) * C + ch * 4) = *reinterpret_cast<const float4*>(RA + r * 256 + ((ch ^ (r & 15)) << 4));
    }
    __syncthreads();                                       // the output staging is free again before the next tile's images are written
    PRL(7);
    cur = nxt;
  }
  if (ovf) p.io->err[DPDF_ERRW_RANGE] = 1;
  if (tid == 0) {
    if (wait_w) mbar_wait(bars, 0);                        // a CTA that found no tile still owns nine bulk copies in flight
    const int done = atomicAdd(p.ctr + 2, 1);
    if (done == (int)gridDim.x - 1) {                      // every CTA of this launch has drawn its last tile id: reset for the next launch
      p.ctr[0] = 0; p.ctr[1] = 0; p.ctr[2] = 0;
      __threadfence();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

void launch_dprnn_post_tc(Engine& e, int blk, int B, cudaStream_t st) {
  PostTcParams p{};
  p.io = e.io_dev;
  p.B = B;
  auto fill = [&](PostTcBranch& b, const DprnnW& w, const float* hcat, const float* xin, float* xout, float* hstate, int Fp) {
    b.hcat = hcat; b.xin = xin; b.xout = xout; b.Fp = Fp;
    b.per_slot = (long long)e.d.N * Fp * C;
    b.hstate = hstate + (size_t)blk * Fp * C;
    b.tc_fc_w = w.tc_fc_w; b.tc_gates = w.tc_gates; b.tc_fc2_w = w.tc_fc2_w;
    b.fc_b = w.fc_b; b.ln_g = w.ln_g; b.ln_b = w.ln_b; b.bias = w.r_bias;
    b.fc2_b = w.fc2_b; b.ln2_g = w.ln2_g; b.ln2_b = w.ln2_b;
  };
  fill(p.br[0], e.w.dprnn_df[blk], e.sc.hcat_d, e.sc.c1, e.sc.c1, e.st.inter_df, NDF / 2);
  fill(p.br[1], e.w.dprnn_erb[blk], e.sc.hcat_e, blk == 0 ? e.sc.e3 : e.sc.xe, e.sc.xe, e.st.inter_erb, e.d.fe[3]);
  // CTA pairs (cta_group::2, Engine::post_pair): 1 = always, 2 = only when the kernel runs after its sweep (a pair needs
  // both SMs of a TPC, which the CTAs of a concurrent sweep fragment)
  const bool pair = e.post_pair == 1 || (e.post_pair == 2 && !e.overlap_now);
  // two tiles per 1024-thread CTA sharing a 5-slab weight ring (Engine::post_dual): 1 = always, 2 = only after the sweep
  const bool dual = !pair && (e.post_dual == 1 || (e.post_dual == 2 && !e.overlap_now));
  auto even = [&](int t) { return (pair || dual) ? (t + 1) & ~1 : t; };
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(dual ? 2 * TC_NT : TC_NT);
  cfg.dynamicSmemBytes = dual ? PostLayout<2>::SMEM : (pair ? PostLayout<1>::SMEM : PostLayout<0>::SMEM);
  auto launch = [&]() {
    if (dual) { cfg.gridDim.x /= 2; cudaLaunchKernelEx(&cfg, k_dprnn_post_tc<2>, p); }
    else if (pair) cudaLaunchKernelEx(&cfg, k_dprnn_post_tc<1>, p);
    else cudaLaunchKernelEx(&cfg, k_dprnn_post_tc<0>, p);
  };
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (pair) {
    attr[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
    attr[cfg.numAttrs].val.clusterDim.x = 2;
    attr[cfg.numAttrs].val.clusterDim.y = 1;
    attr[cfg.numAttrs].val.clusterDim.z = 1;
    ++cfg.numAttrs;
  }
  auto pdl_attr = [&]() {
    attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    ++cfg.numAttrs;
  };
  if (!e.overlap_now && e.post_res) {
    // persistent form with resident weights: one CTA per SM, each dedicated to one branch (the branches have different
    // weights), CTAs split in proportion to the branches' tile counts, tiles drawn from per-lane atomic counters
    p.tiles0 = (int)(((long long)B * (NDF / 2) + 127) / 128);
    p.tiles1 = (int)(((long long)B * e.d.fe[3] + 127) / 128);
    const int G = std::min(e.num_sms, p.tiles0 + p.tiles1);
    p.ctas_erb = G < 2 ? 0 : std::min(G - 1, std::max(1, (int)((long long)G * p.tiles1 / (p.tiles0 + p.tiles1))));
    p.ctr = e.post_ctr_dev + 4 * e.cur_lane;
    cfg.dynamicSmemBytes = POST_RES_SMEM;
    cfg.gridDim = dim3((unsigned)G);
    cfg.numAttrs = 0;
    if (e.pdl_now && !e.pdl_first) pdl_attr();
    cudaLaunchKernelEx(&cfg, k_dprnn_post_res, p);
    return;
  }
  if (!e.overlap_now) {
    p.tiles0 = even((int)(((long long)B * (NDF / 2) + 127) / 128));
    p.tiles1 = even((int)(((long long)B * e.d.fe[3] + 127) / 128));
    p.pf_dist = e.post_pf * e.num_sms;
    cfg.gridDim = dim3((unsigned)(p.tiles0 + p.tiles1));
    if (e.pdl_now && !e.pdl_first) pdl_attr();
    launch();
    return;
  }
  // Overlapped with the sweep that feeds it: launched as a programmatic dependent of the intra kernel (it may start as
  // soon as every intra CTA has started, i.e. none of them can be starved of an SM by a waiting post CTA) and
  // synchronised with it through the per-CTA progress counters only.
  p.progress = e.progress_dev + (size_t)e.cur_lane * 4 * e.progress_tiles;
  p.stiles = (B + 127) / 128;
  p.dup = intra_tc_dup(e, B);
  p.ptiles = (B * p.dup + 127) / 128;
  p.dup_e = intra_tc_dup_erb(e, B);
  p.ptiles_e = (B * p.dup_e + 127) / 128;
  p.tiles1 = even(p.stiles * e.d.fe[3]);
  p.tiles0 = even(p.stiles * (NDF / 2));
  cfg.gridDim = dim3((unsigned)(p.tiles0 + p.tiles1));
  pdl_attr();
  launch();
}

void init_dprnn_tc_kernels() {
  cudaFuncSetAttribute(k_dprnn_post_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PostLayout<0>::SMEM);
  cudaFuncSetAttribute(k_dprnn_post_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PostLayout<1>::SMEM);
  cudaFuncSetAttribute(k_dprnn_post_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PostLayout<2>::SMEM);
  cudaFuncSetAttribute(k_dprnn_post_res, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POST_RES_SMEM);
}

}  // namespace dpdf
