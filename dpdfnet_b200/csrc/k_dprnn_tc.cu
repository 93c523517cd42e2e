// k_dprnn_post_tc: the position-parallel half of a DPRNN block on the 5th-generation tensor cores.
//
// Same math as k_dprnn_post (fc_intra + LayerNorm + residual, inter-frame GRUCell, fc_inter + LayerNorm +
// residual; layers.py:178-196) for a 128-row tile, but every matrix product is a chain of
// tcgen05.mma.kind::tf32 instructions with the FP32 accumulators in tensor memory.  FP32 accuracy is kept
// with the error-compensated 3xTF32 scheme: every operand is split as x = hi + lo with hi, lo exactly
// representable in TF32 (cvt.rna), and D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (the dropped lo*lo term and
// the rounding of lo are ~2^-24 relative).  Weights are split and laid out on the host
// (weights.py:umma_operand); activations are split on the fly while they are staged into shared memory.
//
// Operand layout: K-major, SWIZZLE_NONE ("interleave"): 8 rows x 16 B core matrices (128 B contiguous), core
// matrices adjacent in K are LBO = 128 B apart, 8-row groups are SBO = (K/4)*128 B apart.
// One CTA = 128 threads; thread t owns TMEM lane t = tile row t, so LayerNorm and the GRU gate math are
// row-local register code straight out of tcgen05.ld (no shuffles).
#include "engine.h"

namespace dpdf {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes) {
  // start address [0,14) >>4, LBO [16,30) >>4 (=128 B), SBO [32,46) >>4, descriptor version 1 at bit 46, SWIZZLE_NONE
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
  hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
  lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// float offset of element (r, k) of a [rows][K] operand image
template <int K>
__device__ __forceinline__ int core_off(int r, int k) { return (r >> 3) * (K * 8) + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

// Stage a [128][K] global tile into hi / lo operand images.  A warp moves an 8-row x 16-float block per
// iteration: 64 B contiguous per row from HBM/L2, 512 contiguous bytes (four core matrices) into smem.
template <int K, typename RowPtr>
__device__ __forceinline__ void stage_split(float* hi, float* lo, int valid, RowPtr row_ptr, int warp, int lane) {
  const int rr = lane >> 2, cc = lane & 3;
  constexpr int KB = K / 16;
#pragma unroll 4
  for (int blk = warp; blk < 16 * KB; blk += 4) {
    const int rg = blk / KB, kb = blk % KB;
    const int r = rg * 8 + rr, k = kb * 16 + cc * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < valid) v = __ldg(reinterpret_cast<const float4*>(row_ptr(r) + k));
    float4 h, l;
    split4(v, h, l);
    const int off = rg * (K * 8) + (kb * 4 + cc) * 32 + rr * 4;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// Copy a [128][64] plain-float operand image (optionally hi + lo) back to row-major global memory, coalesced.
template <typename RowPtr, typename Pred>
__device__ __forceinline__ void unstage64(const float* a, const float* b, int valid, RowPtr row_ptr, Pred pred, int warp, int lane) {
  const int rr = lane >> 2, cc = lane & 3;
  for (int blk = warp; blk < 16 * 4; blk += 4) {
    const int rg = blk >> 2, kb = blk & 3;
    const int r = rg * 8 + rr, k = kb * 16 + cc * 4;
    if (r >= valid || !pred(r)) continue;
    const int off = rg * 512 + (kb * 4 + cc) * 32 + rr * 4;
    float4 v = *reinterpret_cast<const float4*>(a + off);
    if (b) {
      const float4 w = *reinterpret_cast<const float4*>(b + off);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    *reinterpret_cast<float4*>(row_ptr(r) + k) = v;
  }
}

__device__ __forceinline__ void layernorm64(float (&v)[64], const float* g, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += v[i];
  const float mean = s * (1.0f / 64.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) { v[i] -= mean; q += v[i] * v[i]; }
  const float rstd = rsqrtf(q * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = v[i] * rstd * g[i] + b[i];
}

}  // namespace

struct PostTcBranch {
  const float* hcat;      // [rows][128]
  const float* xin;       // [rows][64]
  float* xout;            // [rows][64]
  float* hstate;          // inter-GRU state of this block: + slot*per_slot + f*64
  long long per_slot;
  int Fp;
  const float *tc_fc_w, *tc_gates, *tc_fc2_w;         // operand images (hi | lo)
  const float *fc_b, *ln_g, *ln_b, *bias, *fc2_b, *ln2_g, *ln2_b;
};
struct PostTcParams {
  const IoDesc* io;
  PostTcBranch br[2];
  int tiles0, B;
};

constexpr int TC_A = 32768;       // floats: activation operand region (128 KB)
constexpr int TC_W = 16384;       // floats: weight operand region (64 KB)
constexpr size_t POST_TC_SMEM = (size_t)(TC_A + TC_W + 640) * sizeof(float) + 128 * sizeof(long long) + 128 * sizeof(int) + 64;

__global__ void __launch_bounds__(128, 1) k_dprnn_post_tc(PostTcParams p) {
  extern __shared__ __align__(128) float smem[];
  float* RA = smem;
  float* RW = RA + TC_A;
  float* sp = RW + TC_W;
  long long* s_hoff = reinterpret_cast<long long*>(sp + 640);
  int* s_commit = reinterpret_cast<int*>(s_hoff + 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_commit + 128);     // [0] phase barrier, [1],[2] weight-slab buffers
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bi = (int)blockIdx.x >= p.tiles0 ? 1 : 0;
  const PostTcBranch& q = p.br[bi];
  const long long row0 = (long long)(blockIdx.x - (bi ? p.tiles0 : 0)) * 128;
  const long long nrows = (long long)p.B * q.Fp;
  const int valid = (int)min((long long)128, nrows - row0);

  {
    long long off = 0;
    int commit = 0;
    if (tid < valid) {
      const long long row = row0 + tid;
      const int b = (int)(row / q.Fp), f = (int)(row % q.Fp);
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
  sp[192 + tid] = q.bias[tid];
  sp[320 + tid] = q.bias[128 + tid];
  if (tid == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // ---- stage phase-1 operands: hcat tile (split on the fly) and the fc_intra operand image -------------------
  {
    const float* hc = q.hcat + row0 * 2 * C;
    stage_split<128>(RA, RA + 16384, valid, [&](int r) { return hc + (size_t)r * 2 * C; }, warp, lane);
    for (int i = tid; i < TC_W / 4; i += 128) cp_async16(RW + i * 4, q.tc_fc_w + i * 4);
    cp_async_commit();
    cp_async_wait<0>();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // M=128, N=64, tf32, f32 acc
  const uint32_t a_base = smem_u32(RA), w_base = smem_u32(RW);

  // D[tmem_col .. +64) (+)= A[128][K] * W[64][K]^T, three TF32 passes per 8-wide k-step (single issuing thread)
  auto gemm3 = [&](uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int K, uint32_t col, uint32_t accumulate) {
    const uint32_t sbo = (uint32_t)K * 32;            // (K/4) * 128 B
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t dah = umma_desc(a_hi + ks * 256, sbo), dal = umma_desc(a_lo + ks * 256, sbo);
      const uint64_t dbh = umma_desc(b_hi + ks * 256, sbo), dbl = umma_desc(b_lo + ks * 256, sbo);
      umma_tf32(tmem + col, dah, dbh, IDESC, accumulate);
      umma_tf32(tmem + col, dal, dbh, IDESC, 1);
      umma_tf32(tmem + col, dah, dbl, IDESC, 1);
      accumulate = 1;
    }
  };

  // ---- phase 1: acc[0..64) = hcat * fc_intra^T --------------------------------------------------------------
  if (tid == 0) {
    gemm3(a_base, a_base + 65536, w_base, w_base + 32768, 128, 0, 0);
    umma_commit(bars);
  }
  mbar_wait(bars, 0);
  tc_fence_after();

  float* y_hi = RA;                 // [128][64] operand images, K = 64
  float* y_lo = RA + 8192;
  float* h_hi = RA + 16384;
  float* h_lo = RA + 24576;
  {
    float v[64];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float t[16];
      tmem_ld16(lane_base + c * 16, t);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[c * 16 + i] = t[i] + sp[c * 16 + i];
    }
    layernorm64(v, sp + 64, sp + 128);
    const float* xr = q.xin + (size_t)(row0 + tid) * C;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < valid) x = __ldg(reinterpret_cast<const float4*>(xr + c * 4));
      const float4 y = make_float4(v[c * 4] + x.x, v[c * 4 + 1] + x.y, v[c * 4 + 2] + x.z, v[c * 4 + 3] + x.w);
      float4 h, l;
      split4(y, h, l);
      const int off = core_off<64>(tid, c * 4);
      *reinterpret_cast<float4*>(y_hi + off) = h;       // all MMAs that read the hcat image have completed
      *reinterpret_cast<float4*>(y_lo + off) = l;
    }
  }
  __syncthreads();                                        // every thread is done reading the fc_intra image / hcat rows
  // h_prev tile (split) + first two gate slabs
  stage_split<64>(h_hi, h_lo, valid, [&](int r) { return q.hstate + s_hoff[r]; }, warp, lane);
  auto load_slab = [&](int pidx, int buf) {               // processing order -> slab id in the blob (Wih r,z,n | Whh r,z,n)
    const int slab = (pidx & 1) ? 3 + (pidx >> 1) : (pidx >> 1);
    const float* src = q.tc_gates + (size_t)slab * 8192;
    float* dst = RW + buf * 8192;
    for (int i = tid; i < 2048; i += 128) cp_async16(dst + i * 4, src + i * 4);
    cp_async_commit();
  };
  load_slab(0, 0);
  load_slab(1, 1);
  cp_async_wait<0>();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- phase 2: inter-frame GRU gate pre-activations in TMEM: r [64,128) z [128,192) in [192,256) hn [256,320) ---
  const uint32_t yh = a_base, yl = a_base + 32768, hh = a_base + 65536, hl = a_base + 98304;
  auto issue_slab = [&](int pidx) {
    if (tid == 0) {
      const int buf = pidx & 1;
      const bool use_h = (pidx & 1) != 0;
      const int gate = pidx >> 1;                                   // 0 r, 1 z, 2 n
      const uint32_t col = gate < 2 ? 64 + 64 * gate : (use_h ? 256 : 192);
      const uint32_t acc = (gate < 2 && use_h) ? 1u : 0u;
      const uint32_t wb = w_base + buf * 32768;
      gemm3(use_h ? hh : yh, use_h ? hl : yl, wb, wb + 16384, 64, col, acc);
      umma_commit(bars + 1 + buf);
    }
  };
  issue_slab(0);
  issue_slab(1);
  for (int pidx = 0; pidx < 4; ++pidx) {
    mbar_wait(bars + 1 + (pidx & 1), (pidx >> 1) & 1);              // slab pidx consumed -> its buffer is free
    load_slab(pidx + 2, pidx & 1);
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    issue_slab(pidx + 2);
  }
  mbar_wait(bars + 1, 0);
  mbar_wait(bars + 2, 0);
  tc_fence_after();
  // prefetch the fc_inter operand image while the gates are evaluated (both slab buffers are free now)
  for (int i = tid; i < 2048; i += 128) cp_async16(RW + i * 4, q.tc_fc2_w + i * 4);
  cp_async_commit();

#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float gr[16], gz[16], gi[16], gh[16];
    tmem_ld16(lane_base + 64 + c * 16, gr);
    tmem_ld16(lane_base + 128 + c * 16, gz);
    tmem_ld16(lane_base + 192 + c * 16, gi);
    tmem_ld16(lane_base + 256 + c * 16, gh);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const int k = c * 16 + c4 * 4;
      const int off = core_off<64>(tid, k);
      const float4 ph = *reinterpret_cast<const float4*>(h_hi + off);
      const float4 pl = *reinterpret_cast<const float4*>(h_lo + off);
      const float hp[4] = {ph.x + pl.x, ph.y + pl.y, ph.z + pl.z, ph.w + pl.w};
      float hn[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int u = k + e, i = c4 * 4 + e;
        const float r = sigmoidf_(gr[i] + sp[192 + u]);
        const float z = sigmoidf_(gz[i] + sp[256 + u]);
        const float n = tanhf_(gi[i] + sp[320 + u] + r * (gh[i] + sp[384 + u]));
        hn[e] = (1.0f - z) * n + z * hp[e];
      }
      float4 h, l;
      split4(make_float4(hn[0], hn[1], hn[2], hn[3]), h, l);
      *reinterpret_cast<float4*>(h_hi + off) = h;      // in place: this thread is the only reader of its row
      *reinterpret_cast<float4*>(h_lo + off) = l;
    }
  }
  cp_async_wait<0>();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- phase 3: acc[320..384) = h_new * fc_inter^T; meanwhile commit the new state to the slot arena ----------
  if (tid == 0) {
    gemm3(hh, hl, w_base, w_base + 16384, 64, 320, 0);
    umma_commit(bars);
  }
  unstage64(h_hi, h_lo, valid, [&](int r) { return q.hstate + s_hoff[r]; }, [&](int r) { return s_commit[r] != 0; }, warp, lane);
  mbar_wait(bars, 1);
  tc_fence_after();
  {
    float v[64];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float t[16];
      tmem_ld16(lane_base + 320 + c * 16, t);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[c * 16 + i] = t[i] + sp[448 + c * 16 + i];
    }
    layernorm64(v, sp + 512, sp + 576);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int off = core_off<64>(tid, c * 4);
      const float4 a = *reinterpret_cast<const float4*>(y_hi + off);
      const float4 b = *reinterpret_cast<const float4*>(y_lo + off);
      *reinterpret_cast<float4*>(y_hi + off) =
          make_float4(v[c * 4] + a.x + b.x, v[c * 4 + 1] + a.y + b.y, v[c * 4 + 2] + a.z + b.z, v[c * 4 + 3] + a.w + b.w);
    }
  }
  tc_fence_before();
  __syncthreads();
  float* xo = q.xout + row0 * C;
  unstage64(y_hi, nullptr, valid, [&](int r) { return xo + (size_t)r * C; }, [](int) { return true; }, warp, lane);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
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
  p.tiles0 = (int)(((long long)B * (NDF / 2) + 127) / 128);
  const int tiles1 = (int)(((long long)B * e.d.fe[3] + 127) / 128);
  k_dprnn_post_tc<<<p.tiles0 + tiles1, 128, POST_TC_SMEM, st>>>(p);
}

void init_dprnn_tc_kernels() {
  cudaFuncSetAttribute(k_dprnn_post_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POST_TC_SMEM);
}

}  // namespace dpdf
