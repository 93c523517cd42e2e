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
#include "tc_common.cuh"

namespace dpdf {

namespace {

using namespace tc;

// float offset of element (r, k) of a [128][64] operand image (K-major, SWIZZLE_NONE, SBO = 2048 B)
__device__ __forceinline__ int core_off64(int r, int k) { return (r >> 3) * 512 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

constexpr int TC_NT = 512;        // 16 warps: warp w -> TMEM lane quadrant w & 3, 16-column group w >> 2
constexpr int IMG = 8192;         // floats of one [128][64] operand image (32 KB)

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
};

constexpr size_t POST_TC_SMEM = (size_t)(4 * IMG + 2 * IMG + 640 + 1024) * sizeof(float) + 128 * sizeof(long long) +
                                128 * sizeof(int) + 64;

// One CTA = one 128-row tile.  Thread 0 is the single MMA issuer and streams the nine 32 KB weight slabs of the
// block (fc_intra K-halves, six GRU gate slabs, fc_inter) through a two-buffer ring with 1-D bulk copies that
// complete on mbarriers, two slabs ahead of the tensor core; the other 511 threads never wait for weights.
__global__ void __launch_bounds__(TC_NT, 1) k_dprnn_post_tc(PostTcParams p) {
  extern __shared__ __align__(128) float smem[];
  float* RA = smem;                                 // four activation operand images
  float* RW = RA + 4 * IMG;                         // two weight slab buffers
  float* sp = RW + 2 * IMG;                         // small parameters
  float* red = sp + 640;                            // [2][4][128] LayerNorm partials
  long long* s_hoff = reinterpret_cast<long long*>(red + 1024);
  int* s_commit = reinterpret_cast<int*>(s_hoff + 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_commit + 128);   // [0,1] slab full, [2,3] slab consumed, [4] phase result ready
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3, cg = warp >> 2, row = qd * 32 + lane;
  const int bi = (int)blockIdx.x >= p.tiles0 ? 1 : 0;
  const PostTcBranch& q = p.br[bi];
  const long long row0 = (long long)(blockIdx.x - (bi ? p.tiles0 : 0)) * 128;
  const long long nrows = (long long)p.B * q.Fp;
  const int valid = (int)min((long long)128, nrows - row0);

  if (tid < 128) {
    long long off = 0;
    int commit = 0;
    if (tid < valid) {
      const long long r = row0 + tid;
      const int b = (int)(r / q.Fp), f = (int)(r % q.Fp);
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
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();                                       // barriers initialised, s_hoff visible

  // ---- weight slab ring (thread 0 only) --------------------------------------------------------------------
  auto slab_src = [&](int i) -> const float* {
    if (i < 2) return q.tc_fc_w + (size_t)i * IMG;
    if (i == 8) return q.tc_fc2_w;
    const int pidx = i - 2;                                // processing order r(y),r(h),z(y),z(h),n(y),n(h)
    return q.tc_gates + (size_t)((pidx & 1) ? 3 + (pidx >> 1) : (pidx >> 1)) * IMG;
  };
  auto load_slab = [&](int i) {
    const int buf = i & 1;
    if (i >= 2) mbar_wait(bars + 2 + buf, ((i - 2) >> 1) & 1);      // MMAs of the previous tenant have completed
    mbar_expect_tx(bars + buf, IMG * 4);
    bulk_g2s(RW + buf * IMG, slab_src(i), IMG * 4, bars + buf);
  };
  if (tid == 0) { load_slab(0); load_slab(1); }

  // ---- stage the hcat tile as two K=64 operand image pairs (split on the fly) --------------------------------
  const int rr = lane >> 2, cc = lane & 3;
  {
    const float* hc = q.hcat + row0 * 2 * C;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                          // 128 blocks of 8 rows x 16 floats, 8 per warp, all in flight
      const int blk = warp + 16 * i, rg = blk >> 3, kb = blk & 7;
      const int r = rg * 8 + rr;
      v[i] = r < valid ? __ldg(reinterpret_cast<const float4*>(hc + (size_t)r * 2 * C + kb * 16 + cc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int blk = warp + 16 * i, rg = blk >> 3, kb = blk & 7;
      float4 h, l;
      split4(v[i], h, l);
      const int off = (kb >> 2) * 2 * IMG + rg * 512 + ((kb & 3) * 4 + cc) * 32 + rr * 4;
      *reinterpret_cast<float4*>(RA + off) = h;
      *reinterpret_cast<float4*>(RA + IMG + off) = l;
    }
  }
  // prefetch what the first epilogue needs while the tensor core works: residual input and h_prev
  float4 xv[4], hv[4];
  {
    const float* xr = q.xin + (size_t)(row0 + row) * C + cg * 16;
#pragma unroll
    for (int c = 0; c < 4; ++c) xv[c] = row < valid ? __ldg(reinterpret_cast<const float4*>(xr + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {                          // 64 blocks, 4 per warp
      const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
      const int r = rg * 8 + rr;
      hv[i] = r < valid ? __ldg(reinterpret_cast<const float4*>(q.hstate + s_hoff[r] + kb * 16 + cc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(qd * 32) << 16) + cg * 16;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // M=128, N=64, tf32, f32 acc
  const uint32_t a_base = smem_u32(RA), w_base = smem_u32(RW);

  // slab i: D[col .. col+64) (+)= A[128][64] * W_i[64][64]^T, three TF32 passes per 8-wide k-step
  auto run_slab = [&](int i, int a_img, uint32_t col, uint32_t accumulate) {
    const int buf = i & 1;
    mbar_wait(bars + buf, (i >> 1) & 1);                   // slab landed (async proxy write -> async proxy read)
    const uint32_t ah = a_base + a_img * (IMG * 4), al = ah + IMG * 4;
    const uint32_t bh = w_base + buf * (IMG * 4), bl = bh + IMG * 2;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t dah = umma_desc(ah + ks * 256, 2048), dal = umma_desc(al + ks * 256, 2048);
      const uint64_t dbh = umma_desc(bh + ks * 256, 2048), dbl = umma_desc(bl + ks * 256, 2048);
      umma_tf32(tmem + col, dah, dbh, IDESC, accumulate);
      umma_tf32(tmem + col, dal, dbh, IDESC, 1);
      umma_tf32(tmem + col, dah, dbl, IDESC, 1);
      accumulate = 1;
    }
    umma_commit(bars + 2 + buf);
  };

  // ---- phase 1: acc[0,64) = hcat * fc_intra^T ---------------------------------------------------------------
  if (tid == 0) {
    run_slab(0, 0, 0, 0);
    run_slab(1, 2, 0, 1);
    umma_commit(bars + 4);
    load_slab(2);
    load_slab(3);
  }
  mbar_wait(bars + 4, 0);
  tc_fence_after();

  float* y_hi = RA;
  float* y_lo = RA + IMG;
  float* h_hi = RA + 2 * IMG;
  float* h_lo = RA + 3 * IMG;
  // row statistics over 64 columns held by the four column-group warps of a lane quadrant
  auto layernorm16 = [&](float (&v)[16], const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    red[cg * 128 + row] = s;
    __syncthreads();
    const float mean = (red[row] + red[128 + row] + red[256 + row] + red[384 + row]) * (1.0f / 64.0f);
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] -= mean; qq += v[i] * v[i]; }
    red[512 + cg * 128 + row] = qq;
    __syncthreads();
    const float rstd = rsqrtf((red[512 + row] + red[640 + row] + red[768 + row] + red[896 + row]) * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] * rstd * g[cg * 16 + i] + b[cg * 16 + i];
  };
  {
    float v[16];
    tmem_ld16(lane_base, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += sp[cg * 16 + i];
    layernorm16(v, sp + 64, sp + 128);
#pragma unroll
    for (int c = 0; c < 4; ++c) {                           // y = LN(...) + x, staged as the hi / lo operand of the gate GEMMs
      const float4 y = make_float4(v[c * 4] + xv[c].x, v[c * 4 + 1] + xv[c].y, v[c * 4 + 2] + xv[c].z, v[c * 4 + 3] + xv[c].w);
      float4 h, l;
      split4(y, h, l);
      const int off = core_off64(row, cg * 16 + c * 4);
      *reinterpret_cast<float4*>(y_hi + off) = h;
      *reinterpret_cast<float4*>(y_lo + off) = l;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                           // h_prev tile (loaded before the wait) -> hi / lo images
      const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
      float4 h, l;
      split4(hv[i], h, l);
      const int off = rg * 512 + (kb * 4 + cc) * 32 + rr * 4;
      *reinterpret_cast<float4*>(h_hi + off) = h;
      *reinterpret_cast<float4*>(h_lo + off) = l;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- phase 2: GRU gate pre-activations in TMEM: r [64,128) z [128,192) in [192,256) hn [256,320) ------------
  if (tid == 0) {
    run_slab(2, 0, 64, 0);    // Wih_r * y
    run_slab(3, 2, 64, 1);    // Whh_r * h
    load_slab(4);
    load_slab(5);
    run_slab(4, 0, 128, 0);   // Wih_z * y
    run_slab(5, 2, 128, 1);   // Whh_z * h
    load_slab(6);
    load_slab(7);
    run_slab(6, 0, 192, 0);   // Wih_n * y
    run_slab(7, 2, 256, 0);   // Whh_n * h
    umma_commit(bars + 4);
    load_slab(8);
  }
  mbar_wait(bars + 4, 1);
  tc_fence_after();
  {
    float gr[16], gz[16], gi[16], gh[16];
    tmem_ld16(lane_base + 64, gr);
    tmem_ld16(lane_base + 128, gz);
    tmem_ld16(lane_base + 192, gi);
    tmem_ld16(lane_base + 256, gh);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const int k = cg * 16 + c4 * 4;
      const int off = core_off64(row, k);
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
      *reinterpret_cast<float4*>(h_hi + off) = h;      // in place: this thread is the only reader of these 16 bytes
      *reinterpret_cast<float4*>(h_lo + off) = l;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- phase 3: acc[320,384) = h_new * fc_inter^T; meanwhile commit the new state to the slot arena -----------
  if (tid == 0) {
    run_slab(8, 2, 320, 0);
    umma_commit(bars + 4);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
    const int r = rg * 8 + rr;
    if (r < valid && s_commit[r]) {
      const int off = rg * 512 + (kb * 4 + cc) * 32 + rr * 4;
      const float4 a = *reinterpret_cast<const float4*>(h_hi + off), b = *reinterpret_cast<const float4*>(h_lo + off);
      *reinterpret_cast<float4*>(q.hstate + s_hoff[r] + kb * 16 + cc * 4) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  }
  mbar_wait(bars + 4, 0);
  tc_fence_after();
  {
    float v[16];
    tmem_ld16(lane_base + 320, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += sp[448 + cg * 16 + i];
    layernorm16(v, sp + 512, sp + 576);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int off = core_off64(row, cg * 16 + c * 4);
      const float4 a = *reinterpret_cast<const float4*>(y_hi + off);
      const float4 b = *reinterpret_cast<const float4*>(y_lo + off);
      *reinterpret_cast<float4*>(y_hi + off) =
          make_float4(v[c * 4] + a.x + b.x, v[c * 4 + 1] + a.y + b.y, v[c * 4 + 2] + a.z + b.z, v[c * 4 + 3] + a.w + b.w);
    }
  }
  tc_fence_before();
  __syncthreads();
  float* xo = q.xout + row0 * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {                             // coalesced write-out of the block output
    const int blk = warp + 16 * i, rg = blk >> 2, kb = blk & 3;
    const int r = rg * 8 + rr;
    if (r < valid)
      *reinterpret_cast<float4*>(xo + (size_t)r * C + kb * 16 + cc * 4) =
          *reinterpret_cast<const float4*>(y_hi + rg * 512 + (kb * 4 + cc) * 32 + rr * 4);
  }
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
  k_dprnn_post_tc<<<p.tiles0 + tiles1, TC_NT, POST_TC_SMEM, st>>>(p);
}

void init_dprnn_tc_kernels() {
  cudaFuncSetAttribute(k_dprnn_post_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POST_TC_SMEM);
}

}  // namespace dpdf
