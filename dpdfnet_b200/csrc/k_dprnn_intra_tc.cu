// k_dprnn_intra_tc: the bidirectional intra-frame GRU of a DPRNN block (layers.py:126-132, 176-177) on the
// 5th-generation tensor cores.
//
// One CTA = (branch, direction, tile of 128 streams) and walks the F' frequency positions sequentially.  Per step
//     P[128 x 256] = x_t * W_ih^T  (N = 192: r | z | in)   +   h_{t-1} * W_hh^T  (r, z on top; hn in its own 64 columns)
// is a chain of tcgen05.mma.kind::f16 instructions with the FP32 accumulator in tensor memory.  FP32 accuracy is
// kept by the error-compensated split x = hi + lo (both FP16, 11-bit significands; three products hi*hi + lo*hi
// + hi*lo; weights.py:fp16_split): the same accuracy as 3xTF32 at half the shared-memory footprint and twice the
// tensor rate, which is what lets both 192x64 weight matrices of a direction stay resident in shared memory
// (96 KB) next to the operand images of x_t, x_{t+1} and h.
//
// Pipeline per step t (TMEM is double buffered, 2 x 256 columns):
//   thread 0     : h-part MMAs of step t (critical path) -> commit -> x-part MMAs of step t+1 into the other buffer
//   all threads  : wait commit, tcgen05.ld the four gate pre-activations of (row = TMEM lane, 16 units), gate
//                  math in registers (h_{t-1} never leaves registers), write h_t as FP16 hi/lo operand rows for
//                  the next step and as FP32 to hcat[b][f][dir*64 + u]; convert the prefetched x_{t+2} tile.
// The x-part MMAs and the global loads of x are hidden behind the gate math of the previous step.
#include "engine.h"
#include "tc_common.cuh"

namespace dpdf {

namespace {

using namespace tc;

constexpr int ITC_NT = 512;                 // 16 warps: warp w -> TMEM lane quadrant w & 3, 16-unit group w >> 2
constexpr int W_IMG = 192 * 64 * 2;         // bytes of one FP16 [192][64] weight image
constexpr int A_IMG = 128 * 64 * 2;         // bytes of one FP16 [128][64] activation image
constexpr int OFF_X = 4 * W_IMG;            // x images: [2 buffers][hi | lo]
constexpr int OFF_H = OFF_X + 4 * A_IMG;    // h images: [hi | lo]
constexpr int OFF_BIAS = OFF_H + 2 * A_IMG; // [4][64] floats
constexpr int OFF_BAR = OFF_BIAS + 1024;    // two mbarriers + TMEM base slot
constexpr size_t INTRA_TC_SMEM = OFF_BAR + 64;

}  // namespace

struct IntraTcParams {
  const float* x[2];      // [B][Fp][64]    (index 0 = df branch, 1 = erb branch)
  float* hcat[2];         // [B][Fp][128]
  int Fp[2];
  const float* wimg[2];   // [2 dirs][W_ih hi | W_ih lo | W_hh hi | W_hh lo] FP16 operand images
  const float* bias[2];   // [2][4][64]
  int tiles;              // ceil(B / 128)
  int B;
};

__global__ void __launch_bounds__(ITC_NT, 1) k_dprnn_intra_tc(IntraTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* Wsm = smem_raw;
  unsigned char* Xsm = smem_raw + OFF_X;
  unsigned char* Hsm = smem_raw + OFF_H;
  float* sb = reinterpret_cast<float*>(smem_raw + OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + OFF_BAR);       // [0] weights landed, [1] step accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3, cg = warp >> 2, row = qd * 32 + lane;
  const int item = blockIdx.x;
  const int br = item / (2 * p.tiles);
  const int dir = (item % (2 * p.tiles)) / p.tiles;
  const int tile = item % p.tiles;
  const int T = br ? p.Fp[1] : p.Fp[0];
  const int b0 = tile * 128;
  const float* __restrict__ xg = br ? p.x[1] : p.x[0];
  float* __restrict__ hg = br ? p.hcat[1] : p.hcat[0];

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid < 256) sb[tid] = __ldg((br ? p.bias[1] : p.bias[0]) + dir * 4 * C + tid);
  for (int i = tid; i < 2 * A_IMG / 16; i += ITC_NT) reinterpret_cast<uint4*>(Hsm)[i] = make_uint4(0u, 0u, 0u, 0u);   // h_0 = 0
  __syncthreads();                                           // barriers initialised
  if (tid == 0) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(br ? p.wimg[1] : p.wimg[0]) + (size_t)dir * 4 * W_IMG;
    mbar_expect_tx(bars, 4 * W_IMG);
#pragma unroll
    for (int i = 0; i < 4; ++i) bulk_g2s(Wsm + i * W_IMG, src + (size_t)i * W_IMG, W_IMG, bars);
  }

  // ---- x tile staging: 128 rows x 8 chunks of 8 floats, two (row, chunk) items per thread ------------------
  // warp item j = warp + 16 i covers row group j >> 1, K half j & 1: lanes = 8 rows x 4 chunks, so the 16-byte
  // operand rows a warp writes are four contiguous 128-byte core matrices (conflict-free)
  const int xr_[2] = {((warp) >> 1) * 8 + (lane & 7), ((warp + 16) >> 1) * 8 + (lane & 7)};
  const int xkc = (warp & 1) * 4 + (lane >> 3);
  auto load_x = [&](int t, float (&v)[2][8]) {
    const int f = dir ? T - 1 - t : t;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int bb = b0 + xr_[i];
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
      if (bb < p.B) {
        const float4* src = reinterpret_cast<const float4*>(xg + ((size_t)bb * T + f) * C + xkc * 8);
        a = __ldg(src);
        c = __ldg(src + 1);
      }
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
      v[i][4] = c.x; v[i][5] = c.y; v[i][6] = c.z; v[i][7] = c.w;
    }
  };
  auto store_x = [&](int buf, const float (&v)[2][8]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 hi, lo;
      split8_f16(v[i], hi, lo);
      unsigned char* dst = Xsm + buf * 2 * A_IMG + img16_off(xr_[i], xkc);
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + A_IMG) = lo;
    }
  };
  float xv[2][8];
  load_x(0, xv);
  store_x(0, xv);
  if (T > 1) {
    load_x(1, xv);
    store_x(1, xv);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- MMA issue (thread 0) ---------------------------------------------------------------------------------
  const uint32_t w_base = smem_u32(Wsm), x_base = smem_u32(Xsm), h_base = smem_u32(Hsm);
  auto mma3 = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                         // K = 64 in steps of 16 halves = two core matrices = 256 B
      const uint64_t dah = umma_desc(a_hi + ks * 256, 1024), dal = umma_desc(a_lo + ks * 256, 1024);
      const uint64_t dbh = umma_desc(b_hi + ks * 256, 1024), dbl = umma_desc(b_lo + ks * 256, 1024);
      umma_f16(d, dah, dbh, idesc, accumulate);
      umma_f16(d, dal, dbh, idesc, 1);
      umma_f16(d, dah, dbl, idesc, 1);
      accumulate = 1;
    }
  };
  auto x_mma = [&](int t) {                                  // P[t & 1][0, 192) = x_t * W_ih^T
    const uint32_t xa = x_base + (t & 1) * 2 * A_IMG;
    mma3(tmem + (t & 1) * 256, xa, xa + A_IMG, w_base, w_base + W_IMG, idesc_f16(128, 192), 0);
  };
  auto h_mma = [&](int t) {
    const uint32_t d = tmem + (t & 1) * 256;
    const uint32_t whi = w_base + 2 * W_IMG, wlo = w_base + 3 * W_IMG;
    mma3(d, h_base, h_base + A_IMG, whi, wlo, idesc_f16(128, 128), 1);                               // r, z += h * W_hh[r,z]^T
    mma3(d + 192, h_base, h_base + A_IMG, whi + 16 * 1024, wlo + 16 * 1024, idesc_f16(128, 64), 0);  // hn = h * W_hh[n]^T
  };
  if (tid == 0) {
    mbar_wait(bars, 0);                                      // weight images landed (async proxy -> async proxy)
    x_mma(0);
    h_mma(0);
    umma_commit(bars + 1);
    if (T > 1) x_mma(1);
  }

  // ---- the sweep --------------------------------------------------------------------------------------------
  float h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = 0.f;
  const int bme = b0 + row;
  const bool live = bme < p.B;
  const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16) + cg * 16;
  for (int t = 0; t < T; ++t) {
    if (t + 2 < T) load_x(t + 2, xv);                        // in flight during the wait
    mbar_wait(bars + 1, t & 1);
    tc_fence_after();
    const int f = dir ? T - 1 - t : t;
    float* hdst = hg + ((size_t)bme * T + f) * 2 * C + dir * C + cg * 16;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t gr[8], gz[8], gi[8], gh[8];
      const uint32_t ta = lane_addr + (t & 1) * 256 + half * 8;
      tmem_ld8_nowait(ta, gr);
      tmem_ld8_nowait(ta + 64, gz);
      tmem_ld8_nowait(ta + 128, gi);
      tmem_ld8_nowait(ta + 192, gh);
      tmem_ld_wait();
      float bu[4][8];                                        // biases of these 8 units (warp-uniform: broadcast LDS.128)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4* b4 = reinterpret_cast<const float4*>(sb + g * C + cg * 16 + half * 8);
        const float4 u0 = b4[0], u1 = b4[1];
        bu[g][0] = u0.x; bu[g][1] = u0.y; bu[g][2] = u0.z; bu[g][3] = u0.w;
        bu[g][4] = u1.x; bu[g][5] = u1.y; bu[g][6] = u1.z; bu[g][7] = u1.w;
      }
      float hn[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float r = sigmoidf_(__uint_as_float(gr[e]) + bu[0][e]);
        const float z = sigmoidf_(__uint_as_float(gz[e]) + bu[1][e]);
        const float n = tanhf_(__uint_as_float(gi[e]) + bu[2][e] + r * (__uint_as_float(gh[e]) + bu[3][e]));
        hn[e] = (1.0f - z) * n + z * h[half * 8 + e];
        h[half * 8 + e] = hn[e];
      }
      uint4 hi, lo;
      split8_f16(hn, hi, lo);
      unsigned char* dst = Hsm + img16_off(row, cg * 2 + half);
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + A_IMG) = lo;
      if (live) {
        *reinterpret_cast<float4*>(hdst + half * 8) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(hdst + half * 8 + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
      }
    }
    if (t + 2 < T) store_x(t & 1, xv);                       // x_mma(t) (reader of this buffer) completed with the commit
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      if (t + 1 < T) {
        h_mma(t + 1);
        umma_commit(bars + 1);
      }
      if (t + 2 < T) x_mma(t + 2);
    }
  }
  tc_fence_after();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

void launch_dprnn_intra_tc(Engine& e, int blk, int B, cudaStream_t st) {
  IntraTcParams p{};
  p.x[0] = e.sc.c1;
  p.x[1] = blk == 0 ? e.sc.e3 : e.sc.xe;
  p.hcat[0] = e.sc.hcat_d;
  p.hcat[1] = e.sc.hcat_e;
  p.Fp[0] = NDF / 2;
  p.Fp[1] = e.d.fe[3];
  p.wimg[0] = e.w.dprnn_df[blk].tc_intra;  p.bias[0] = e.w.dprnn_df[blk].i_bias;
  p.wimg[1] = e.w.dprnn_erb[blk].tc_intra; p.bias[1] = e.w.dprnn_erb[blk].i_bias;
  p.B = B;
  p.tiles = (B + 127) / 128;
  k_dprnn_intra_tc<<<4 * p.tiles, ITC_NT, INTRA_TC_SMEM, st>>>(p);
}

void init_dprnn_intra_tc_kernels() {
  cudaFuncSetAttribute(k_dprnn_intra_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTRA_TC_SMEM);
}

}  // namespace dpdf
