// k_gru_tc: GRUCell(256) (a8 - a10: the embedding GRU and the two decoder GRU stacks; layers.py:1117-1188,
// torch.nn.GRUCell semantics) with both gate GEMMs on the tensor cores.
//
// One CTA = (tile of 128 streams, chunk of 64 hidden units, cell of the launch): gate pre-activations
//     [128 x 256] = x[128 x 256] * W_ih[chunk]^T  (r | z | in)  +  h[128 x 256] * W_hh[chunk]^T  (r, z on top; hn apart)
// as a K-pipelined chain of tcgen05.mma.kind::f16 instructions (FP16 hi/lo split, FP32 accumulate in TMEM, see
// tc_common.cuh).  K = 512 is walked in eight 64-wide stages: the eight converter warps load the FP32 activation
// chunk, split it and store it as K-major operand images (2-deep ring), the issuer warp streams the matching
// 48 KB weight slab ([192 gate rows][64 k], hi | lo; weights.py: tc_w) through a 3-deep ring of bulk copies and
// fires 24 MMAs per stage; stage completion (tcgen05.commit) frees both rings.  Epilogue: thread = (TMEM lane =
// stream, 32 units): gates, h' = (1 - z) n + z h  ->  hout (the state itself is committed by k_gru_commit).
//
// Template parameter UC = hidden units per CTA.  UC = 64 (above) is the throughput form.  At latency batch sizes the
// launch is bound by the weight stream of each CTA - eight dependent 48 KB slabs through a 3-deep ring, 27 us per launch
// at 1024 streams for 5 us of MMAs - so UC = 32 halves the slab (the three 4 KB row blocks r | z | n of the unit half are
// pulled out of the same packed image), makes the ring six deep in the same shared memory and spreads the weight stream
// over twice as many SMs.
#include "engine.h"
#include "tc_common.cuh"

namespace dpdf {

namespace {

using namespace tc;

constexpr int GT_CONV = 256;                 // converter / epilogue threads
constexpr int GT_NT = GT_CONV + 32;          // + issuer warp
constexpr int GT_AIMG = 128 * 64 * 2;        // one FP16 [128][64] image
constexpr int GT_WSLAB = 2 * 192 * 64 * 2;   // packed slab of a 64-unit chunk: [192 gate rows][64 k], hi | lo (48 KB)
constexpr int GT_WRING = 3 * GT_WSLAB;       // bytes of the weight ring
constexpr int GT_OFF_W = 2 * 2 * GT_AIMG;    // A ring: 2 stages x (hi | lo)
constexpr int GT_OFF_MISC = GT_OFF_W + GT_WRING;
constexpr size_t GRU_TC_SMEM = GT_OFF_MISC + 1024 + 192;   // + slot offsets [128] int64 ... barriers

}  // namespace

struct GRUTcParams {
  const IoDesc* io;
  GRUProblem prob[2];
  int B;
};

template <int UC>
__global__ void __launch_bounds__(GT_NT, 1) k_gru_tc(GRUTcParams p) {
  constexpr int SLAB = GT_WSLAB * UC / 64;                    // this CTA's slab: [3 UC gate rows][64 k], hi | lo
  constexpr int NW = GT_WRING / SLAB;                         // ring depth: 3 (UC = 64) or 6 (UC = 32)
  pdl_trigger();        // griddepcontrol.wait comes after the prologue and the first weight slabs: neither depends on the previous kernel
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* Asm = smem_raw;
  unsigned char* Wsm = smem_raw + GT_OFF_W;
  long long* s_hoff = reinterpret_cast<long long*>(smem_raw + GT_OFF_MISC);
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem_raw + GT_OFF_MISC + 1024);   // [NW]
  uint64_t* done = full_w + NW;                                                    // [8], one per stage, single use
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 8);

  const GRUProblem& q = p.prob[blockIdx.z];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * 128;
  const int uc = blockIdx.y;                                  // unit chunk: hidden units [UC uc, UC uc + UC)
  const int valid = min(128, p.B - b0);

  if (tid < 128) s_hoff[tid] = tid < valid ? (long long)io_slot(p.io, b0 + tid) * q.hs_stride : 0;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NW + 8; ++i) mbar_init(full_w + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<4 * UC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---- issuer warp: weight slab ring + MMAs -----------------------------------------------------------------
    // packed per 64-unit chunk (weights.py: tc_w): 8 stages x [r 64 | z 64 | n 64 rows][64 k] hi | lo, 8 rows per KB
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(q.w.tc_w) + (size_t)(uc * UC / 64) * 8 * GT_WSLAB;
    // Step i of the walk is K stage (i + 4) & 7: the RECURRENT half (h_prev * W_hh, stages 4..7) comes first.  It depends on
    // nothing but the state of the previous hop, so as a programmatic dependent (dense tail, Engine::tail_pdl) the CTA
    // loads, converts and multiplies it while the kernel that produces x is still running; only the x half waits.
    auto load_w = [&](int i) {
      mbar_expect_tx(full_w + i % NW, SLAB);
      unsigned char* dst = Wsm + (i % NW) * SLAB;
      const unsigned char* src = wsrc + (size_t)((i + 4) & 7) * GT_WSLAB;
      if constexpr (UC == 64) {
        bulk_g2s(dst, src, GT_WSLAB, full_w + i % NW);
      } else {                                               // the unit half's row blocks of r, z and n, out of the hi and the lo image
        const int sub = (uc % (64 / UC)) * UC * 128;         // byte offset of UC rows inside a 64-row gate block
#pragma unroll
        for (int img = 0; img < 2; ++img)
#pragma unroll
          for (int gate = 0; gate < 3; ++gate)
            bulk_g2s(dst + img * (SLAB / 2) + gate * UC * 128, src + img * (GT_WSLAB / 2) + gate * 8192 + sub, UC * 128, full_w + i % NW);
      }
    };
    if (elect_one()) {                                       // elect.sync, not `lane == 0`: tc_common.cuh:elect_one
#pragma unroll
      for (int s = 0; s < NW; ++s) load_w(s);
    }
    // (no griddepcontrol.wait in this warp: it touches weights and shared memory only)
    for (int s = 0; s < 8; ++s) {                               // s = step of the walk (see load_w)
      const int buf = s & 1;
      if (buf == 0) asm volatile("bar.sync 1, %0;" ::"n"(GT_NT) : "memory");      // A images of stage s written
      else asm volatile("bar.sync 2, %0;" ::"n"(GT_NT) : "memory");
      if (elect_one()) {
        tc_fence_after();
        mbar_wait(full_w + s % NW, (s / NW) & 1);
        const uint32_t ah = smem_u32(Asm) + buf * 2 * GT_AIMG, al = ah + GT_AIMG;
        const uint32_t bh = smem_u32(Wsm) + (s % NW) * SLAB, bl = bh + SLAB / 2;
        const bool hpart = s < 4;
        const uint32_t ncol = hpart ? 3u * UC : 2u * UC;     // in (x part) / hn (h part)
        const uint32_t nfirst = (s & 3) == 0 ? 0u : 1u;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t dah = umma_desc(ah + ks * 256, 1024), dal = umma_desc(al + ks * 256, 1024);
          const uint64_t dbh = umma_desc(bh + ks * 256, 1024), dbl = umma_desc(bl + ks * 256, 1024);
          const uint64_t dnh = umma_desc(bh + 2 * UC * 128 + ks * 256, 1024), dnl = umma_desc(bl + 2 * UC * 128 + ks * 256, 1024);
          umma_f16(tmem, dah, dbh, idesc_f16(128, 2 * UC), (s > 0 || ks > 0) ? 1u : 0u);   // r | z: x and h parts add up
          umma_f16(tmem, dal, dbh, idesc_f16(128, 2 * UC), 1);
          umma_f16(tmem, dah, dbl, idesc_f16(128, 2 * UC), 1);
          umma_f16(tmem + ncol, dah, dnh, idesc_f16(128, UC), ks > 0 ? 1u : nfirst);
          umma_f16(tmem + ncol, dal, dnh, idesc_f16(128, UC), 1);
          umma_f16(tmem + ncol, dah, dnl, idesc_f16(128, UC), 1);
        }
        umma_commit(done + s);
        if (s >= 1 && s + NW - 1 < 8) {                       // slab ring slot of stage s - 1 is free once its MMAs are done
          mbar_wait(done + s - 1, 0);
          load_w(s + NW - 1);
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
    uint32_t ovf = 0;                                         // FP16 range guard (tc_common.cuh:f16_nonfinite)
    // the activation chunk of stage s + 1 is in flight while stage s is converted: with the loads issued only when their
    // stage came up, every one of the eight stages ate a full L2 round trip (27 us per launch at 1024 streams)
    auto load_stage = [&](int s, float4 (&v)[8]) {              // s = step of the walk: h_prev chunks first (previous hop's state),
      const int k0 = (s & 3) * 64 + g * 4;                      // then, once the grid dependency has resolved, the x chunks
      const bool hpart = s < 4;
      if (s == 4) pdl_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < valid) v[i] = __ldg(reinterpret_cast<const float4*>((hpart ? q.hstate + s_hoff[r] : q.x + (size_t)(b0 + r) * H) + k0));
      }
    };
    auto convert_stage = [&](int s, const float4 (&v)[8]) {
      const int buf = s & 1;
      if (s >= 2) mbar_wait(done + s - 2, 0);                 // MMAs that read this image pair are complete
      unsigned char* img = Asm + buf * 2 * GT_AIMG;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        uint2 h, l;
        split2_f16(v[i].x, v[i].y, h.x, l.x);
        split2_f16(v[i].z, v[i].w, h.y, l.y);
        ovf |= f16_nonfinite(h.x) | f16_nonfinite(h.y);
        unsigned char* dst = img + (r >> 3) * 1024 + (g >> 1) * 128 + (r & 7) * 16 + (g & 1) * 8;
        *reinterpret_cast<uint2*>(dst) = h;
        *reinterpret_cast<uint2*>(dst + GT_AIMG) = l;
      }
      fence_async_smem();
      if (buf == 0) asm volatile("bar.arrive 1, %0;" ::"n"(GT_NT) : "memory");
      else asm volatile("bar.arrive 2, %0;" ::"n"(GT_NT) : "memory");
    };
    float4 va[8], vb[8];
    load_stage(0, va);
#pragma unroll 1
    for (int s = 0; s < 8; s += 2) {
      load_stage(s + 1, vb);
      convert_stage(s, va);
      if (s + 2 < 8) load_stage(s + 2, va);
      convert_stage(s + 1, vb);
    }
    if (ovf) p.io->err[DPDF_ERRW_RANGE] = 1;
    // ---- epilogue: thread = (stream row = TMEM lane, UC / 2 units) ------------------------------------------------
    mbar_wait(done + 7, 0);
    tc_fence_after();
    const int qd = warp & 3, half = warp >> 2, row = qd * 32 + lane;
    const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16) + half * (UC / 2);
    const float* bias = q.w.bias;
#pragma unroll 1
    for (int c = 0; c < UC / 16; ++c) {
      uint32_t gr[8], gz[8], gi[8], gh[8];
      tmem_ld8_nowait(ta + c * 8, gr);
      tmem_ld8_nowait(ta + UC + c * 8, gz);
      tmem_ld8_nowait(ta + 2 * UC + c * 8, gi);
      tmem_ld8_nowait(ta + 3 * UC + c * 8, gh);
      const int u = uc * UC + half * (UC / 2) + c * 8;
      float hp[8];
      if (row < valid) {
        const float4 a = *reinterpret_cast<const float4*>(q.hstate + s_hoff[row] + u);
        const float4 b = *reinterpret_cast<const float4*>(q.hstate + s_hoff[row] + u + 4);
        hp[0] = a.x; hp[1] = a.y; hp[2] = a.z; hp[3] = a.w; hp[4] = b.x; hp[5] = b.y; hp[6] = b.z; hp[7] = b.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) hp[e] = 0.f;
      }
      tmem_ld_wait();
      float hn[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float rg = sigmoidf_(__uint_as_float(gr[e]) + __ldg(bias + u + e));
        const float zg = sigmoidf_(__uint_as_float(gz[e]) + __ldg(bias + H + u + e));
        const float ng = tanhf_(__uint_as_float(gi[e]) + __ldg(bias + 2 * H + u + e) + rg * (__uint_as_float(gh[e]) + __ldg(bias + 3 * H + u + e)));
        hn[e] = (1.0f - zg) * ng + zg * hp[e];
      }
      if (row < valid) {
        float* dst = q.hout + (size_t)(b0 + row) * H + u;
        *reinterpret_cast<float4*>(dst) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<4 * UC>(tmem);
}

void launch_gru_tc(Engine& e, const GRUProblem* probs, int nprob, int B, cudaStream_t st) {
  GRUTcParams p{};
  p.io = e.io_dev;
  p.B = B;
  for (int i = 0; i < nprob; ++i) p.prob[i] = probs[i];
  // 32-unit CTAs while twice the grid still fits one wave (Engine::gru_uc = 0: auto)
  const int tiles = (std::max(B, e.total_B) + 127) / 128;
  const bool small = e.gru_uc == 32 || (e.gru_uc == 0 && tiles * (H / 32) * nprob <= e.num_sms);
  if (small) launch_k(e, k_gru_tc<32>, dim3((B + 127) / 128, H / 32, nprob), dim3(GT_NT), GRU_TC_SMEM, st, p);
  else launch_k(e, k_gru_tc<64>, dim3((B + 127) / 128, H / 64, nprob), dim3(GT_NT), GRU_TC_SMEM, st, p);
}

void init_gru_tc_kernels() {
  cudaFuncSetAttribute(k_gru_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRU_TC_SMEM);
  cudaFuncSetAttribute(k_gru_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRU_TC_SMEM);
}

}  // namespace dpdf
