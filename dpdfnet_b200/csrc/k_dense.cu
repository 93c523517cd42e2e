// Dense per-stream layers of a8-a10: grouped linears and the five 256-wide GRU cells.
//   k_gl  : GroupedLinear (layers.py:1020-1046) - one CTA per (64-stream tile, group); fused bias,
//           optional addend and ReLU / tanh; input may be the concatenation of two buffers.
//   k_gru : torch.nn.GRUCell(256,256) (layers.py:1206-1259) for a 64-stream x 32-unit tile; x and
//           h_prev tiles stay resident in shared memory while the six [32 x 256] weight slabs stream
//           through a cp.async double buffer in 64-wide K chunks; gates fused in registers.
#include "engine.h"

namespace dpdf {

constexpr int GL_MAXP = 4;
struct GLParams {
  GLProblem prob[GL_MAXP];
  int g0[GL_MAXP + 1];       // first blockIdx.y of each problem
  int nprob, B;
};

template <int NJ>
__global__ void __launch_bounds__(256) k_gl(GLParams p) {
  pdl_trigger();        // griddepcontrol.wait sits between the weight loads and the activation loads
  extern __shared__ __align__(16) float smem[];
  int pi = 0;
#pragma unroll
  for (int i = 1; i < GL_MAXP; ++i)
    if (i < p.nprob && (int)blockIdx.y >= p.g0[i]) pi = i;
  const GLProblem& q = p.prob[pi];
  const int g = blockIdx.y - p.g0[pi];
  const int Kg = q.w.Kg, Ng = q.w.Ng;
  const int LD = Kg + 4;
  float* As = smem;                 // [64][LD]
  float* Ws = smem + 64 * LD;       // [Ng][LD]
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * 64;
  const int valid = min(64, p.B - b0);
  const int kch = Kg / 4;
  const int incol = g * Kg;
  const float* src;
  int ld;
  if (q.in1 == nullptr || incol < q.split) { src = q.in0 + incol; ld = q.ld0; }
  else { src = q.in1 + (incol - q.split); ld = q.ld1; }
  const float* wg = q.w.w + (size_t)g * Ng * Kg;
  for (int i = tid; i < Ng * kch; i += 256) {
    const int r = i / kch, c = (i % kch) * 4;
    cp_async16(Ws + r * LD + c, wg + (size_t)r * Kg + c);
  }
  pdl_wait();                                                 // the activations are the previous kernel's output
  for (int i = tid; i < 64 * kch; i += 256) {
    const int r = i / kch, c = (i % kch) * 4;
    if (r < valid) cp_async16(As + r * LD + c, src + (size_t)(b0 + r) * ld + c);
    else *reinterpret_cast<float4*>(As + r * LD + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int row = tid >> 2, ol = tid & 3;
  float2 acc[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = make_float2(0.f, 0.f);
  const float* ap = As + row * LD;
  for (int k = 0; k < Kg; k += 4) {
    const float4 a = *reinterpret_cast<const float4*>(ap + k);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int o = ol + 4 * j;
      if (o < Ng) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + o * LD + k);
        acc[j] = ffma2(lo2(a), lo2(w), acc[j]);
        acc[j] = ffma2(hi2(a), hi2(w), acc[j]);
      }
    }
  }
  if (row < valid) {
    const int b = b0 + row;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int o = ol + 4 * j;
      if (o < Ng) {
        const int col = g * Ng + o;
        float y = acc[j].x + acc[j].y + __ldg(q.w.b + col);
        if (q.addend) y += q.addend[(size_t)b * q.lda + col];
        if (q.act == 1) y = fmaxf(y, 0.f);
        else if (q.act == 2) y = tanhf(y);
        q.out[(size_t)b * q.ldo + q.col0 + col] = y;
      }
    }
  }
}

void launch_gl(Engine& e, const GLProblem* probs, int nprob, int B, cudaStream_t st) {
  GLParams p{};
  p.nprob = nprob;
  p.B = B;
  int groups = 0, maxnj = 0;
  size_t smem = 0;
  for (int i = 0; i < nprob; ++i) {
    p.prob[i] = probs[i];
    p.g0[i] = groups;
    groups += probs[i].w.G;
    maxnj = max(maxnj, (probs[i].w.Ng + 3) / 4);
    smem = max(smem, (size_t)(64 + probs[i].w.Ng) * (probs[i].w.Kg + 4) * sizeof(float));
  }
  p.g0[nprob] = groups;
  dim3 grid((B + 63) / 64, groups);
  if (maxnj <= 4) launch_k(e, k_gl<4>, dim3(grid), dim3(256), smem, st, p);
  else if (maxnj <= 8) launch_k(e, k_gl<8>, dim3(grid), dim3(256), smem, st, p);
  else if (maxnj <= 16) launch_k(e, k_gl<16>, dim3(grid), dim3(256), smem, st, p);
  else launch_k(e, k_gl<20>, dim3(grid), dim3(256), smem, st, p);
}

// ---------------------------------------------------------------------------------------------
constexpr int GRU_MAXP = 2;
constexpr int G_LD = 68;
struct GRUParams {
  const IoDesc* io;
  GRUProblem prob[GRU_MAXP];
  int B;
};
template <int TM>
constexpr size_t gru_smem() { return (size_t)(2 * (32 * TM + 96) * G_LD) * sizeof(float) + 32 * TM * sizeof(long long); }

// One CTA = (32*TM streams) x (32 hidden units, all three gates).  K = 256 (x) + 256 (h_prev) is walked in
// eight 64-wide chunks; chunk kc+1 (activation tile + the [96 x 64] weight slab) streams in through
// cp.async while chunk kc is multiplied.  Thread tile: TM rows x 4 units x {r, z, n} -> 16 smem loads
// feed 96 packed FFMA2 per 4 k (TM = 4).
template <int TM>
__global__ void __launch_bounds__(256, 1) k_gru(GRUParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int ROWS = 32 * TM;
  extern __shared__ __align__(16) float smem[];
  float* Abuf = smem;                          // [2][ROWS][68]
  float* Wbuf = Abuf + 2 * ROWS * G_LD;        // [2][96][68]
  long long* s_hoff = reinterpret_cast<long long*>(Wbuf + 2 * 96 * G_LD);
  const GRUProblem& q = p.prob[blockIdx.z];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int b0 = blockIdx.x * ROWS;
  const int u0 = blockIdx.y * 32;
  const int valid = min(ROWS, p.B - b0);

  if (tid < ROWS) s_hoff[tid] = tid < valid ? (long long)io_slot(p.io, b0 + tid) * q.hs_stride : 0;
  __syncthreads();

  auto load_chunk = [&](int kc, int buf) {
    const bool hpart = kc >= 4;
    const int k0 = (kc & 3) * 64;
    float* Ad = Abuf + buf * ROWS * G_LD;
    float* Wd = Wbuf + buf * 96 * G_LD;
    for (int i = tid; i < ROWS * 16; i += 256) {
      const int r = i >> 4, c = (i & 15) * 4;
      float* dst = Ad + r * G_LD + c;
      if (r < valid) cp_async16(dst, (hpart ? q.hstate + s_hoff[r] : q.x + (size_t)(b0 + r) * H) + k0 + c);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* wb = hpart ? q.w.whh : q.w.wih;
    for (int i = tid; i < 96 * 16; i += 256) {
      const int r = i >> 4, c = (i & 15) * 4;
      cp_async16(Wd + r * G_LD + c, wb + (size_t)((r >> 5) * H + u0 + (r & 31)) * H + k0 + c);
    }
  };

  float2 ar[TM][4], az[TM][4], ain[TM][4], ahn[TM][4];
  acc_zero(ar); acc_zero(az); acc_zero(ain); acc_zero(ahn);

  auto mac = [&](const float* As, const float* Ws, float2 (&a0)[TM][4], float2 (&a1)[TM][4], float2 (&a2)[TM][4]) {
    const float* ap = As + ty * G_LD;
    const float* wp = Ws + tx * G_LD;
#pragma unroll 2
    for (int k = 0; k < 64; k += 4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(ap + i * 32 * G_LD + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(wp + (j * 8) * G_LD + k);
        const float4 w1 = *reinterpret_cast<const float4*>(wp + (32 + j * 8) * G_LD + k);
        const float4 w2 = *reinterpret_cast<const float4*>(wp + (64 + j * 8) * G_LD + k);
#pragma unroll
        for (int i = 0; i < TM; ++i) {          // runs share a[i] (operand-reuse cache), accumulators independent
          a0[i][j] = ffma2(lo2(a[i]), lo2(w0), a0[i][j]);
          a1[i][j] = ffma2(lo2(a[i]), lo2(w1), a1[i][j]);
          a2[i][j] = ffma2(lo2(a[i]), lo2(w2), a2[i][j]);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          a0[i][j] = ffma2(hi2(a[i]), hi2(w0), a0[i][j]);
          a1[i][j] = ffma2(hi2(a[i]), hi2(w1), a1[i][j]);
          a2[i][j] = ffma2(hi2(a[i]), hi2(w2), a2[i][j]);
        }
      }
    }
  };

  load_chunk(0, 0);
  cp_async_commit();
  for (int kc = 0; kc < 8; ++kc) {
    cp_async_wait<0>();
    __syncthreads();
    if (kc + 1 < 8) load_chunk(kc + 1, (kc + 1) & 1);
    cp_async_commit();
    const float* As = Abuf + (kc & 1) * ROWS * G_LD;
    const float* Ws = Wbuf + (kc & 1) * 96 * G_LD;
    if (kc < 4) mac(As, Ws, ar, az, ain);
    else mac(As, Ws, ar, az, ahn);
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = ty + 32 * i;
    if (row >= valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int u = u0 + tx + 8 * j;
      const float rg = sigmoidf_(ar[i][j].x + ar[i][j].y + __ldg(q.w.bias + u));
      const float zg = sigmoidf_(az[i][j].x + az[i][j].y + __ldg(q.w.bias + H + u));
      const float ng = tanhf_(ain[i][j].x + ain[i][j].y + __ldg(q.w.bias + 2 * H + u) +
                              rg * (ahn[i][j].x + ahn[i][j].y + __ldg(q.w.bias + 3 * H + u)));
      const float hprev = q.hstate[s_hoff[row] + u];
      q.hout[(size_t)(b0 + row) * H + u] = (1.0f - zg) * ng + zg * hprev;
    }
  }
  // Every unit-chunk CTA of a cell reads the full h_prev rows, so the state may only be overwritten
  // once all of them are done: hout is the hand-off buffer and k_gru_commit (next launch) copies it.
}

struct GRUCommitParams {
  const IoDesc* io;
  const float* hout[5];
  float* hstate[5];
  int stride[5];
  int n, B;
};
// Every unit-chunk CTA of a cell reads the full h_prev rows, and nothing else reads the GRU states inside a
// hop, so all five cells' new states are written back by one launch at the end of the hop.
__global__ void k_gru_commit(GRUCommitParams p) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (stream, 16-float chunk) of one cell: four 16-byte copies in flight per thread
  if (idx >= p.B * (H / 16)) return;
  const int b = idx / (H / 16), t = idx % (H / 16);            // 16 threads per row: float4 t, t + 16, t + 32, t + 48 (coalesced)
  if (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) return;
  const int slot = io_slot(p.io, b);
  const int cell = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(p.hout[cell] + (size_t)b * H) + t;
  float4* dst = reinterpret_cast<float4*>(p.hstate[cell] + (size_t)slot * p.stride[cell]) + t;
  const float4 v0 = src[0], v1 = src[16], v2 = src[32], v3 = src[48];
  dst[0] = v0; dst[16] = v1; dst[32] = v2; dst[48] = v3;
}

void launch_gru(Engine& e, const GRUProblem* probs, int nprob, int B, cudaStream_t st) {
  GRUParams p{};
  p.io = e.io_dev;
  p.B = B;
  for (int i = 0; i < nprob; ++i) p.prob[i] = probs[i];
  if (B >= 4 * e.num_sms) {          // enough streams to fill the chip with 128-row tiles
    dim3 grid((B + 127) / 128, H / 32, nprob);
    launch_k(e, k_gru<4>, dim3(grid), dim3(256), gru_smem<4>(), st, p);
  } else {
    dim3 grid((B + 63) / 64, H / 32, nprob);
    launch_k(e, k_gru<2>, dim3(grid), dim3(256), gru_smem<2>(), st, p);
  }
}

void launch_gru_commit(Engine& e, const GRUProblem* probs, int nprob, int B, cudaStream_t st) {
  GRUCommitParams p{};
  p.io = e.io_dev;
  p.n = nprob;
  p.B = B;
  for (int i = 0; i < nprob; ++i) { p.hout[i] = probs[i].hout; p.hstate[i] = probs[i].hstate; p.stride[i] = probs[i].hs_stride; }
  dim3 grid((B * (H / 16) + 255) / 256, nprob);
  launch_k(e, k_gru_commit, dim3(grid), dim3(256), 0, st, p);
}

void init_dense_kernels() {
  cudaFuncSetAttribute(k_gru<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gru_smem<4>());
  cudaFuncSetAttribute(k_gru<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gru_smem<2>());
}

}  // namespace dpdf
