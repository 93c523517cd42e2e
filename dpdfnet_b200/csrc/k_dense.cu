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
  for (int i = tid; i < 64 * kch; i += 256) {
    const int r = i / kch, c = (i % kch) * 4;
    if (r < valid) cp_async16(As + r * LD + c, src + (size_t)(b0 + r) * ld + c);
    else *reinterpret_cast<float4*>(As + r * LD + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* wg = q.w.w + (size_t)g * Ng * Kg;
  for (int i = tid; i < Ng * kch; i += 256) {
    const int r = i / kch, c = (i % kch) * 4;
    cp_async16(Ws + r * LD + c, wg + (size_t)r * Kg + c);
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
  if (maxnj <= 4) k_gl<4><<<grid, 256, smem, st>>>(p);
  else if (maxnj <= 8) k_gl<8><<<grid, 256, smem, st>>>(p);
  else if (maxnj <= 16) k_gl<16><<<grid, 256, smem, st>>>(p);
  else k_gl<20><<<grid, 256, smem, st>>>(p);
}

// ---------------------------------------------------------------------------------------------
constexpr int GRU_MAXP = 2;
constexpr int G_LDA = 260, G_LDW = 68;
struct GRUParams {
  const IoDesc* io;
  GRUProblem prob[GRU_MAXP];
  int B;
};
constexpr size_t GRU_SMEM = (size_t)(2 * 64 * G_LDA + 2 * 32 * G_LDW) * sizeof(float) + 64 * sizeof(long long);

__global__ void __launch_bounds__(256, 1) k_gru(GRUParams p) {
  extern __shared__ __align__(16) float smem[];
  float* Ax = smem;                     // [64][260]
  float* Ah = Ax + 64 * G_LDA;          // [64][260]
  float* Wb = Ah + 64 * G_LDA;          // [2][32][68]
  long long* s_hoff = reinterpret_cast<long long*>(Wb + 2 * 32 * G_LDW);
  const GRUProblem& q = p.prob[blockIdx.z];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int b0 = blockIdx.x * 64;
  const int u0 = blockIdx.y * 32;
  const int valid = min(64, p.B - b0);

  if (tid < 64) {
    s_hoff[tid] = tid < valid ? (long long)io_slot(p.io, b0 + tid) * q.hs_stride : -1;
  }
  __syncthreads();
  tile_load_async<256, G_LDA, 256>(Ax, 64, valid, [&](int r) { return q.x + (size_t)(b0 + r) * H; });
  tile_load_async<256, G_LDA, 256>(Ah, 64, valid, [&](int r) { return q.hstate + s_hoff[r]; });
  cp_async_commit();

  // chunk c: gates r, z -> (ih k0..3, hh k0..3); gate n -> (hh k0..3) then (ih k0..3)
  auto chunk_src = [&](int c, const float*& wsrc, int& use_h, int& kc) {
    int g, part;
    if (c < 16) { g = c >> 3; part = (c >> 2) & 1; }
    else { g = 2; part = c < 20 ? 1 : 0; }
    kc = c & 3;
    use_h = part;
    const float* base = part ? q.w.whh : q.w.wih;
    wsrc = base + (size_t)(g * H + u0) * H + kc * 64;
  };
  auto prefetch = [&](int c) {
    const float* wsrc; int use_h, kc;
    chunk_src(c, wsrc, use_h, kc);
    float* dst = Wb + (c & 1) * 32 * G_LDW;
    for (int i = tid; i < 32 * 16; i += 256) cp_async16(dst + (i >> 4) * G_LDW + (i & 15) * 4, wsrc + (size_t)(i >> 4) * H + (i & 15) * 4);
  };
  prefetch(0);
  cp_async_commit();

  float2 acc[2][4];
  float rg[2][4], zg[2][4];
  acc_zero(acc);
  for (int c = 0; c < 24; ++c) {
    cp_async_wait<0>();
    __syncthreads();
    if (c + 1 < 24) prefetch(c + 1);
    cp_async_commit();
    const float* wsrc; int use_h, kc;
    chunk_src(c, wsrc, use_h, kc);
    tile_mac<64, G_LDA, G_LDW, 2, 4>((use_h ? Ah : Ax) + kc * 64, Wb + (c & 1) * 32 * G_LDW, acc, tx, ty);
    if (c == 7 || c == 15 || c == 19 || c == 23) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int u = u0 + tx + 8 * j;
          const float s = acc[i][j].x + acc[i][j].y;
          if (c == 7) rg[i][j] = sigmoidf_(s + __ldg(q.w.bias + u));
          else if (c == 15) zg[i][j] = sigmoidf_(s + __ldg(q.w.bias + H + u));
          else if (c == 19) rg[i][j] *= s + __ldg(q.w.bias + 3 * H + u);
          else {
            const int row = ty + 32 * i;
            const float ng = tanhf_(s + __ldg(q.w.bias + 2 * H + u) + rg[i][j]);
            const float hn = (1.0f - zg[i][j]) * ng + zg[i][j] * Ah[row * G_LDA + u];
            if (row < valid) q.hout[(size_t)(b0 + row) * H + u] = hn;
          }
        }
      acc_zero(acc);
    }
  }
  // Every unit-chunk CTA of a cell reads the full h_prev rows, so the state may only be overwritten
  // once all of them are done: hout is the hand-off buffer and k_gru_commit (next launch) copies it.
}

__global__ void k_gru_commit(const IoDesc* io, const float* hout, float* hstate, int hs_stride, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // float4 index
  if (idx >= B * (H / 4)) return;
  const int b = idx / (H / 4), c = (idx % (H / 4)) * 4;
  if (io_flags(io, b) & DPDF_FLAG_WARMUP_) return;
  *reinterpret_cast<float4*>(hstate + (size_t)io_slot(io, b) * hs_stride + c) =
      *reinterpret_cast<const float4*>(hout + (size_t)b * H + c);
}

void launch_gru(Engine& e, const GRUProblem* probs, int nprob, int B, cudaStream_t st) {
  GRUParams p{};
  p.io = e.io_dev;
  p.B = B;
  for (int i = 0; i < nprob; ++i) p.prob[i] = probs[i];
  dim3 grid((B + 63) / 64, H / 32, nprob);
  k_gru<<<grid, 256, GRU_SMEM, st>>>(p);
  for (int i = 0; i < nprob; ++i)
    k_gru_commit<<<(B * (H / 4) + 255) / 256, 256, 0, st>>>(e.io_dev, probs[i].hout, probs[i].hstate, probs[i].hs_stride, B);
}

void init_dense_kernels() {
  cudaFuncSetAttribute(k_gru, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRU_SMEM);
}

}  // namespace dpdf
