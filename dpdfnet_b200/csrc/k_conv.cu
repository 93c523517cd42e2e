// Convolutional stages (a5, conv parts of a9 and a10).
//   k_erb_conv0  : 3x3 conv 1->64 over the 3-frame feature ring + BN + ReLU   (onnx_model/dpdfnet.py:206-211)
//   k_sepconv    : [depthwise 1x3 (stride / sub-pixel) | grouped 3x3 from the df ring] prologue feeding a
//                  64x64 pointwise GEMM tile + BN + ReLU; also fuses the decoder "pathway" adds
//                  (layers.py:761-834, 895-973; onnx_model/dpdfnet.py:212-227, 361-363)
//   k_conv0_out  : pathway add + 64->1 1x3 conv + BN + sigmoid -> ERB / per-bin gains (:364)
//   k_df_pathway : 5-frame c0 ring -> grouped (2 x 32->5, 5x1) + pointwise 10x10 + BN + ReLU, added to the
//                  tanh'd df_out coefficients and pushed into the coefficient ring (:503-515)
#include "engine.h"

namespace dpdf {

// ---------------------------------------------------------------------------------------------
struct Conv0Params {
  IoDesc* io;
  State st;
  const float *w, *bias;   // [9][64], [64]
  float* e0;               // [B][fe0][64]
  int fe0, fe_feat, B;
};

// thread = (b, chunk of CH0_L positions, 4 channels): the 9 x 4 weights stay in registers and the three ring rows slide
// along f, so a position costs three (broadcast) tap loads, 36 FMAs and one 16-byte store; the 16 channel threads of a
// position write one whole 256-byte row.  (The first form - thread = (b, f, 4 channels), 25 loads per 16-byte store - was
// load-issue bound: 127 us for the 252 MB it writes at 2048 streams of the 48 kHz model, 3 x the HBM time.)
constexpr int CH0_L = 8;
__global__ void __launch_bounds__(256) k_erb_conv0(Conv0Params p) {
  pdl_trigger();
  pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) {      // hop counter tick (see IoDesc)
    p.io->t_out = p.io->t_in;
    p.io->t_in = p.io->t_in + 1;
  }
  const int nch = (p.fe0 + CH0_L - 1) / CH0_L;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;      // < 2^32: B * nch * 16
  const unsigned total = (unsigned)p.B * nch * (C / 4);
  if (idx >= total) return;
  const int c = (idx & 15) * 4;
  const unsigned bc = idx >> 4;
  const int ch = bc % nch, b = bc / nch;
  const int f0 = ch * CH0_L, f1 = min(p.fe0, f0 + CH0_L);
  const int slot = io_slot(p.io, b);
  const int pos = p.st.pos[slot];
  const float* ring = p.st.erb_ring + (size_t)slot * 3 * p.fe_feat;
  const float* rowp[3] = {ring + ((pos + 1) % 3) * p.fe_feat, ring + ((pos + 2) % 3) * p.fe_feat, ring + (pos % 3) * p.fe_feat};
  float4 w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = __ldg(reinterpret_cast<const float4*>(p.w + i * C + c));
  const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bias + c));
  float xm[3], x0[3], xp[3];                      // taps at f - 1, f, f + 1 of the three frames
#pragma unroll
  for (int kt = 0; kt < 3; ++kt) {
    xm[kt] = f0 > 0 ? rowp[kt][f0 - 1] : 0.f;
    x0[kt] = rowp[kt][f0];
  }
  float* out = p.e0 + ((size_t)b * p.fe0 + f0) * C + c;
  for (int f = f0; f < f1; ++f, out += C) {
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) xp[kt] = f + 1 < p.fe0 ? rowp[kt][f + 1] : 0.f;
    float4 acc = bias;
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const float x[3] = {xm[kt], x0[kt], xp[kt]};
#pragma unroll
      for (int kf = 0; kf < 3; ++kf) {
        const float4 ww = w[kt * 3 + kf];
        acc.x = fmaf(ww.x, x[kf], acc.x); acc.y = fmaf(ww.y, x[kf], acc.y); acc.z = fmaf(ww.z, x[kf], acc.z); acc.w = fmaf(ww.w, x[kf], acc.w);
      }
      xm[kt] = x0[kt];
      x0[kt] = xp[kt];
    }
    *reinterpret_cast<float4*>(out) = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
  }
}

void launch_erb_conv0(Engine& e, int B, cudaStream_t st) {
  Conv0Params p{e.io_dev, e.st, e.w.erb_conv0_w, e.w.erb_conv0_b, e.sc.e0, e.d.fe[0], e.d.fe_feat, B};
  const long long total = (long long)B * ((e.d.fe[0] + CH0_L - 1) / CH0_L) * (C / 4);
  launch_k(e, k_erb_conv0, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, p);
}

// ---------------------------------------------------------------------------------------------
constexpr int SEP_MAXP = 4;
constexpr int SEP_LD = 68;
struct SepParams {
  const IoDesc* io;
  State st;
  SepProblem prob[SEP_MAXP];
  int nprob, B;
};
constexpr int SEP_TM = 2;                 // 64-row tiles: 35 KB smem, <= 128 registers -> 2+ CTAs per SM hide the prologue latency
constexpr int SEP_ROWS = 32 * SEP_TM;
constexpr size_t SEP_SMEM = (size_t)(SEP_ROWS * SEP_LD + 64 * SEP_LD + 64) * sizeof(float);

__global__ void __launch_bounds__(256, 2) k_sepconv(SepParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                    // [SEP_ROWS][68]
  float* Ws = As + SEP_ROWS * SEP_LD;  // [64][68]
  float* bs = Ws + 64 * SEP_LD;        // [64]
  int pi = 0;
#pragma unroll
  for (int i = 1; i < SEP_MAXP; ++i)
    if (i < p.nprob && (int)blockIdx.x >= p.prob[i].tile0) pi = i;
  const SepProblem& q = p.prob[pi];
  const int tid = threadIdx.x;
  const long long row0 = (long long)(blockIdx.x - q.tile0) * SEP_ROWS;
  const long long nrows = (long long)p.B * q.Fout;

  // pointwise weights + bias -> smem (async)
  for (int i = tid; i < 64 * 16; i += 256) cp_async16(Ws + (i >> 4) * SEP_LD + (i & 15) * 4, q.pw + (i >> 4) * 64 + (i & 15) * 4);
  if (tid < 16) cp_async16(bs + tid * 4, q.bias + tid * 4);
  cp_async_commit();

  // prologue: A[row][c].  A thread keeps the same 4 channels for all of its rows, so the depthwise / grouped
  // 3x3 taps and the pathway affine are loaded once; all rows' activation loads are issued back to back.
  {
    const int c = (tid & 15) * 4;
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
    for (int i = 0; i < SEP_ROWS / 16; ++i) {
      const int r = (tid >> 4) + 16 * i;
      const long long row = row0 + r;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < nrows) {
        const int b = (int)(row / q.Fout), fo = (int)(row % q.Fout);
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
      *reinterpret_cast<float4*>(As + r * SEP_LD + c) = acc;
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  const int tx = tid & 7, ty = tid >> 3;
  float2 acc[SEP_TM][8];
  acc_zero(acc);
  tile_mac<64, SEP_LD, SEP_LD, SEP_TM, 8>(As, Ws, acc, tx, ty);
  __syncthreads();                          // everyone done reading As
#pragma unroll
  for (int i = 0; i < SEP_TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = tx + 8 * j;
      As[(ty + 32 * i) * SEP_LD + col] = fmaxf(acc[i][j].x + acc[i][j].y + bs[col], 0.f);
    }
  __syncthreads();
  for (int it = tid; it < SEP_ROWS * 16; it += 256) {
    const int r = it >> 4, c = (it & 15) * 4;
    const long long row = row0 + r;
    if (row >= nrows) continue;
    float4 v = *reinterpret_cast<const float4*>(As + r * SEP_LD + c);
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
}

void launch_sepconv(Engine& e, const SepProblem* probs, int nprob, int B, cudaStream_t st) {
  SepParams p{};
  p.io = e.io_dev;
  p.st = e.st;
  p.nprob = nprob;
  p.B = B;
  int tiles = 0;
  for (int i = 0; i < nprob; ++i) {
    p.prob[i] = probs[i];
    p.prob[i].tile0 = tiles;
    tiles += (int)(((long long)B * probs[i].Fout + SEP_ROWS - 1) / SEP_ROWS);
  }
  launch_k(e, k_sepconv, dim3(tiles), dim3(256), SEP_SMEM, st, p);
}

// ---------------------------------------------------------------------------------------------
struct Conv0OutParams {
  const float *e0, *d1, *pa, *pb, *w, *bias;   // w [3][64]
  float* m;
  int fe0, B;
};

// warp = (b, chunk of `chunk` positions), lane = 2 channels.  u[f] = relu(e0[f] * a + pb) + d1[f] of a row is formed once
// and slides through the three taps, four positions per round (eight 8-byte loads in flight per lane), and the four
// lane-partial sums of a round are reduced together with 6 shuffles instead of 20.  (The first form - one warp per
// position, three rows of e0 and d1 re-read per output - took 187 us for the 504 MB it reads at 2048 streams of the
// 48 kHz model, 2.4 x the HBM time.)
__global__ void __launch_bounds__(256) k_conv0_out(Conv0OutParams p, int chunk, int nch) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.B * nch) return;
  const int b = warp / nch, f0 = (warp % nch) * chunk, f1 = min(p.fe0, f0 + chunk);
  const float2 a = __ldg(reinterpret_cast<const float2*>(p.pa) + lane);
  const float2 pb = __ldg(reinterpret_cast<const float2*>(p.pb) + lane);
  const float2 w0 = __ldg(reinterpret_cast<const float2*>(p.w) + lane), w1 = __ldg(reinterpret_cast<const float2*>(p.w + C) + lane),
               w2 = __ldg(reinterpret_cast<const float2*>(p.w + 2 * C) + lane);
  const float bias = __ldg(p.bias);
  const float2* e0 = reinterpret_cast<const float2*>(p.e0 + (size_t)b * p.fe0 * C) + lane;
  const float2* d1 = reinterpret_cast<const float2*>(p.d1 + (size_t)b * p.fe0 * C) + lane;
  auto urow = [&](int fi) -> float2 {                       // zero outside [0, fe0): the conv's zero padding
    if (fi < 0 || fi >= p.fe0) return make_float2(0.f, 0.f);
    const float2 e = __ldg(e0 + (size_t)fi * (C / 2)), d = __ldg(d1 + (size_t)fi * (C / 2));
    return make_float2(fmaxf(fmaf(e.x, a.x, pb.x), 0.f) + d.x, fmaxf(fmaf(e.y, a.y, pb.y), 0.f) + d.y);
  };
  auto dot = [&](const float2& um, const float2& u0, const float2& up) {
    float acc = w0.x * um.x;
    acc = fmaf(w0.y, um.y, acc);
    acc = fmaf(w1.x, u0.x, acc); acc = fmaf(w1.y, u0.y, acc);
    acc = fmaf(w2.x, up.x, acc); acc = fmaf(w2.y, up.y, acc);
    return acc;
  };
  float2 um = urow(f0 - 1), u0 = urow(f0);
  for (int f = f0; f < f1; f += 4) {
    float2 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = urow(f + 1 + j);     // rows past f1 are only read for the halo of position f1 - 1
    float s0 = dot(um, u0, u[0]), s1 = dot(u0, u[0], u[1]), s2 = dot(u[0], u[1], u[2]), s3 = dot(u[1], u[2], u[3]);
    um = u[2];
    u0 = u[3];
    // four sums over the 32 lanes at once: halve the values twice, then reduce the remaining one over 8 lanes
    const bool hi16 = lane & 16, hi8 = lane & 8;
    const float k0 = hi16 ? s2 : s0, k1 = hi16 ? s3 : s1, g0 = hi16 ? s0 : s2, g1 = hi16 ? s1 : s3;
    const float t0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 16), t1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 16);   // lanes < 16: positions 0, 1; others 2, 3
    float r = (hi8 ? t1 : t0) + __shfl_xor_sync(0xffffffffu, hi8 ? t0 : t1, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    const int j = (hi16 ? 2 : 0) + (hi8 ? 1 : 0);           // the position this lane's sum belongs to
    if ((lane & 7) == 0 && f + j < f1) p.m[(size_t)b * p.fe0 + f + j] = sigmoidf_(r + bias);
  }
}

void launch_conv0_out(Engine& e, int B, cudaStream_t st) {
  Conv0OutParams p{e.sc.e0, e.sc.d1, e.w.convp_a[3], e.w.convp_b[3], e.w.conv0_out_w, e.w.conv0_out_b, e.sc.m, e.d.fe[0], B};
  // positions per warp, a multiple of 4: up to 32 (6 % halo rows) once the grid is large, 8 for small grids, where one
  // warp walking all 32 bands of a stream is a 23 us serial chain at 1024 streams (13 us with 8)
  const int cap = (long long)B * e.d.fe[0] >= (1 << 18) ? 32 : 8;
  const int nch = (e.d.fe[0] + cap - 1) / cap, chunk = ((e.d.fe[0] + nch - 1) / nch + 3) & ~3;
  const int nch2 = (e.d.fe[0] + chunk - 1) / chunk;
  const long long warps = (long long)B * nch2;
  launch_k(e, k_conv0_out, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, st, p, chunk, nch2);
}

// ---------------------------------------------------------------------------------------------
struct DfPathParams {
  const IoDesc* io;
  State st;
  const float *w, *pw, *bias;   // [10][5][32], [10][10], [10]
  const float* co;              // [B][96][10] tanh(df_out)
  int B;
  float* tmp;                   // [B][96][10] pathway term alone (k_df_pathway<.., true> -> k_df_combine)
};

// One warp = four consecutive bins of a stream; lane = (bin, 4-channel chunk).  Every load of the 5-frame c0 ring
// is a fully used 128-byte line per bin, weights are broadcast float4 from shared memory, the 8 chunk-lanes of a
// bin are reduced with xor shuffles.  HBM-bound on the ring (120 KB per stream-frame).
// EARLY = true: only the pathway term is computed and left in p.tmp - it depends on the c0 ring alone, so the launch sits
// on a forked stream right behind df_conv0 and runs beside the DPRNN stack (which leaves most SMs idle at latency batch
// sizes) instead of on the critical decoder tail; k_df_combine then adds tanh(df_out) and pushes the coefficient ring.
template <bool RING16, bool EARLY = false>       // RING16: c0 ring stored in half precision (option c0_fp16); templates so that the FP32 path costs nothing
__global__ void __launch_bounds__(256) k_df_pathway(DfPathParams p) {
  pdl_trigger();
  __shared__ __align__(16) float ws[2 * ORD * 8 * 5 * 4];        // [g][kt][ci/4][o][4]
  __shared__ float pws[100], bs[10];
  const int tid = threadIdx.x, lane = tid & 31;
  // The 6.8 KB of weights are staged ONCE per CTA (before the grid dependency resolves) and the CTA then walks its share
  // of the (stream, 4-bin) items: with one item per warp every one of the 3072 CTAs of a 1024-stream hop paid the staging
  // and its barrier before its first ring load (47 us for 126 MB, where this kernel is the critical path of the forked
  // decoder tail).
  for (int i = tid; i < 2 * ORD * 8 * 5 * 4; i += 256) {
    const int e = i & 3, o = (i >> 2) % 5, c4 = (i / 20) % 8, kt = (i / 160) % ORD, g = i / 800;
    ws[i] = __ldg(p.w + ((g * 5 + o) * ORD + kt) * 32 + c4 * 4 + e);
  }
  if (tid < 100) pws[tid] = __ldg(p.pw + tid);
  if (tid < 10) bs[tid] = __ldg(p.bias + tid);
  pdl_wait();
  __syncthreads();
  const long long nitems = (long long)p.B * (NDF / 4);
  const int c8 = lane & 7;
  for (long long witem = (long long)blockIdx.x * 8 + (tid >> 5); witem < nitems; witem += (long long)gridDim.x * 8) {   // (b, group of 4 bins)
    const int b = (int)(witem / (NDF / 4)), f = (int)(witem % (NDF / 4)) * 4 + (lane >> 3);
    const int slot = io_slot(p.io, b);
    const int pos = p.st.pos[slot];
    const float* ring = p.st.c0_ring + (size_t)slot * ORD * NDF * C;
    float4 x[ORD][2];
    if (RING16) {                                      // half the bytes of the 120 KB ring read per stream
      const __half* ringh = reinterpret_cast<const __half*>(ring);
#pragma unroll
      for (int kt = 0; kt < ORD; ++kt) {
        const __half* src = ringh + ((size_t)((pos + 1 + kt) % ORD) * NDF + f) * C + c8 * 4;
        x[kt][0] = unpack4_f16(*reinterpret_cast<const uint2*>(src));
        x[kt][1] = unpack4_f16(*reinterpret_cast<const uint2*>(src + 32));
      }
    } else {
#pragma unroll
      for (int kt = 0; kt < ORD; ++kt) {
        const float* src = ring + ((size_t)((pos + 1 + kt) % ORD) * NDF + f) * C + c8 * 4;
        x[kt][0] = *reinterpret_cast<const float4*>(src);
        x[kt][1] = *reinterpret_cast<const float4*>(src + 32);
      }
    }
    float t[10];
#pragma unroll
    for (int o = 0; o < 10; ++o) t[o] = 0.f;
#pragma unroll
    for (int kt = 0; kt < ORD; ++kt)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4* wp = reinterpret_cast<const float4*>(ws) + (g * ORD + kt) * 40 + c8 * 5;
        const float4 xv = x[kt][g];
#pragma unroll
        for (int o = 0; o < 5; ++o) {
          const float4 w = wp[o];
          t[g * 5 + o] = fmaf(w.x, xv.x, fmaf(w.y, xv.y, fmaf(w.z, xv.z, fmaf(w.w, xv.w, t[g * 5 + o]))));
        }
      }
#pragma unroll
    for (int o = 0; o < 10; ++o) {
      t[o] += __shfl_xor_sync(0xffffffffu, t[o], 1);
      t[o] += __shfl_xor_sync(0xffffffffu, t[o], 2);
      t[o] += __shfl_xor_sync(0xffffffffu, t[o], 4);
    }
    const bool warm = (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) != 0;
    float* dst = p.st.coef_ring + (((size_t)slot * 3 + pos % 3) * NDF + f) * 10;
    const float* cop = p.co + ((size_t)b * NDF + f) * 10;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int oo = c8 + 8 * rep;                    // lanes 0..7 of a bin finish outputs 0..7, lanes 0,1 also 8,9
      if (oo < 10) {
        float u = bs[oo];
#pragma unroll
        for (int i = 0; i < 10; ++i) u = fmaf(pws[oo * 10 + i], t[i], u);
        if constexpr (EARLY) p.tmp[((size_t)b * NDF + f) * 10 + oo] = fmaxf(u, 0.f);
        else dst[oo] = warm ? 0.f : cop[oo] + fmaxf(u, 0.f);
      }
    }
  }
}

// coefficient ring slot of this hop <- tanh(df_out) + pathway term (k_df_pathway<.., true>); thread = four coefficients
__global__ void __launch_bounds__(256) k_df_combine(DfPathParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int PER = NDF * 2 * ORD / 4;                     // float4 per stream
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)p.B * PER) return;
  const int b = idx / PER, r = idx % PER;
  const int slot = io_slot(p.io, b);
  const int pos = p.st.pos[slot];
  const bool warm = (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) != 0;
  const float4 a = *reinterpret_cast<const float4*>(p.co + (size_t)idx * 4), t = *reinterpret_cast<const float4*>(p.tmp + (size_t)idx * 4);
  float4* dst = reinterpret_cast<float4*>(p.st.coef_ring + ((size_t)slot * 3 + pos % 3) * NDF * 10) + r;
  *dst = warm ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
}

void launch_df_pathway_early(Engine& e, int B, cudaStream_t st) {
  DfPathParams p{e.io_dev, e.st, e.w.dfp_w, e.w.dfp_pw, e.w.dfp_b, e.sc.co, B, e.sc.dfp};
  const long long ctas = ((long long)B * (NDF / 4) + 7) / 8;
  const unsigned grid = (unsigned)std::min<long long>(ctas, 4LL * e.num_sms);
  if (e.st.c0_fp16) launch_k(e, k_df_pathway<true, true>, dim3(grid), dim3(256), 0, st, p);
  else launch_k(e, k_df_pathway<false, true>, dim3(grid), dim3(256), 0, st, p);
}

void launch_df_combine(Engine& e, int B, cudaStream_t st) {
  DfPathParams p{e.io_dev, e.st, e.w.dfp_w, e.w.dfp_pw, e.w.dfp_b, e.sc.co, B, e.sc.dfp};
  const long long total = (long long)B * (NDF * 2 * ORD / 4);
  launch_k(e, k_df_combine, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, p);
}

void launch_df_pathway(Engine& e, int B, cudaStream_t st) {
  DfPathParams p{e.io_dev, e.st, e.w.dfp_w, e.w.dfp_pw, e.w.dfp_b, e.sc.co, B, nullptr};
  const long long ctas = ((long long)B * (NDF / 4) + 7) / 8;             // one item per warp ...
  const unsigned grid = (unsigned)std::min<long long>(ctas, 4LL * e.num_sms);   // ... or the four resident CTAs per SM (62 registers) walking the items
  if (e.st.c0_fp16) launch_k(e, k_df_pathway<true>, dim3(grid), dim3(256), 0, st, p);
  else launch_k(e, k_df_pathway<false>, dim3(grid), dim3(256), 0, st, p);
}

// ---------------------------------------------------------------------------------------------
// The same 5-frame grouped conv as k_df_pathway, as pending partial sums: when the
// frame of hop t arrives it is multiplied by all five taps at once and added to the outputs of hops t .. t+4
// (acc slot (pos + d) % 5 <- + W[4 - d] * c0_t); slot pos % 5 is then complete, consumed and cleared.  Per stream
// and hop this reads the new frame (24 KB, still in L2 from k_sepconv_tc) and 19 KB of accumulators instead of the
// 120 KB ring.  The raw ring is still written (by the df_conv0 kernel) because the reference-layout state export needs it.
struct DfPathPsParams {
  const IoDesc* io;
  State st;
  const float *w, *pw, *bias;   // [10][5][32], [10][10], [10]
  const float* co;              // [B][96][10] tanh(df_out)
  const float* c0;              // [B][96][64] this hop's frame
  int B;
};

__global__ void __launch_bounds__(256) k_df_pathway_ps(DfPathPsParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float ws[2 * ORD * 8 * 5 * 4];        // [g][kt][ci/4][o][4]
  __shared__ float pws[100], bs[10], tsm[8][4][10];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 2 * ORD * 8 * 5 * 4; i += 256) {
    const int e = i & 3, o = (i >> 2) % 5, c4 = (i / 20) % 8, kt = (i / 160) % ORD, g = i / 800;
    ws[i] = __ldg(p.w + ((g * 5 + o) * ORD + kt) * 32 + c4 * 4 + e);
  }
  if (tid < 100) pws[tid] = __ldg(p.pw + tid);
  if (tid < 10) bs[tid] = __ldg(p.bias + tid);
  __syncthreads();
  const long long witem = ((long long)blockIdx.x * 256 + tid) >> 5;       // (b, group of 4 bins)
  if (witem >= (long long)p.B * (NDF / 4)) return;
  const int b = (int)(witem / (NDF / 4)), jb = lane >> 3, f = (int)(witem % (NDF / 4)) * 4 + jb, c8 = lane & 7;
  const int slot = io_slot(p.io, b);
  const int pos = p.st.pos[slot];
  const bool warm = (io_flags(p.io, b) & DPDF_FLAG_WARMUP_) != 0;
  float4 x[2];
  {
    const float* src = p.c0 + ((size_t)b * NDF + f) * C + c8 * 4;
    x[0] = warm ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(src);
    x[1] = warm ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(src + 32);
  }
  // t[d * 10 + g * 5 + o]: this lane's 4-channel share of W[tap 4 - d] * frame for the output d hops ahead
  float t[56];
#pragma unroll
  for (int d = 0; d < ORD; ++d)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const float4* wp = reinterpret_cast<const float4*>(ws) + (g * ORD + (ORD - 1 - d)) * 40 + c8 * 5;
#pragma unroll
      for (int o = 0; o < 5; ++o) {
        const float4 w = wp[o];
        t[d * 10 + g * 5 + o] = fmaf(w.x, x[g].x, fmaf(w.y, x[g].y, fmaf(w.z, x[g].z, w.w * x[g].w)));
      }
    }
#pragma unroll
  for (int i = 50; i < 56; ++i) t[i] = 0.f;
  // reduce-scatter over the 8 channel lanes of a bin (xor 4, 2, 1): lane c8 ends with the seven complete sums
  // base .. base + 6, base = 28 b2 + 14 b1 + 7 b0
  const bool b2 = (c8 & 4) != 0, b1 = (c8 & 2) != 0, b0 = (c8 & 1) != 0;
  float r1[28], r2[14], r3[7];
#pragma unroll
  for (int i = 0; i < 28; ++i) {
    const float keep = b2 ? t[28 + i] : t[i], send = b2 ? t[i] : t[28 + i];
    r1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 14; ++i) {
    const float keep = b1 ? r1[14 + i] : r1[i], send = b1 ? r1[i] : r1[14 + i];
    r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float keep = b0 ? r2[7 + i] : r2[i], send = b0 ? r2[i] : r2[7 + i];
    r3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  const int base = (b2 ? 28 : 0) + (b1 ? 14 : 0) + (b0 ? 7 : 0);
  float* acc = p.st.dfp_acc + (size_t)slot * ORD * NDF * 10;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int idx = base + i;
    if (idx < 50) {
      const int d = idx / 10, oo = idx - d * 10;
      float* a = acc + ((size_t)f * ORD + (pos + d) % ORD) * 10 + oo;      // [96][5][10]: the 50 sums of a bin are contiguous
      const float v = *a + r3[i];
      if (d == 0) { tsm[warp][jb][oo] = v; *a = 0.f; }        // complete: consumed below, slot restarts for hop t + 5
      else *a = v;
    }
  }
  __syncwarp();
  float* dst = p.st.coef_ring + (((size_t)slot * 3 + pos % 3) * NDF + f) * 10;
  const float* cop = p.co + ((size_t)b * NDF + f) * 10;
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const int oo = c8 + 8 * rep;                      // lanes 0..7 of a bin finish outputs 0..7, lanes 0,1 also 8,9
    if (oo < 10) {
      float u = bs[oo];
#pragma unroll
      for (int i = 0; i < 10; ++i) u = fmaf(pws[oo * 10 + i], tsm[warp][jb][i], u);
      dst[oo] = warm ? 0.f : cop[oo] + fmaxf(u, 0.f);
    }
  }
}

void launch_df_pathway_ps(Engine& e, int B, cudaStream_t st) {
  DfPathPsParams p{e.io_dev, e.st, e.w.dfp_w, e.w.dfp_pw, e.w.dfp_b, e.sc.co, e.sc.c0, B};
  const long long threads = (long long)B * (NDF / 4) * 32;
  launch_k(e, k_df_pathway_ps, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, p);
}

void init_conv_kernels() {
  cudaFuncSetAttribute(k_sepconv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEP_SMEM);
}

}  // namespace dpdf
