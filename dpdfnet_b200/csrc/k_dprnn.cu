// Dual-path RNN block (a6 / a7): the two kernels that carry ~80 % of the FLOPs.
//
// k_dprnn_intra : bidirectional intra-frame GRU over the F' frequency positions of one hop
//                 (layers.py:126-132, 176-177).  Persistent per (stream tile, direction): the
//                 192x64 input and recurrent weight matrices of the direction live in REGISTERS for
//                 the whole sweep (each thread owns one hidden unit x one quarter of K for all three
//                 gates, 96 floats), the hidden state ping-pongs in shared memory, partial dot
//                 products are reduce-scattered over the 4 K-lanes with warp shuffles so that every
//                 lane finishes the gate math of a different stream, x_t tiles are prefetched with
//                 cp.async one step ahead.  All inner products run on packed FFMA2.
// k_dprnn_post  : everything position-parallel in the block, fused over a 128-row tile:
//                 fc_intra + LayerNorm + residual, inter-frame GRUCell against the per-stream state,
//                 fc_inter + LayerNorm + residual (layers.py:178-196).
#include "engine.h"

namespace dpdf {

struct IntraParams {
  const float* x[2];      // [B][Fp][64]    (index 0 = df branch, 1 = erb branch)
  float* hcat[2];         // [B][Fp][128]
  int Fp[2];
  const float* wih[2];    // [2][192][64]
  const float* whh[2];    // [2][192][64]
  const float* bias[2];   // [2][4][64]
  int tiles;              // ceil(B / BT)
  int B;
};

constexpr int ILD = 68;   // smem row stride (floats): 4s + 16u + jj bank pattern of the h exchange is conflict-free

// Position of element k of a 64-float row in shared memory.  The row is stored as [half][q][4]: the eight
// K-slices (q) of one 128-bit load are contiguous, so every broadcast LDS.128 touches one 128-byte line.
__device__ __forceinline__ int row_pos(int k) { return ((k >> 2) & 1) * 32 + (k >> 3) * 4 + (k & 3); }

template <int BT>
__global__ void __launch_bounds__(256, 1) k_dprnn_intra(IntraParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float xs[2][BT][ILD];
  __shared__ __align__(16) float hs[2][BT][ILD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // thread = (unit pair {up, up+32}, K-eighth q): 2 units x 6 matrices x 8 k = 96 stationary weights
  const int q = lane & 7, jj = lane >> 3, up = warp * 4 + jj;
  const int item = blockIdx.x;
  const int br = item / (2 * p.tiles);
  const int dir = (item % (2 * p.tiles)) / p.tiles;
  const int tile = item % p.tiles;
  const int T = p.Fp[br];
  const int b0 = tile * BT;
  const float* __restrict__ xg = p.x[br];
  float* __restrict__ hg = p.hcat[br];

  float2 wi[2][3][4], wh[2][3][4];
  {
    const float* Wih = p.wih[br] + (size_t)dir * 192 * C;
    const float* Whh = p.whh[br] + (size_t)dir * 192 * C;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const float4* si = reinterpret_cast<const float4*>(Wih + (size_t)(g * C + up + 32 * u) * C + q * 8);
        const float4* sh = reinterpret_cast<const float4*>(Whh + (size_t)(g * C + up + 32 * u) * C + q * 8);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 a = __ldg(si + c), b = __ldg(sh + c);
          wi[u][g][2 * c] = lo2(a); wi[u][g][2 * c + 1] = hi2(a);
          wh[u][g][2 * c] = lo2(b); wh[u][g][2 * c + 1] = hi2(b);
        }
      }
  }
  // after the reduce-scatter lane q finishes (stream q>>1 of the batch, unit up + 32*(q&1))
  const int myu = q & 1, myj = up + 32 * myu, mypos = row_pos(myj);
  const float* bias = p.bias[br] + dir * 4 * C;
  const float b_r = __ldg(bias + myj), b_z = __ldg(bias + C + myj), b_in = __ldg(bias + 2 * C + myj), b_hn = __ldg(bias + 3 * C + myj);

  for (int i = tid; i < BT * ILD; i += 256) (&hs[0][0][0])[i] = 0.f;     // h0 = 0 every frame

  auto prefetch = [&](int t, int buf) {
    const int f = dir ? T - 1 - t : t;
    for (int i = tid; i < BT * 16; i += 256) {
      const int s = i >> 4, c = i & 15;
      const int b = b0 + s;
      float* dst = &xs[buf][s][(c & 1) * 32 + (c >> 1) * 4];
      if (b < p.B) cp_async16(dst, xg + ((size_t)b * T + f) * C + c * 4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_h = [&](int buf, int t) {
    const int f = dir ? T - 1 - t : t;
    for (int i = tid; i < BT * 16; i += 256) {
      const int s = i >> 4, c = i & 15;
      const int b = b0 + s;
      if (b < p.B)
        *reinterpret_cast<float4*>(hg + ((size_t)b * T + f) * 2 * C + dir * C + c * 4) =
            *reinterpret_cast<const float4*>(&hs[buf][s][(c & 1) * 32 + (c >> 1) * 4]);
    }
  };

  prefetch(0, 0);
  cp_async_commit();
  int cur = 0;
  struct Frag { float4 x[2], h[2]; };
  for (int t = 0; t < T; ++t) {
    cp_async_wait<0>();
    __syncthreads();                       // x_t landed, h_t complete, previous buffers free
    if (t + 1 < T) prefetch(t + 1, (t + 1) & 1);
    cp_async_commit();
    if (t > 0) store_h(cur, t - 1);
    const float(*xb)[ILD] = xs[t & 1];
    const float(*hb)[ILD] = hs[cur];
    float(*hn)[ILD] = hs[cur ^ 1];
    auto load_frag = [&](Frag& f, int row) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        f.x[c] = *reinterpret_cast<const float4*>(&xb[row][c * 32 + q * 4]);
        f.h[c] = *reinterpret_cast<const float4*>(&hb[row][c * 32 + q * 4]);
      }
    };
    Frag fr;
    load_frag(fr, 0);
#pragma unroll 2
    for (int sb = 0; sb < BT / 4; ++sb) {
      float v[4][2][4];                                  // [stream][unit][gate r,z,in,hn]
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        Frag nx;                                         // software pipeline: next stream's operands
        load_frag(nx, min(sb * 4 + s + 1, BT - 1));      // are in flight while this one is multiplied
        // Issue order matters: the register file feeds one 64-bit operand per lane per cycle, so an FFMA2 only
        // sustains its 2-cycle rate when one of its three operand pairs comes from the operand-reuse cache.
        // Runs of six FFMA2 share the same x (or h) pair; the six accumulators of a run are independent.
        float2 ar[2], az[2], ain[2], ahn[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) ar[u] = az[u] = ain[u] = ahn[u] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 xv = fr.x[c], hv = fr.h[c];
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
            const float2 xo = hl ? hi2(xv) : lo2(xv);
            const float2 ho = hl ? hi2(hv) : lo2(hv);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              ar[u] = ffma2(wi[u][0][2 * c + hl], xo, ar[u]);
              az[u] = ffma2(wi[u][1][2 * c + hl], xo, az[u]);
              ain[u] = ffma2(wi[u][2][2 * c + hl], xo, ain[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              ar[u] = ffma2(wh[u][0][2 * c + hl], ho, ar[u]);
              az[u] = ffma2(wh[u][1][2 * c + hl], ho, az[u]);
              ahn[u] = ffma2(wh[u][2][2 * c + hl], ho, ahn[u]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          v[s][u][0] = ar[u].x + ar[u].y;
          v[s][u][1] = az[u].x + az[u].y;
          v[s][u][2] = ain[u].x + ain[u].y;
          v[s][u][3] = ahn[u].x + ahn[u].y;
        }
        fr = nx;
      }
      // reduce-scatter over the 8 K-lanes (xor 4, 2, 1): lane q ends with the four complete gate sums of
      // (stream sb*4 + (q>>1), unit up + 32*(q&1))
      float w2[2][2][4], w1[2][4], r[4];
      const bool b4 = (q & 4) != 0, b2 = (q & 2) != 0, b1 = (q & 1) != 0;
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float send = b4 ? v[s2][u][g] : v[s2 + 2][u][g];
            const float keep = b4 ? v[s2 + 2][u][g] : v[s2][u][g];
            w2[s2][u][g] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float send = b2 ? w2[0][u][g] : w2[1][u][g];
          const float keep = b2 ? w2[1][u][g] : w2[0][u][g];
          w1[u][g] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float send = b1 ? w1[0][g] : w1[1][g];
        const float keep = b1 ? w1[1][g] : w1[0][g];
        r[g] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      const int srow = sb * 4 + (q >> 1);
      const float hprev = hb[srow][mypos];
      const float rg = sigmoidf_(r[0] + b_r);
      const float zg = sigmoidf_(r[1] + b_z);
      const float ng = tanhf_(r[2] + b_in + rg * (r[3] + b_hn));
      hn[srow][mypos] = (1.0f - zg) * ng + zg * hprev;
    }
    cur ^= 1;
  }
  __syncthreads();
  store_h(cur, T - 1);
}

void launch_dprnn_intra(Engine& e, int blk, int B, cudaStream_t st) {
  IntraParams p{};
  p.x[0] = e.sc.c1;
  p.x[1] = blk == 0 ? e.sc.e3 : e.sc.xe;
  p.hcat[0] = e.sc.hcat_d;
  p.hcat[1] = e.sc.hcat_e;
  p.Fp[0] = NDF / 2;
  p.Fp[1] = e.d.fe[3];
  p.wih[0] = e.w.dprnn_df[blk].i_wih;  p.whh[0] = e.w.dprnn_df[blk].i_whh;  p.bias[0] = e.w.dprnn_df[blk].i_bias;
  p.wih[1] = e.w.dprnn_erb[blk].i_wih; p.whh[1] = e.w.dprnn_erb[blk].i_whh; p.bias[1] = e.w.dprnn_erb[blk].i_bias;
  p.B = B;
  int bt = e.intra_bt;
  if (bt == 0)    // one CTA per SM (register-resident weights): small tiles until the df CTAs fill the chip
    bt = B <= 4 * e.num_sms ? 8 : (B <= 16 * e.num_sms ? 16 : 32);
  p.tiles = (B + bt - 1) / bt;
  const int grid = 4 * p.tiles;
  if (bt == 8) launch_k(e, k_dprnn_intra<8>, dim3(grid), dim3(256), 0, st, p);
  else if (bt == 16) launch_k(e, k_dprnn_intra<16>, dim3(grid), dim3(256), 0, st, p);
  else launch_k(e, k_dprnn_intra<32>, dim3(grid), dim3(256), 0, st, p);
}

// ---------------------------------------------------------------------------------------------
struct PostBranch {
  const float* hcat;      // [rows][128]
  const float* xin;       // [rows][64]
  float* xout;            // [rows][64]
  float* hstate;          // inter-GRU state of this block: + slot*per_slot + f*64
  long long per_slot;
  int Fp;
  const float *fc_w, *fc_b, *ln_g, *ln_b, *wih, *whh, *bias, *fc2_w, *fc2_b, *ln2_g, *ln2_b;
};
struct PostParams {
  const IoDesc* io;
  PostBranch br[2];
  int tiles0;     // tiles of branch 0
  int B;
};

constexpr int P_LDH = 132, P_LD = 68;
constexpr int P_RA = 192 * P_LD * 2;            // 26112 floats: max(hcat tile + fc_w, W_ih + W_hh)
constexpr int P_RX = 128 * P_LD;
constexpr size_t POST_SMEM = (size_t)(P_RA + 2 * P_RX + 640) * sizeof(float) + 128 * sizeof(long long) + 128 * sizeof(int);
static_assert(128 * P_LDH + 64 * P_LDH <= P_RA, "phase-1 operands must fit the weight region");

__global__ void __launch_bounds__(256, 1) k_dprnn_post(PostParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* RA = smem;
  float* RX = RA + P_RA;
  float* RH = RX + P_RX;
  float* sp = RH + P_RX;                                  // small parameters, 640 floats
  long long* s_hoff = reinterpret_cast<long long*>(sp + 640);
  int* s_commit = reinterpret_cast<int*>(s_hoff + 128);

  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int bi = (int)blockIdx.x >= p.tiles0 ? 1 : 0;
  const PostBranch& q = p.br[bi];
  const long long row0 = (long long)(blockIdx.x - (bi ? p.tiles0 : 0)) * 128;
  const long long nrows = (long long)p.B * q.Fp;
  const int valid = (int)min((long long)128, nrows - row0);

  if (tid < 128) {
    long long off = -1;
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
  sp[192 + tid] = q.bias[tid];                             // [4][64]
  __syncthreads();                                         // s_hoff visible for the state loader

  // ---- phase 1: y = LN(fc_intra(hcat)) + x --------------------------------------------------
  {
    const float* hc = q.hcat + row0 * 2 * C;
    const float* xi = q.xin + row0 * C;
    tile_load_async<128, P_LDH, 256>(RA, 128, valid, [&](int r) { return hc + (size_t)r * 2 * C; });
    tile_load_async<128, P_LDH, 256>(RA + 128 * P_LDH, 64, 64, [&](int r) { return q.fc_w + (size_t)r * 2 * C; });
    tile_load_async<64, P_LD, 256>(RX, 128, valid, [&](int r) { return xi + (size_t)r * C; });
    tile_load_async<64, P_LD, 256>(RH, 128, valid, [&](int r) { return q.hstate + s_hoff[r]; });
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  float2 acc[4][8];
  acc_zero(acc);
  tile_mac<128, P_LDH, P_LDH, 4, 8>(RA, RA + 128 * P_LDH, acc, tx, ty);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) v[jj] = acc[i][jj].x + acc[i][jj].y + sp[tx + 8 * jj];
    row_layernorm8(v, sp + 64, sp + 128, tx);
    float* xr = RX + (ty + 32 * i) * P_LD;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) xr[tx + 8 * jj] += v[jj];
  }
  __syncthreads();                                         // RA free; RX = y

  // ---- phase 2: inter-frame GRUCell ----------------------------------------------------------
  tile_load_async<64, P_LD, 256>(RA, 192, 192, [&](int r) { return q.wih + (size_t)r * C; });
  tile_load_async<64, P_LD, 256>(RA + 192 * P_LD, 192, 192, [&](int r) { return q.whh + (size_t)r * C; });
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const float* Wih = RA;
  const float* Whh = RA + 192 * P_LD;
  float rg[4][8], zg[4][8];
  acc_zero(acc);
  tile_mac<64, P_LD, P_LD, 4, 8>(RX, Wih, acc, tx, ty);
  tile_mac<64, P_LD, P_LD, 4, 8>(RH, Whh, acc, tx, ty);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) rg[i][jj] = sigmoidf_(acc[i][jj].x + acc[i][jj].y + sp[192 + tx + 8 * jj]);
  acc_zero(acc);
  tile_mac<64, P_LD, P_LD, 4, 8>(RX, Wih + 64 * P_LD, acc, tx, ty);
  tile_mac<64, P_LD, P_LD, 4, 8>(RH, Whh + 64 * P_LD, acc, tx, ty);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) zg[i][jj] = sigmoidf_(acc[i][jj].x + acc[i][jj].y + sp[256 + tx + 8 * jj]);
  acc_zero(acc);
  tile_mac<64, P_LD, P_LD, 4, 8>(RH, Whh + 128 * P_LD, acc, tx, ty);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) rg[i][jj] *= acc[i][jj].x + acc[i][jj].y + sp[384 + tx + 8 * jj];
  acc_zero(acc);
  tile_mac<64, P_LD, P_LD, 4, 8>(RX, Wih + 128 * P_LD, acc, tx, ty);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* hr = RH + (ty + 32 * i) * P_LD;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float ng = tanhf_(acc[i][jj].x + acc[i][jj].y + sp[320 + tx + 8 * jj] + rg[i][jj]);
      rg[i][jj] = (1.0f - zg[i][jj]) * ng + zg[i][jj] * hr[tx + 8 * jj];
    }
  }
  __syncthreads();                                         // all reads of h_prev and W done
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* hr = RH + (ty + 32 * i) * P_LD;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) hr[tx + 8 * jj] = rg[i][jj];
  }
  tile_load_async<64, P_LD, 256>(RA, 64, 64, [&](int r) { return q.fc2_w + (size_t)r * C; });
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();                                         // RH = h_new, RA = fc_inter weights
  for (int it = tid; it < 128 * 16; it += 256) {           // commit the new inter state
    const int r = it >> 4, c = (it & 15) * 4;
    if (r < valid && s_commit[r])
      *reinterpret_cast<float4*>(q.hstate + s_hoff[r] + c) = *reinterpret_cast<const float4*>(RH + r * P_LD + c);
  }

  // ---- phase 3: out = LN(fc_inter(h_new)) + y -------------------------------------------------
  acc_zero(acc);
  tile_mac<64, P_LD, P_LD, 4, 8>(RH, RA, acc, tx, ty);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) v[jj] = acc[i][jj].x + acc[i][jj].y + sp[448 + tx + 8 * jj];
    row_layernorm8(v, sp + 512, sp + 576, tx);
    float* xr = RX + (ty + 32 * i) * P_LD;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) xr[tx + 8 * jj] += v[jj];
  }
  __syncthreads();
  float* xo = q.xout + row0 * C;
  for (int it = tid; it < 128 * 16; it += 256) {
    const int r = it >> 4, c = (it & 15) * 4;
    if (r < valid) *reinterpret_cast<float4*>(xo + (size_t)r * C + c) = *reinterpret_cast<const float4*>(RX + r * P_LD + c);
  }
}

void launch_dprnn_post(Engine& e, int blk, int B, cudaStream_t st) {
  PostParams p{};
  p.io = e.io_dev;
  p.B = B;
  auto fill = [&](PostBranch& b, const DprnnW& w, const float* hcat, const float* xin, float* xout, float* hstate, int Fp) {
    b.hcat = hcat; b.xin = xin; b.xout = xout; b.Fp = Fp;
    b.per_slot = (long long)e.d.N * Fp * C;
    b.hstate = hstate + (size_t)blk * Fp * C;
    b.fc_w = w.fc_w; b.fc_b = w.fc_b; b.ln_g = w.ln_g; b.ln_b = w.ln_b;
    b.wih = w.r_wih; b.whh = w.r_whh; b.bias = w.r_bias;
    b.fc2_w = w.fc2_w; b.fc2_b = w.fc2_b; b.ln2_g = w.ln2_g; b.ln2_b = w.ln2_b;
  };
  fill(p.br[0], e.w.dprnn_df[blk], e.sc.hcat_d, e.sc.c1, e.sc.c1, e.st.inter_df, NDF / 2);
  fill(p.br[1], e.w.dprnn_erb[blk], e.sc.hcat_e, blk == 0 ? e.sc.e3 : e.sc.xe, e.sc.xe, e.st.inter_erb, e.d.fe[3]);
  p.tiles0 = (int)(((long long)B * (NDF / 2) + 127) / 128);
  const int tiles1 = (int)(((long long)B * e.d.fe[3] + 127) / 128);
  launch_k(e, k_dprnn_post, dim3(p.tiles0 + tiles1), dim3(256), POST_SMEM, st, p);
}

void init_dprnn_kernels() {
  cudaFuncSetAttribute(k_dprnn_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POST_SMEM);
}

}  // namespace dpdf
