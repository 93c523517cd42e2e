// Shared device helpers for the DPDFNet-B200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "dpdfnet_b200 kernels are written for sm_100a (packed FFMA2, 227 KB smem); no fallback path"
#endif

namespace dpdf {

constexpr int C = 64;        // conv channels / DPRNN hidden
constexpr int H = 256;       // embedding GRU width
constexpr int NDF = 96;      // deep-filter bins
constexpr int ORD = 5;       // deep-filter taps

#define DPDF_FLAG_WARMUP_ 1
#define DPDF_FLAG_ZERO_FEAT_ 2
#define DPDF_FLAG_ZERO_SPEC_ 8

// Per-call I/O descriptor, resident in device memory so that a captured CUDA graph of one hop
// can be replayed for every hop of a run without touching kernel parameters.
struct IoDesc {
  const float* in;        // pcm [B][in_stride] or spec [B][F][2]
  float* out;
  long long in_stride;    // floats between rows
  long long out_stride;
  const int* slot_ids;    // nullptr = identity
  const int* flags;       // nullptr = 0
  int t_in;               // hop index read by the analysis kernel
  int t_out;              // hop index read by the synthesis kernel
  int mode;               // 0 = pcm, 1 = spec
  int slot_base;          // identity slot mapping: slot = slot_base + b (lanes of one batched step, api.cu:run_step)
  int* err;               // device-visible error words in mapped host memory (api.cu:check_device_errors):
                          //   [0] overlapped post kernel gave up waiting for the sweep (tile skipped, hop invalid)
                          //   [1] an activation left the FP16 range / was non-finite in a tensor-core operand converter
};
#define DPDF_ERRW_OVERLAP 0
#define DPDF_ERRW_RANGE 1

__device__ __forceinline__ int io_slot(const IoDesc* io, int b) {
  return io->slot_ids ? __ldg(io->slot_ids + b) : io->slot_base + b;
}
// ---- c0 ring in half precision (option "c0_fp16") ------------------------------------------
// 58 % of a stream's state is the five-frame c0 ring that only the df pathway conv reads (30 720 floats, 120 KB per stream
// and hop).  Stored as FP16 it costs half the HBM bytes on both sides; the conv then sees its inputs rounded to 11
// significant bits (2^-12 relative), which moves the deep-filter coefficients by ~1e-5 and the waveform by < 1e-5
// (tests/test_gpu_hardening.py::test_c0_ring_fp16).  Frames are compact [96][64] halves at the start of the slot's region.
__device__ __forceinline__ uint2 pack4_f16(const float4& v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
__device__ __forceinline__ float4 unpack4_f16(const uint2& u) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ int io_flags(const IoDesc* io, int b) {
  return io->flags ? __ldg(io->flags + b) : 0;
}

// ---- programmatic dependent launch (engine.h:launch_k) ------------------------------------------------------
// Every kernel of a hop opens with pdl_trigger(); pdl_wait();  -- the trigger lets the next kernel of the chain become
// resident (launch latency, prologue) while this one runs, the wait blocks until the kernel before this one has
// completed and flushed its writes.  Both are no-ops for a kernel launched without the attribute.
#ifdef DPDF_NO_PDL      // microbenchmark switch (tools/ubench): what the two instructions cost a plainly launched kernel
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---- packed FP32 math (Blackwell FFMA2: two FMAs per lane per issue) ------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// ---- activations ------------------------------------------------------------------------------
// ex2.approx based (max rel. error 2^-22): ~1e-7 absolute on the gates, far inside the 1e-4 budget
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
  // 1 - 2/(1+e^{2x}); saturates correctly for |x| large (e^{2x} -> inf or 0)
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}

// ---- cp.async (LDGSTS) ----------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Copy a [rows][COLS] row-major global tile (row pointers given by functor) into smem with row
// stride LD using 16-byte cp.async; rows >= valid_rows are zero-filled.  COLS % 4 == 0.
template <int COLS, int LD, int NT, typename RowPtr>
__device__ __forceinline__ void tile_load_async(float* smem, int rows, int valid_rows, RowPtr row_ptr) {
  constexpr int CH = COLS / 4;
  for (int i = threadIdx.x; i < rows * CH; i += NT) {
    int r = i / CH, c = (i % CH) * 4;
    float* dst = smem + r * LD + c;
    if (r < valid_rows) cp_async16(dst, row_ptr(r) + c);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- register-tiled smem GEMM -----------------------------------------------------------------
// acc[TM][TN] (k-even / k-odd partial sums packed in a float2) += A[rows][K] * W[cols][K]^T
// 256 threads as 8 (tx, columns) x 32 (ty, rows); thread rows = ty + 32*i, cols = tx + 8*j.
// A and W are k-contiguous in smem with strides LDA / LDW floats; (LD/4) odd makes every
// 128-bit load conflict-free (4 distinct rows / 8 distinct cols per warp, rest broadcast).
template <int K, int LDA, int LDW, int TM, int TN>
__device__ __forceinline__ void tile_mac(const float* __restrict__ As, const float* __restrict__ Ws,
                                         float2 (&acc)[TM][TN], int tx, int ty) {
  static_assert(K % 4 == 0 && LDA % 4 == 0 && LDW % 4 == 0, "alignment");
  const float* ap = As + ty * LDA;
  const float* wp = Ws + tx * LDW;
#pragma unroll 4
  for (int k = 0; k < K; k += 4) {
    float4 a[TM], w[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(ap + i * 32 * LDA + k);
#pragma unroll
    for (int j = 0; j < TN; ++j) w[j] = *reinterpret_cast<const float4*>(wp + j * 8 * LDW + k);
    // Issue order: the register file feeds one 64-bit operand per lane per cycle, so an FFMA2 sustains its
    // 2-cycle rate only if one operand pair sits in the operand-reuse cache -> keep a[i] fixed over a run of j,
    // and never put the two updates of one accumulator back to back (4.4-cycle dependent-issue latency).
#pragma unroll
    for (int i = 0; i < TM; ++i) {
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = ffma2(lo2(a[i]), lo2(w[j]), acc[i][j]);
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = ffma2(hi2(a[i]), hi2(w[j]), acc[i][j]);
    }
  }
}

template <int TM, int TN>
__device__ __forceinline__ void acc_zero(float2 (&acc)[TM][TN]) {
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = make_float2(0.f, 0.f);
}

// LayerNorm over the 64 columns of a row spread over the 8 tx lanes (TN = 8 values per lane).
// Two-pass (mean, then biased variance of the centred values), eps = 1e-5 like torch.nn.LayerNorm.
__device__ __forceinline__ void row_layernorm8(float (&v)[8], const float* __restrict__ g,
                                               const float* __restrict__ b, int tx) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float mean = s * (1.0f / 64.0f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] -= mean;
    q += v[j] * v[j];
  }
  q += __shfl_xor_sync(0xffffffffu, q, 1);
  q += __shfl_xor_sync(0xffffffffu, q, 2);
  q += __shfl_xor_sync(0xffffffffu, q, 4);
  const float rstd = rsqrtf(q * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = v[j] * rstd * g[tx + 8 * j] + b[tx + 8 * j];
}

}  // namespace dpdf
