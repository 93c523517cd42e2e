/*
 * dpdfnet_b200 - C ABI of the B200-native batched DPDFNet streaming engine.
 *
 * This library replaces, for many streams at once, the one seam through which the reference
 * touches its inference backend:
 *
 *     runtime.session.run([spec_e, state_out], {spec: f32[1,1,F,2], state_in: f32[S]})
 *         reference call sites: package/src/dpdfnet/api.py:98-101, 154-157
 *                               package/src/dpdfnet/stream.py:129-135
 *         producer:             package/src/dpdfnet/onnx_backend.py:81-99 (build_runtime_model)
 *
 * plus the host DSP that brackets it in the streaming path (causal windowed rfft before, irfft *
 * window + overlap-add after: stream.py:119-126, 138-156), which the engine fuses on the device.
 *
 * Conventions
 *   - All functions return 0 on success, a negative dpdf_status otherwise; the message is
 *     available from dpdf_last_error() (thread local).  Nothing throws across the ABI.
 *   - "device" pointers are CUDA device pointers on the engine's device, "host" pointers are
 *     ordinary host memory.  `cuda_stream` is a cudaStream_t passed as void* (NULL = default).
 *   - An engine handle is single-writer: calls on one handle must not overlap.
 *   - Streams live in numbered slots [0, max_streams).  `slot_ids` (int32[B], may be NULL meaning
 *     0..B-1) maps batch rows to slots; a slot may appear at most once per call.
 *   - `flags` (int32[B], may be NULL meaning all zero) are per-row DPDF_FLAG_* bits.
 */
#ifndef DPDFNET_B200_H_
#define DPDFNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPDF_ABI_VERSION 1

typedef struct dpdf_engine dpdf_engine;

enum dpdf_status {
  DPDF_OK = 0,
  DPDF_ERR_INVALID = -1,   /* bad argument / spec mismatch   (reference: ValueError)          */
  DPDF_ERR_WEIGHTS = -2,   /* malformed or incomplete blob   (reference: FileNotFound/Value)  */
  DPDF_ERR_CUDA = -3,      /* CUDA runtime failure           (reference: RuntimeError)        */
  DPDF_ERR_NOMEM = -4
};

/* Per-row step flags (see DESIGN.md "offline-exact schedule"). */
#define DPDF_FLAG_WARMUP 1    /* network not run: feature/c0/coef ring slots <- 0, GRU states kept   */
#define DPDF_FLAG_ZERO_FEAT 2 /* normalised features forced to 0 (offline look-ahead padding)        */
#define DPDF_FLAG_ZERO_SPEC 8 /* step_pcm: analysed spectrum taken as 0 (offline DF look-ahead pad)  */

/* Static model description; mirrors dpdfnet_b200/spec.py:ModelSpec (reference constructor
 * arguments: onnx_model/dpdfnet.py:522-565, onnx_model/dpdfnet_48khz_hr.py:586-630). */
typedef struct dpdf_spec {
  int32_t abi_version;    /* DPDF_ABI_VERSION */
  int32_t sample_rate;    /* 16000 | 48000 */
  int32_t win;            /* 320 | 960 */
  int32_t hop;            /* win / 2 */
  int32_t freq_bins;      /* win / 2 + 1 */
  int32_t n_blocks;       /* DPRNN blocks per branch (0,2,4,8) */
  int32_t hr48;           /* 1 = DPDFNet48HR (per-bin features and mask) */
  int32_t fe_feat;        /* 32 | 481 */
  int32_t fe[4];          /* frequency widths of e0..e3 */
  int32_t erb_strides[3]; /* erb_conv1..3 frequency strides */
  int32_t dec_up[3];      /* convt3, convt2, convt1 up-sampling factors (1 = plain conv) */
  int32_t erb_widths[32]; /* ERB band widths in bins (sum = freq_bins) */
  int32_t state_size;     /* S: floats of the reference's flat state vector */
} dpdf_spec;

/* Create an engine for `max_streams` slots on CUDA device `device` from a packed weight blob
 * (dpdfnet_b200/weights.py:pack_checkpoint).  Replaces onnx_backend.py:81-99. */
int dpdf_create(const dpdf_spec* spec, const void* weights, size_t weights_bytes,
                int32_t max_streams, int32_t device, dpdf_engine** out);
int dpdf_destroy(dpdf_engine* e);

/* Re-initialise slots (norm states <- mu0/s0, everything else <- 0).  `slots` is a HOST array;
 * NULL / n <= 0 resets every slot.  Replaces `state = runtime.init_state.copy()` (api.py:91,
 * stream.py:69) and the buffer clears of StreamEnhancer.reset (stream.py:62-72). */
int dpdf_reset(dpdf_engine* e, const int32_t* slots_host, int32_t n, void* cuda_stream);

/* One hop for B streams, ONNX-shaped: un-normalised spectra in and out, f32[B, F, 2]
 * (the graph multiplies by wnorm on the way in and 1/wnorm on the way out,
 * onnx_model/export_dpdfnet_to_onnx.py:21-25).  Device pointers, asynchronous on `cuda_stream`. */
int dpdf_step_spec(dpdf_engine* e, const float* spec_in, float* spec_out, const int32_t* slot_ids,
                   const int32_t* flags, int32_t B, void* cuda_stream);

/* One hop for B streams, fused DSP: `hop` new PCM samples per stream in, `hop` enhanced samples
 * out (delayed by one window + 4 frames like StreamEnhancer, stream.py:117-156).  Row b of
 * pcm_in / pcm_out starts at pcm_in + b*in_stride / pcm_out + b*out_stride (floats). */
int dpdf_step_pcm(dpdf_engine* e, const float* pcm_in, int64_t in_stride, float* pcm_out,
                  int64_t out_stride, const int32_t* slot_ids, const int32_t* flags, int32_t B,
                  void* cuda_stream);

/* T consecutive hops: step t reads pcm_in[b*in_stride + t*hop ...] and writes the same offset of
 * pcm_out.  One CUDA-graph replay per hop; flags apply to every hop (normally NULL). */
int dpdf_run_pcm(dpdf_engine* e, const float* pcm_in, int64_t in_stride, float* pcm_out,
                 int64_t out_stride, const int32_t* slot_ids, const int32_t* flags, int32_t B,
                 int32_t T, void* cuda_stream);

/* Store one hop per stream as analysis history without producing a frame (the reference emits
 * nothing until a full window has been buffered, stream.py:116). */
int dpdf_prime_pcm(dpdf_engine* e, const float* pcm_in, int64_t in_stride, const int32_t* slot_ids,
                   int32_t B, void* cuda_stream);

/* Host-buffer variants (the call a reference-side binding makes): pageable or pinned HOST
 * arrays, host->device and device->host copies included, synchronous on return. */
int dpdf_step_spec_host(dpdf_engine* e, const float* spec_in, float* spec_out,
                        const int32_t* slot_ids, const int32_t* flags, int32_t B);
int dpdf_step_pcm_host(dpdf_engine* e, const float* pcm_in, float* pcm_out, const int32_t* slot_ids,
                       const int32_t* flags, int32_t B);
int dpdf_run_pcm_host(dpdf_engine* e, const float* pcm_in, float* pcm_out, const int32_t* slot_ids,
                      int32_t B, int32_t T);

/* Pipelined host entry for callers that always have the next hop in hand (a server feeding fixed 10 ms packets): submit
 * returns at once with a ticket; the host->device copy of ticket t and the device->host copy of ticket t-1 overlap the
 * kernels of the neighbouring ticket on separate CUDA streams.  At most two tickets are in flight (submit blocks until
 * ticket t-2 has been delivered); pcm_in / pcm_out must stay valid until dpdf_wait(ticket) returns and should be pinned
 * (pageable memory makes the copies synchronous).  Same arithmetic as dpdf_step_pcm_host, hop for hop.  Do not mix with the
 * synchronous *_host calls while tickets are outstanding. */
int dpdf_submit_pcm_host(dpdf_engine* e, const float* pcm_in, float* pcm_out, const int32_t* slot_ids,
                         const int32_t* flags, int32_t B, int64_t* ticket);
int dpdf_wait(dpdf_engine* e, int64_t ticket);

/* Flat state vector of one slot in the reference layout (onnx_model/dpdfnet.py:737-746):
 * drop-in for StreamEnhancer._state / the `state_in`,`state_out` tensors.  HOST buffers of
 * dpdf_state_size() floats; synchronous. */
int dpdf_state_size(const dpdf_engine* e);
int dpdf_state_export(dpdf_engine* e, int32_t slot, float* flat_host);
int dpdf_state_import(dpdf_engine* e, int32_t slot, const float* flat_host);

/* Introspection for tests and benches. */
int dpdf_debug_tensor(dpdf_engine* e, const char* name, float* out_host, size_t max_floats,
                      size_t* numel_per_stream);    /* stage output of the last step, [B, numel] */
int dpdf_kernel_launches(const dpdf_engine* e);    /* kernels launched by the last step */
/* Options (all default to the measured-best choice): "graph" 0/1 one CUDA graph per hop;
 * "lanes" 0 = auto, 1..8 kernel-chain lanes per batched step; "free_lanes" 0/1 lanes of a multi-hop run replay their
 * own graphs on their own streams; "overlap" 0/1 (+ "overlap_max") post kernel overlapped with the intra sweep;
 * "decoder_fork" 0/1 decoder tails on forked streams; "intra_tc" / "sep_tc" / "gru_tc" 0 FFMA2 / 1 tcgen05 / 2 by batch
 * size (+ "intra_tc_min"); "intra_dup" 0 auto / 1 / 2 / 4 rows per stream in the tcgen05 intra-GRU tile (128 / D streams
 * per CTA); "intra_frag" 0/1 fragment form of that kernel where it runs 32 streams per CTA (two rows per stream, .16x128b
 * TMEM fragments; on, up to "frag_max" = 6 144 streams per step); "intra_sr" 0 / 1 / 2 split rows (hi | lo operand halves in the D rows of a stream, two MMA passes) never /
 * whenever D > 1 / with D = 4 only when the fragment form is off (default 2); "dfp_early" 0/1 df pathway conv on a forked
 * stream behind df_conv0 + k_df_combine (measured slower, off); "lane_min" smallest lane in streams (128), "sweep_prio" launch
 * priority of the sweep kernels (measured neutral, 0); "sep_tma" 0/1 tensor-core separable convs as the persistent TMA-fed kernel; "post_tc" 0/1; "intra_bt" 0/8/16/32 stream tile of the FFMA2 intra-GRU kernel; "c0_fp16" 0/1 c0 ring stored in half precision (switch only on freshly reset streams); "ana_nb" / "syn_sb" caps on the
 * streams per CTA of the analysis / synthesis kernels ("ana_force" / "syn_force" force a count, experiments only); "post_pf" L2 prefetch distance of the post kernel; "dfp_ps" 0/1
 * df pathway conv as pending partial sums (switch only on freshly reset streams); "pdl" 0 / 1 chain every kernel of a
 * hop with programmatic dependent launches / 2 every segment but the DPRNN stack; "tail_pdl" 0/1 the dense tail only; "intra_pdl" 0/1 the sweep of block i >= 1 as a programmatic dependent of the previous
 * post kernel; "post_pair" 0/1/2 post kernel as cta_group::2 CTA pairs (never / always / when not overlapped with its sweep;
 * measured slower, off); "post_res" 0/1 post kernel as a persistent kernel with resident weights when it runs after its sweep
 * (measured on par, off); "post_dual" 0/1/2 two tiles per 1024-thread CTA sharing one weight ring (measured slower, off); "gru_uc" 0 auto / 32 / 64 hidden units per CTA of the tcgen05 GRU(256) kernel; "dft_tc" 0 FFMA2 / 1 tcgen05 / 2 by batch size
 * (+ "dft_tc_min") framed DFT and inverse DFT + overlap-add as tensor-core GEMMs; "encoder_fork" 0/1 df encoder chain on a
 * forked stream; "stop_after" k enqueue only the first k kernels of a hop (profiling: tools/chain_profile.py, outputs invalid). */
int dpdf_set_option(dpdf_engine* e, const char* key, int32_t value);
int dpdf_time_kernels(dpdf_engine* e, int32_t B, int32_t iters, float* ms_out, const char** names_out,
                      int32_t max_entries, int32_t* n_entries); /* per-kernel CUDA-event timing */
/* ---- batched streaming resampler (device) ----------------------------------------------------------------
 * Replaces the per-chunk host resampling around the path (reference: audio.py:20-27 ensure_sample_rate ->
 * librosa.resample, called by stream.py:112,163-164 and api.py:86,110) for B streams in lock step:
 *     y[m] = sum_j x[j] * taps[ntaps/2 + m*down - j*up]
 * (the arithmetic of scipy.signal.resample_poly; the zero-phase FIR `taps`, odd length, is designed by the caller,
 * dpdfnet_b200/resample.py).  State per stream: the last few input samples, so chunked output == one-shot output.
 * `n_out` (host) receives the samples written per row; dpdf_resampler_pending() tells it in advance. */
typedef struct dpdf_resampler dpdf_resampler;
int dpdf_resampler_create(int32_t up, int32_t down, const float* taps_host, int32_t ntaps, int32_t max_streams,
                          int32_t device, dpdf_resampler** out);
int dpdf_resampler_destroy(dpdf_resampler* r);
int dpdf_resampler_reset(dpdf_resampler* r, void* cuda_stream);
int64_t dpdf_resampler_pending(const dpdf_resampler* r, int32_t n_new, int32_t flush);
int dpdf_resampler_process(dpdf_resampler* r, const float* in_dev, int64_t in_stride, int32_t n_new, float* out_dev,
                           int64_t out_stride, int32_t B, int32_t flush, int64_t* n_out, void* cuda_stream);

/* Errors raised ON THE DEVICE by earlier hops: a DPRNN post tile that timed out waiting for the overlapped sweep, or an
 * activation outside the FP16 operand range / non-finite in a tensor-core converter (the FP16 hi/lo split of
 * DESIGN.md section 3 is exact only inside +-65504).  Returns 0 or DPDF_ERR_CUDA with the message in dpdf_last_error()
 * and clears the condition.  `synchronize` != 0 waits for the device first.  The *_host entry points call it after
 * their own synchronisation (so they report errors of the hop they ran); the device-pointer entry points call it on
 * entry (so they report errors of hops that finished since the previous call).  The reference raises nothing
 * comparable: ONNX Runtime computes in FP32 throughout. */
int dpdf_poll_error(dpdf_engine* e, int32_t synchronize);

const char* dpdf_last_error(void);
const char* dpdf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DPDFNET_B200_H_ */
