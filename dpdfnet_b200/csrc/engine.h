// Engine-internal structures shared by the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/dpdfnet_b200.h"
#include "common.cuh"

namespace dpdf {

// ---- device views ---------------------------------------------------------------------------
struct Dims {
  int win, hop, F;        // window, hop, frequency bins
  int fe_feat;            // 32 | 481
  int fe[4];              // widths of e0..e3
  int stride[3];          // erb_conv1..3
  int up[3];              // convt3, convt2, convt1
  int N;                  // DPRNN blocks
  int hr48;
  float wnorm, inv_wnorm;
};

struct State {           // all [max_streams][...], float unless noted
  float* mu;             // [fe_feat]
  float* s;              // [96]
  float* erb_ring;       // [3][fe_feat]
  float* df_ring;        // [3][2][96]
  float* inter_erb;      // [N][fe3][64]
  float* inter_df;       // [N][48][64]
  float* h_enc;          // [256]
  float* h_erb;          // [2][256]
  float* h_df;           // [2][256]
  float* c0_ring;        // [5][96][64]
  float* dfp_acc;        // [96][5][10] pending sums of the df pathway conv (k_df_pathway_ps); slot (pos + d) % 5 = output d hops ahead
  float* mask_ring;      // [3][F][2]
  float* coef_ring;      // [3][96][10]
  float* dfspec_ring;    // [5][F][2]
  float* in_hist;        // [hop]
  float* ola;            // [hop]
  int* pos;              // frames pushed so far
  int c0_fp16;           // c0_ring holds FP16 frames [5][96][64] halves at the start of each slot's (FP32-sized) region (option "c0_fp16")
};

struct Scratch {         // all [max_streams][...]
  float *e0, *e1, *e2, *e3, *c0, *c1, *xe;
  float *hcat_e, *hcat_d;
  float *emb_e, *cemb, *g0, *henc, *emb;
  float *x1, *herb1, *herb2, *ed, *ed2;
  float *x2, *hdf1, *hdf2, *cc, *co, *dfp;     // dfp: pathway term of the df coefficients [96][10] (k_df_pathway early form)
  float *d3, *d2, *d1, *m;
  float *spec_tc, *yspec_tc;      // k_dft_tc: spectrum of the hop [spec_tc_ld], masked / filtered spectrum Y [yspec_tc_ld] (zero padded)
};

struct SepW { const float *dw, *pw, *b, *tc_pw; };     // tc_pw: FP16 hi/lo tcgen05 image of pw (weights.py:umma_operand16)
struct GLW { const float *w, *b; int G, Ng, Kg; };
struct GRUW { const float *wih, *whh, *bias, *tc_w; };   // tc_w: FP16 hi/lo slab images for k_gru_tc (weights.py:gru_tc_images)
struct DprnnW {
  const float *i_wih, *i_whh, *i_bias, *fc_w, *fc_b, *ln_g, *ln_b;
  const float *r_wih, *r_whh, *r_bias, *fc2_w, *fc2_b, *ln2_g, *ln2_b;
  const float *tc_fc_w, *tc_gates, *tc_fc2_w;      // tcgen05 operand images (hi | lo), weights.py:umma_operand
  const float* tc_intra_bias;                      // [2][4][64] with the same exponent scales as the images
  const float* tc_intra;                           // FP16 operand images of the intra GRU, weights.py:umma_operand16
  const float* tc_intra_f = nullptr;               // df branch: the same with W_hh's K axis in fragment order (k_dprnn_intra_tc.cu:intra_sweep_f)
};

struct Weights {
  const float *dft_fwd, *dft_inv, *mu0, *s0;
  const float* dft_tc_scale;                 // [4] power-of-two (in, out) scales of the analysis and of the synthesis GEMM
  const float *dft_fwd_tc, *dft_inv_tc;      // FP16 hi | lo operand images of the two bases (weights.py:dft_tc_images), or nullptr (older blobs)
  const float *erb_conv0_w, *erb_conv0_b;
  SepW erb_conv[3], df_conv1, convt[3];
  const float *df_conv0_w, *df_conv0_pw, *df_conv0_b, *df_conv0_tc_pw;
  std::vector<DprnnW> dprnn_erb, dprnn_df;
  GLW erb_fc_emb, df_fc_emb, enc_in, enc_out, erbdec_in, erbdec_out, erbdec_fc, dfdec_in, df_skip, df_out;
  GRUW enc_gru, erb_gru[2], df_gru[2];
  const float *convp_a[4], *convp_b[4];      // conv3p, conv2p, conv1p, conv0p
  const float *conv0_out_w, *conv0_out_b;
  const float *dfp_w, *dfp_pw, *dfp_b;
  const float* band_inv_w;                   // [32] 1/width
  const int* band_start;                     // [33]
  const int* band_of_bin;                    // [F]
};

// ---- kernel launchers (defined in k_*.cu) ---------------------------------------------------
struct Engine;

void launch_prime(Engine& e, const float* pcm, long long stride, const int* slot_ids, int B, cudaStream_t st);
void launch_analysis(Engine& e, int B, cudaStream_t st);
bool dft_on_tc(const Engine& e, int B);
void launch_dft_tc(Engine& e, int B, cudaStream_t st);     // before launch_analysis when dft_on_tc
void launch_idft_tc(Engine& e, int B, cudaStream_t st);    // after launch_synthesis when dft_on_tc
void init_dft_tc_kernels();
void launch_synthesis(Engine& e, int B, cudaStream_t st);
void launch_reset(Engine& e, const int* slots_dev, int n, cudaStream_t st);

void launch_erb_conv0(Engine& e, int B, cudaStream_t st);

struct SepProblem {
  int mode;                 // 0: depthwise(+pathway) prologue, 1: df_conv0 grouped 3x3 from the df ring
  const float* in1;         // [B][Fin][64]
  const float* in2;         // pathway source [B][Fin][64] or nullptr
  const float *pa, *pb;     // pathway affine
  const float *dw, *pw, *bias;
  const float* tc_pw;       // tensor-core image of pw (launch_sepconv_tc)
  float* out;               // [B][Fout][64]  (mode 1: c0 ring slot)
  int Fin, Fout, stride, up;
  int tile0;                // first tile index of this problem in the launch
};
void launch_sepconv(Engine& e, const SepProblem* probs, int nprob, int B, cudaStream_t st);
void launch_sepconv_tc(Engine& e, const SepProblem* probs, int nprob, int B, cudaStream_t st);
void launch_sepconv_tma(Engine& e, const SepProblem* probs, int nprob, int B, cudaStream_t st);   // persistent, TMA-fed (k_conv_tma.cu)
bool sepconv_tma_available();
void launch_conv0_out(Engine& e, int B, cudaStream_t st);
void launch_df_pathway(Engine& e, int B, cudaStream_t st);
void launch_df_pathway_early(Engine& e, int B, cudaStream_t st);   // pathway term alone -> Scratch::dfp
void launch_df_combine(Engine& e, int B, cudaStream_t st);         // coefficient ring <- tanh(df_out) + Scratch::dfp
void launch_df_pathway_ps(Engine& e, int B, cudaStream_t st);

void launch_dprnn_intra(Engine& e, int blk, int B, cudaStream_t st);
void launch_dprnn_post(Engine& e, int blk, int B, cudaStream_t st);
void launch_dprnn_post_tc(Engine& e, int blk, int B, cudaStream_t st);
void launch_dprnn_intra_tc(Engine& e, int blk, int B, cudaStream_t st);
int intra_tc_dup(const Engine& e, int B);   // stream tiles of the sweep are 128 / D streams
int intra_tc_dup_erb(const Engine& e, int B);   // ... of the erb branch: 1, or 4 (fragment form, 48 kHz models)

struct GLProblem {
  const float* in0; int ld0;     // input columns [0, split)
  const float* in1; int ld1;     // input columns [split, ...) (may be nullptr)
  int split;
  GLW w;
  float* out; int ldo; int col0; // output written at out[b*ldo + col0 + ...]
  const float* addend; int lda;  // optional: y += addend[b*lda + col]
  int act;                       // 0 none, 1 relu, 2 tanh
};
void launch_gl(Engine& e, const GLProblem* probs, int nprob, int B, cudaStream_t st);

struct GRUProblem {
  const float* x;   // [B][256]
  float* hstate;    // [max_streams][hs_stride] slot-indexed
  int hs_stride;
  GRUW w;
  float* hout;      // [B][256]
};
void launch_gru(Engine& e, const GRUProblem* probs, int nprob, int B, cudaStream_t st);
void launch_gru_tc(Engine& e, const GRUProblem* probs, int nprob, int B, cudaStream_t st);
void launch_gru_commit(Engine& e, const GRUProblem* probs, int nprob, int B, cudaStream_t st);   // nprob <= 5

// ---- the engine -----------------------------------------------------------------------------
struct Engine {
  dpdf_spec spec;
  Dims d;
  int device = 0;
  int max_streams = 0;
  int num_sms = 148;
  float* weights_dev = nullptr;
  size_t weights_floats = 0;
  std::map<std::string, std::pair<size_t, size_t>> wtable;   // name -> (offset, numel)
  Weights w;
  State st;
  Scratch sc;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  int* aux_int = nullptr;         // band tables
  IoDesc* io_dev = nullptr;       // descriptor the kernel launchers bind (points into io_lanes during a laned step)
  IoDesc* io_lanes = nullptr;     // [MAX_LANES] device descriptors, one per lane
  IoDesc* io_host = nullptr;      // pinned
  int* slots_dev = nullptr;       // staging for host slot ids / flags
  int* flags_dev = nullptr;
  float* stage_in = nullptr;      // device staging for *_host entry points
  float* stage_out = nullptr;
  size_t stage_in_floats = 0, stage_out_floats = 0;
  // 2-deep host pipeline (dpdf_submit_pcm_host / dpdf_wait): H2D, hop and D2H of consecutive tickets on three streams
  struct HostPipe {
    float *in[2] = {nullptr, nullptr}, *out[2] = {nullptr, nullptr};
    int *slots[2] = {nullptr, nullptr}, *flags[2] = {nullptr, nullptr};
    size_t cap = 0;               // floats per staging buffer
    int cap_ids = 0;
    cudaEvent_t h2d_done[2] = {}, step_done[2] = {}, d2h_done[2] = {};
    cudaStream_t in_stream = nullptr, out_stream = nullptr;
    long long submitted = 0;
  } pipe;
  float* pinned = nullptr;        // pinned host staging
  size_t pinned_floats = 0;
  cudaStream_t own_stream = nullptr;
  int last_B = 0;
  int launches = 0;               // kernels launched by the last step
  // Lanes: a batched step is split into `lanes` row ranges that run as independent kernel chains on forked streams
  // inside one CUDA graph, so the latency-bound kernels of one lane (the sequential intra-frame GRU sweep occupies
  // 32 SMs per 1024 streams) overlap with the throughput kernels of the others.  Lanes share the state arena (slots
  // are global) and use disjoint row ranges of the scratch arena.
  static constexpr int MAX_LANES = 16;
  int sweep_prio = 0;             // launch priority of the sweep kernels (0 = default, v = -v): a sweep CTA needs a whole SM, lanes' small kernels fragment them
  int prio_now = 0;               // priority of the launch being enqueued (launch_k)
  int lane_min = 128;             // smallest lane (streams); lanes are multiples of it
  int lanes = 0;                  // 0 = auto by batch size
  int total_B = 0;                // batch of the whole step while its lanes are enqueued (kernel-variant choice)
  std::vector<std::pair<float**, size_t>> sc_items;   // scratch pointer members and their floats per stream
  cudaStream_t lane_stream[MAX_LANES] = {};  // [0] is used by free-running multi-hop runs only
  cudaEvent_t lane_fork = nullptr, lane_done[MAX_LANES] = {};
  int use_graph = 1;
  int intra_bt = 0;               // 0 = auto
  int intra_tc = 2;               // intra-frame GRU on tcgen05 (FP16 split): 0 never, 1 always, 2 = when B >= intra_tc_min
  int intra_tc_min = 1;           // with the fragment form (32-stream CTAs) the tcgen05 sweep wins at every batch size (profiles/r4a_*, r4b_*, dpdfnet4 ms/hop FFMA2 / tcgen05: 4 streams 0.595 / 0.491, 64: 0.616 / 0.514, 512: 0.738 / 0.585; dpdfnet8_48khz_hr 512: 1.96 / 1.25); before it the cross-over was 640 streams (profiles/r2v_sweep.log)
  int intra_pdl = 0;              // sweep of block i >= 1 launched as a programmatic dependent of the previous block's post kernel (prologue under its tail)
  int intra_sr = 2;               // k_dprnn_intra_tc "split rows": the D rows of a stream carry the hi | lo operand halves, two MMA passes instead of three; 0 off, 1 whenever D > 1, 2 = with D = 4 only (measured)
  int dfp_early = 0;              // df pathway conv on a forked stream right behind df_conv0 (needs encoder_fork), k_df_combine + gru_commit on the coefficient tail; measured +0.5..0.8 % hop time (the ERB tail is as long): off
  int frag_max = 6144;            // largest step (streams, all lanes) whose sweeps run in fragment form
  int intra_frag_erb = 1;         // ... and for the erb branch of the 48 kHz models while both sweeps fit one wave (intra_tc_dup_erb)
  int intra_frag = 1;             // k_dprnn_intra_tc fragment form where the sweep runs 32 streams per CTA (intra_dup 4): two rows per stream, .16x128b TMEM fragments
  int intra_dup = 0;              // k_dprnn_intra_tc row duplication D (128 / D streams per CTA): 0 = auto (largest D whose sweep fits one wave), 1, 2, 4
  int gru_tc = 2;                 // GRUCell(256) gate GEMMs on tcgen05: 0 never, 1 always, 2 = when B >= gru_tc_min
  int gru_tc_min = 1;             // since the recurrent half runs ahead of the grid dependency and the MMAs issue under elect.sync the tcgen05 cells win at every batch size (profiles/r4d_*: 1 / 64 / 192 streams 0.463 / 0.515 / 0.562 -> 0.381 / 0.432 / 0.487 ms per hop; 256 before)
  int dft_tc = 2;                 // framed DFT / inverse DFT + OLA on tcgen05 (k_dft_tc.cu): 0 never, 1 always, 2 = when B >= dft_tc_min
  int dft_tc_min = 512;
  int spec_tc_ld = 0, yspec_tc_ld = 0;       // floats per stream of the two scratch rows
  int gru_uc = 0;                 // k_gru_tc hidden units per CTA: 0 = auto (32 while the doubled grid fits one wave), 32, 64
  int sep_tc = 2;                 // separable convs with the pointwise GEMM on tcgen05: 0 never, 1 always, 2 = when B >= sep_tc_min
  int sep_tc_min = 256;
  int sep_tma = 1;                // tensor-core separable convs as the persistent TMA-fed kernel (k_conv_tma.cu) instead of k_sepconv_tc
  int ana_force = 0, syn_force = 0;   // experiments: force that many streams per CTA regardless of the grid size (0 = auto)
  int ana_nb = 32, syn_sb = 16;   // caps on the streams per CTA of the analysis / synthesis kernels (8|16|32, 4|8|16)
  int post_pf = 1;                // k_dprnn_post_tc: L2 prefetch distance in units of the SM count (2 CTAs per SM -> 2), 0 = off
  int stop_after = 0, run_idx = 0;   // profiling (tools/chain_profile.py): enqueue only the first stop_after kernels of a hop
  int post_dual = 0;              // k_dprnn_post_tc<2>: two tiles per 1024-thread CTA sharing a 5-slab weight ring: 0 never, 1 always, 2 = after the sweep only.
                                  // Bit-identical, parity-tested, measured 33 % SLOWER (profiles/r2Q_*: post 2.53 -> 3.37 ms per hop at 16 384 streams):
                                  // no weight wait is left, but the two tiles advance in lock step - what makes two co-resident CTAs fast is that
                                  // they are NOT synchronised: one tile's epilogue runs under the other's MMAs and loads
  int post_res = 0;               // k_dprnn_post_res (persistent, resident weights) when the post kernel runs after its sweep.  Bit-identical,
                                  // parity-tested, measured on par / slightly slower (profiles/r2E_*: post 2.53 -> 2.58 ms per hop at 16 384
                                  // streams): without weight waits a tile takes 21-26 k cycles, but one CTA per SM overlaps nothing, and two
                                  // co-resident streaming CTAs at 37-44 k cycles each come to the same 19-22 k per tile
  int* post_ctr_dev = nullptr;    // [MAX_LANES][4] tile counters of k_dprnn_post_res (self-resetting)
  int post_pair = 0;              // k_dprnn_post_tc as CTA pairs (cta_group::2, M = 256): 0 never, 1 always, 2 = when not overlapped with its sweep.
                                  // Bit-identical and parity-tested, but measured SLOWER (profiles/r2A_*: post 2.54 -> 2.79 ms per hop at 16 384
                                  // streams): the 4-deep ring does cut phase 2 (15.0 k -> 8.2 k cycles per tile) but the pair runs in lock step -
                                  // every phase waits for the slower CTA's staging plus a remote mbarrier arrive (phase 1: 2.0 k -> 8.2 k), and
                                  // every slab needs a relay hop from the peer (a 1-D bulk copy cannot signal the leader's barrier)
  int post_tc = 1;                // DPRNN position-parallel half on tcgen05 (3xTF32) instead of FFMA2
  std::map<int, cudaGraphExec_t> graphs;     // keyed by B: one hop of all lanes (forked chains, joined)
  std::map<int, cudaGraphExec_t> lane_graphs; // keyed by B * MAX_LANES + lane: one hop of one lane (free-running lanes of a multi-hop run)
  int launches_per_lane = 0;
  // post kernel overlapped with the intra sweep: per-lane progress counters [lane][2][2][tiles]
  int decoder_fork = 1;           // run the deep-filter coefficient tail beside the ERB decoder's conv stack (forked stream)
  cudaStream_t br_stream[MAX_LANES] = {};
  cudaStream_t dfp_stream[MAX_LANES] = {};          // the early df pathway conv (Engine::dfp_early) runs beside the DPRNN stack
  cudaEvent_t dfp_fork[MAX_LANES] = {}, dfp_join[MAX_LANES] = {};
  cudaEvent_t br_fork[MAX_LANES] = {}, br_join[MAX_LANES] = {};
  int encoder_fork = 1;           // df encoder chain (df_conv0 -> df_conv1) on the forked stream beside erb_conv0..3
  cudaEvent_t enc_fork[MAX_LANES] = {}, enc_join[MAX_LANES] = {};
  int dfp_ps = 0;                 // df pathway conv as pending partial sums: 38 KB of accumulator traffic instead of the 120 KB c0 ring
                                  // read per stream-hop, but measured slower (0.48 vs 0.27 ms at 8192 streams: the 50-value
                                  // reduce-scatter and the dependent read-modify-writes cost more than the ring read saves)
  std::vector<float> dfp_w_host;  // [10][5][32] for rebuilding the pending sums at state import
  int tail_pdl = 1;               // option (measured -0.6 % hop time at 1024 and 16 384 streams, profiles/r2s_sweep.log): programmatic dependent launches for the dense per-stream tail only (grouped linears + GRU
                                  // cells between the last DPRNN block and the decoder fork): eight tiny dependent kernels whose
                                  // prologues and weight prefetches can run under their predecessor
  int pdl = 0;                    // option: chain ALL kernels of a hop with programmatic dependent launches (measured slower: early-resident waiters crowd the running kernel)
  bool pdl_now = false;           // decided per enqueue_step (off while timing with events)
  bool pdl_first = false;         // next launch is the first kernel of a chain: plain launch
  int overlap = 1;                // option: 0 off, 1 on whenever both DPRNN kernels run on tcgen05, lanes are not forced and
  int overlap_max = 4097;         //         the step has fewer than overlap_max streams
  bool overlap_now = false;       // decided per enqueue_step
  int cur_lane = 0;
  int* progress_dev = nullptr;
  int progress_tiles = 0;
  // Device-raised error words (IoDesc::err) in mapped pinned host memory: the kernels store, the host polls without
  // a copy (api.cu:check_device_errors).  [0] overlap wait timed out, [1] FP16 operand range exceeded / non-finite.
  int* err_host = nullptr;
  int* err_dev = nullptr;
  int free_lanes = 1;             // multi-hop runs: every lane replays its own graph on its own stream, joined once at the end
  std::vector<std::pair<std::string, float>> ktimes;
  bool timing = false;
  std::vector<cudaEvent_t> tev;
  std::vector<std::string> tnames;
};

// Kernel launch with the programmatic-stream-serialization attribute when the engine runs its hop as a PDL chain
// (Engine::pdl_now); `first` marks the first kernel of a chain, whose predecessor is not a kernel.
template <typename... KArgs, typename... Args>
inline void launch_k(const Engine& e, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args);

int set_error(int code, const char* msg);   // thread-local message behind dpdf_last_error()
template <typename... KArgs, typename... Args>
inline void launch_k(const Engine& e, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (e.pdl_now && !e.pdl_first) {
    attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    ++cfg.numAttrs;
  }
  if (e.prio_now) {                                        // kernel-node priority (Engine::sweep_prio): lower value = scheduled first
    attr[cfg.numAttrs].id = cudaLaunchAttributePriority;
    attr[cfg.numAttrs].val.priority = -e.prio_now;
    ++cfg.numAttrs;
  }
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

void init_frontend_kernels();
void init_conv_kernels();
void init_dprnn_kernels();
void init_dense_kernels();
void init_dprnn_tc_kernels();
void init_dprnn_intra_tc_kernels();
void init_conv_tc_kernels();
void init_conv_tma_kernels();
void init_gru_tc_kernels();
void enqueue_step(Engine& e, int B, cudaStream_t st);    // all kernels of one hop, in order

}  // namespace dpdf
