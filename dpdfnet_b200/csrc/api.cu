// C ABI of the engine (include/dpdfnet_b200.h): creation from a packed weight blob, the per-hop
// kernel schedule (optionally replayed as a CUDA graph), host-buffer entry points and the
// reference-layout state import / export.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cuda_fp16.h>

#include "engine.h"

using namespace dpdf;

struct dpdf_engine { Engine e; };

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t err__ = (call);                                                                   \
    if (err__ != cudaSuccess)                                                                     \
      return fail(DPDF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

namespace dpdf {
int set_error(int code, const char* msg) {      // for the other translation units behind the same ABI
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
}  // namespace dpdf

extern "C" const char* dpdf_last_error(void) { return g_err; }
extern "C" const char* dpdf_version(void) { return "dpdfnet_b200 0.2 (sm_100a, tcgen05 FP16-split + FFMA2)"; }

// ---------------------------------------------------------------------------------------------
// weight blob
// ---------------------------------------------------------------------------------------------
static const int NAME_LEN = 56;

static int parse_blob(Engine& e, const void* blob, size_t nbytes, const float** payload, size_t* payload_floats) {
  const char* p = static_cast<const char*>(blob);
  if (nbytes < 16 || memcmp(p, "DPDFW001", 8) != 0) return fail(DPDF_ERR_WEIGHTS, "not a DPDFNet-B200 weight blob");
  long long n;
  memcpy(&n, p + 8, 8);
  const size_t rec = NAME_LEN + 16;
  if (n <= 0 || (unsigned long long)n > (nbytes - 16) / rec)        // before any n * rec product: no wrap-around
    return fail(DPDF_ERR_WEIGHTS, "truncated weight blob header");
  size_t head = 16 + (size_t)n * rec;
  head += (128 - head % 128) % 128;
  if (head > nbytes) return fail(DPDF_ERR_WEIGHTS, "truncated weight blob header");
  const size_t avail = (nbytes - head) / sizeof(float);             // payload floats actually present
  size_t maxend = 0;
  for (long long i = 0; i < n; ++i) {
    const char* r = p + 16 + i * rec;
    char name[NAME_LEN + 1];
    memcpy(name, r, NAME_LEN);
    name[NAME_LEN] = 0;
    long long off, numel;
    memcpy(&off, r + NAME_LEN, 8);
    memcpy(&numel, r + NAME_LEN + 8, 8);
    if (off < 0 || numel < 0 || (size_t)off > avail || (size_t)numel > avail - (size_t)off)   // overflow-safe bounds
      return fail(DPDF_ERR_WEIGHTS, "corrupt or truncated entry %s", name);
    e.wtable[name] = {(size_t)off, (size_t)numel};
    if ((size_t)off + (size_t)numel > maxend) maxend = (size_t)off + (size_t)numel;
  }
  if (maxend > avail) return fail(DPDF_ERR_WEIGHTS, "truncated weight blob payload");
  *payload = reinterpret_cast<const float*>(p + head);
  *payload_floats = (nbytes - head) / sizeof(float);
  return 0;
}

static int bind_weights(Engine& e) {
  const char* missing = nullptr;
  static std::string missing_name;
  auto W = [&](const std::string& name, size_t expect) -> const float* {
    auto it = e.wtable.find(name);
    if (it == e.wtable.end() || it->second.second != expect) {
      if (!missing) { missing_name = name; missing = missing_name.c_str(); }
      return nullptr;
    }
    return e.weights_dev + it->second.first;
  };
  const Dims& d = e.d;
  Weights& w = e.w;
  w.dft_fwd = W("const.dft_fwd", (size_t)d.win * d.F * 2);
  w.dft_inv = W("const.dft_inv", (size_t)d.win * d.F * 2);
  {
    // optional (blobs packed before the tensor-core DFT existed lack them): without them the FFMA2 DFT stays
    const size_t ncol = (size_t)(2 * d.F + 127) / 128 * 128, kpad = (size_t)(2 * d.F + 63) / 64 * 64;
    auto itf = e.wtable.find("const.dft_fwd_tc"), iti = e.wtable.find("const.dft_inv_tc"), its = e.wtable.find("const.dft_tc_scale");
    const bool ok = itf != e.wtable.end() && iti != e.wtable.end() && its != e.wtable.end() && its->second.second == 4 && itf->second.second == ncol * d.win &&
                    iti->second.second == (size_t)(d.hop / 80) * (kpad / 64) * 160 * 64 && d.hop % 80 == 0 && d.win % 64 == 0 && d.win <= 1024;
    w.dft_fwd_tc = ok ? e.weights_dev + itf->second.first : nullptr;
    w.dft_inv_tc = ok ? e.weights_dev + iti->second.first : nullptr;
    w.dft_tc_scale = ok ? e.weights_dev + its->second.first : nullptr;
    e.spec_tc_ld = (int)ncol;
    e.yspec_tc_ld = (int)kpad;
    if (!ok) e.dft_tc = 0;
  }
  w.mu0 = W("const.mu0", d.fe_feat);
  w.s0 = W("const.s0", NDF);
  w.erb_conv0_w = W("enc.erb_conv0.w", 9 * C);
  w.erb_conv0_b = W("enc.erb_conv0.b", C);
  auto sep = [&](const std::string& n, int up) { return SepW{W(n + ".dw", (size_t)up * 3 * C), W(n + ".pw", C * C), W(n + ".b", C), W(n + ".tc_pw", C * C)}; };
  for (int i = 0; i < 3; ++i) w.erb_conv[i] = sep("enc.erb_conv" + std::to_string(i + 1), 1);
  w.df_conv0_w = W("enc.df_conv0.w", 9 * C);
  w.df_conv0_pw = W("enc.df_conv0.pw", C * C);
  w.df_conv0_b = W("enc.df_conv0.b", C);
  w.df_conv0_tc_pw = W("enc.df_conv0.tc_pw", C * C);
  w.df_conv1 = sep("enc.df_conv1", 1);
  for (int br = 0; br < 2; ++br) {
    auto& vec = br ? w.dprnn_df : w.dprnn_erb;
    vec.resize(d.N);
    for (int i = 0; i < d.N; ++i) {
      const std::string q = std::string("enc.dprnn_") + (br ? "df." : "erb.") + std::to_string(i);
      DprnnW& x = vec[i];
      x.i_wih = W(q + ".intra.wih", 2 * 192 * C);  x.i_whh = W(q + ".intra.whh", 2 * 192 * C);
      x.i_bias = W(q + ".intra.bias", 2 * 4 * C);
      x.fc_w = W(q + ".intra.fc_w", C * 2 * C);    x.fc_b = W(q + ".intra.fc_b", C);
      x.ln_g = W(q + ".intra.ln_g", C);            x.ln_b = W(q + ".intra.ln_b", C);
      x.r_wih = W(q + ".inter.wih", 192 * C);      x.r_whh = W(q + ".inter.whh", 192 * C);
      x.r_bias = W(q + ".inter.bias", 4 * C);
      x.fc2_w = W(q + ".inter.fc_w", C * C);       x.fc2_b = W(q + ".inter.fc_b", C);
      x.ln2_g = W(q + ".inter.ln_g", C);           x.ln2_b = W(q + ".inter.ln_b", C);
      x.tc_fc_w = W(q + ".tc.fc_w", 2 * C * C);     x.tc_gates = W(q + ".tc.gates", 6 * C * C);
      x.tc_fc2_w = W(q + ".tc.fc2_w", C * C);
      x.tc_intra = W(q + ".tc.intra", 2 * 4 * 192 * C / 2);
      x.tc_intra_bias = W(q + ".tc.intra_bias", 2 * 4 * C);
      {                                                      // optional (blobs packed before the fragment form existed): absent = form off
        auto it = e.wtable.find(q + ".tc.intra_f");
        x.tc_intra_f = it != e.wtable.end() && it->second.second == (size_t)(2 * 4 * 192 * C / 2) ? e.weights_dev + it->second.first : nullptr;
      }
    }
  }
  auto gl = [&](const std::string& n, int G, int Ng, int Kg) { return GLW{W(n + ".w", (size_t)G * Ng * Kg), W(n + ".b", (size_t)G * Ng), G, Ng, Kg}; };
  auto gru = [&](const std::string& n) { return GRUW{W(n + ".wih", 3 * H * H), W(n + ".whh", 3 * H * H), W(n + ".bias", 4 * H), W(n + ".tc_w", 2 * 3 * H * H)}; };
  const int kfc = C * d.fe[3] / 32;
  if (d.hr48) {
    w.erb_fc_emb = gl("enc.erb_fc_emb", 32, 16, kfc);
    w.erbdec_fc = gl("erb_dec.erb_fc_emb", 32, kfc, 16);
  }
  w.df_fc_emb = gl("enc.df_fc_emb", 32, 16, 96);
  w.enc_in = gl("enc.emb_gru.lin_in", 16, 16, 64);
  w.enc_gru = gru("enc.emb_gru.gru.0");
  w.enc_out = gl("enc.emb_gru.lin_out", 16, 32, 16);
  w.erbdec_in = gl("erb_dec.emb_gru.lin_in", 16, 16, 32);
  w.erb_gru[0] = gru("erb_dec.emb_gru.gru.0");
  w.erb_gru[1] = gru("erb_dec.emb_gru.gru.1");
  w.erbdec_out = gl("erb_dec.emb_gru.lin_out", 16, 32, 16);
  for (int i = 0; i < 3; ++i) {
    const std::string n = std::to_string(3 - i);
    w.convp_a[i] = W("erb_dec.conv" + n + "p.a", C);
    w.convp_b[i] = W("erb_dec.conv" + n + "p.b", C);
    w.convt[i] = sep("erb_dec.convt" + n, d.up[i]);
  }
  w.convp_a[3] = W("erb_dec.conv0p.a", C);
  w.convp_b[3] = W("erb_dec.conv0p.b", C);
  w.conv0_out_w = W("erb_dec.conv0_out.w", 3 * C);
  w.conv0_out_b = W("erb_dec.conv0_out.b", 1);
  w.dfp_w = W("df_dec.df_convp.w", 10 * ORD * 32);
  w.dfp_pw = W("df_dec.df_convp.pw", 100);
  w.dfp_b = W("df_dec.df_convp.b", 10);
  w.dfdec_in = gl("df_dec.df_gru.lin_in", 8, 32, 64);
  w.df_gru[0] = gru("df_dec.df_gru.gru.0");
  w.df_gru[1] = gru("df_dec.df_gru.gru.1");
  w.df_skip = gl("df_dec.df_skip", 16, 16, 32);
  w.df_out = gl("df_dec.df_out", 16, 60, 16);
  if (missing) return fail(DPDF_ERR_WEIGHTS, "weight tensor '%s' missing or of unexpected size", missing);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------------------------
static int check_spec(const dpdf_spec* s) {
  if (!s) return fail(DPDF_ERR_INVALID, "spec is NULL");
  if (s->abi_version != DPDF_ABI_VERSION) return fail(DPDF_ERR_INVALID, "ABI version mismatch: %d vs %d", s->abi_version, DPDF_ABI_VERSION);
  if (s->win != 2 * s->hop || s->freq_bins != s->win / 2 + 1 || s->win % 4 || s->win > 1024)
    return fail(DPDF_ERR_INVALID, "unsupported framing win=%d hop=%d bins=%d", s->win, s->hop, s->freq_bins);
  if (s->n_blocks < 0 || s->n_blocks > 64) return fail(DPDF_ERR_INVALID, "bad n_blocks %d", s->n_blocks);
  int sum = 0;
  for (int i = 0; i < 32; ++i) sum += s->erb_widths[i];
  if (sum != s->freq_bins) return fail(DPDF_ERR_INVALID, "ERB widths sum to %d, expected %d", sum, s->freq_bins);
  if (s->hr48 ? (s->fe_feat != s->freq_bins || s->fe[0] != s->freq_bins - 1) : (s->fe_feat != 32 || s->fe[0] != 32))
    return fail(DPDF_ERR_INVALID, "inconsistent feature widths");
  if ((C * s->fe[3]) % 32) return fail(DPDF_ERR_INVALID, "fe[3] not compatible with 32 linear groups");
  return 0;
}

static size_t state_size_of(const Dims& d) {
  return (size_t)d.fe_feat + NDF + 3 * d.fe_feat + (size_t)d.N * d.fe[3] * C + 3 * 2 * NDF + (size_t)d.N * (NDF / 2) * C +
         5 * H + (size_t)ORD * C * NDF + 3 * d.F * 2 + 3 * ORD * NDF * 2 + (size_t)ORD * d.F * 2;
}

extern "C" int dpdf_create(const dpdf_spec* spec, const void* weights, size_t nbytes, int32_t max_streams,
                           int32_t device, dpdf_engine** out) {
  if (!out) return fail(DPDF_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (int rc = check_spec(spec)) return rc;
  if (!weights) return fail(DPDF_ERR_WEIGHTS, "weights is NULL");
  if (max_streams <= 0) return fail(DPDF_ERR_INVALID, "max_streams must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(DPDF_ERR_CUDA, "no CUDA device available: this engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(DPDF_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(DPDF_ERR_CUDA, "device %s is sm_%d%d; the engine is built for sm_100a only", prop.name, prop.major, prop.minor);

  dpdf_engine* h = new dpdf_engine();
  Engine& e = h->e;
  e.spec = *spec;
  e.device = device;
  e.max_streams = max_streams;
  e.num_sms = prop.multiProcessorCount;
  Dims& d = e.d;
  d.win = spec->win; d.hop = spec->hop; d.F = spec->freq_bins; d.fe_feat = spec->fe_feat;
  for (int i = 0; i < 4; ++i) d.fe[i] = spec->fe[i];
  for (int i = 0; i < 3; ++i) { d.stride[i] = spec->erb_strides[i]; d.up[i] = spec->dec_up[i]; }
  d.N = spec->n_blocks; d.hr48 = spec->hr48;
  const double wn = 1.0 / ((double)d.win * d.win / (2.0 * d.hop));
  d.wnorm = (float)wn; d.inv_wnorm = (float)(1.0 / wn);
  if ((int)state_size_of(d) != spec->state_size) {
    delete h;
    return fail(DPDF_ERR_INVALID, "state_size mismatch: spec says %d, layout gives %zu", spec->state_size, state_size_of(d));
  }
  auto bail = [&](int rc) { dpdf_destroy(h); return rc; };

  const float* payload = nullptr;
  size_t payload_floats = 0;
  if (int rc = parse_blob(e, weights, nbytes, &payload, &payload_floats)) return bail(rc);
  e.weights_floats = payload_floats;
  if (cudaMalloc(&e.weights_dev, payload_floats * sizeof(float)) != cudaSuccess) return bail(fail(DPDF_ERR_NOMEM, "cudaMalloc(weights) failed"));
  if (cudaMemcpy(e.weights_dev, payload, payload_floats * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    return bail(fail(DPDF_ERR_CUDA, "weight upload failed"));
  if (int rc = bind_weights(e)) return bail(rc);
  {
    auto it = e.wtable.find("df_dec.df_convp.w");
    if (it != e.wtable.end()) e.dfp_w_host.assign(payload + it->second.first, payload + it->second.first + it->second.second);
  }

  // band tables
  {
    std::vector<int> aux(33 + d.F);
    std::vector<float> invw(32);
    int k = 0;
    for (int b = 0; b < 32; ++b) {
      aux[b] = k;
      for (int i = 0; i < spec->erb_widths[b]; ++i) aux[33 + k + i] = b;
      k += spec->erb_widths[b];
      invw[b] = (float)(1.0 / spec->erb_widths[b]);
    }
    aux[32] = k;
    float* invw_dev = nullptr;
    if (cudaMalloc(&e.aux_int, aux.size() * sizeof(int) + 32 * sizeof(float)) != cudaSuccess) return bail(fail(DPDF_ERR_NOMEM, "cudaMalloc(aux) failed"));
    invw_dev = reinterpret_cast<float*>(e.aux_int + aux.size());
    cudaMemcpy(e.aux_int, aux.data(), aux.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(invw_dev, invw.data(), 32 * sizeof(float), cudaMemcpyHostToDevice);
    e.w.band_start = e.aux_int;
    e.w.band_of_bin = e.aux_int + 33;
    e.w.band_inv_w = invw_dev;
  }

  // arena: state + scratch, 256-byte aligned sub-allocations
  {
    const size_t Bm = max_streams;
    std::vector<std::pair<float**, size_t>> items;
    State& s = e.st;
    Scratch& c = e.sc;
    bool in_scratch = false;
    auto add = [&](float*& ptr, size_t per) {
      items.push_back({&ptr, per * Bm});
      if (in_scratch) e.sc_items.push_back({&ptr, per});
    };
    add(s.mu, d.fe_feat); add(s.s, NDF); add(s.erb_ring, 3 * d.fe_feat); add(s.df_ring, 3 * 2 * NDF);
    add(s.inter_erb, (size_t)std::max(d.N, 1) * d.fe[3] * C); add(s.inter_df, (size_t)std::max(d.N, 1) * (NDF / 2) * C);
    add(s.h_enc, H); add(s.h_erb, 2 * H); add(s.h_df, 2 * H);
    add(s.c0_ring, (size_t)ORD * NDF * C); add(s.dfp_acc, (size_t)ORD * NDF * 10); add(s.mask_ring, 3 * d.F * 2); add(s.coef_ring, 3 * NDF * 2 * ORD);
    add(s.dfspec_ring, (size_t)ORD * d.F * 2); add(s.in_hist, d.hop); add(s.ola, d.hop);
    in_scratch = true;
    add(c.e0, (size_t)d.fe[0] * C); add(c.e1, (size_t)d.fe[1] * C); add(c.e2, (size_t)d.fe[2] * C); add(c.e3, (size_t)d.fe[3] * C);
    add(c.c0, NDF * C); add(c.c1, (NDF / 2) * C); add(c.xe, (size_t)d.fe[3] * C);
    add(c.hcat_e, (size_t)d.fe[3] * 2 * C); add(c.hcat_d, (NDF / 2) * 2 * C);
    add(c.emb_e, 512); add(c.cemb, 512); add(c.g0, H); add(c.henc, H); add(c.emb, 512);
    add(c.x1, H); add(c.herb1, H); add(c.herb2, H); add(c.ed, 512); add(c.ed2, (size_t)d.fe[3] * C);
    add(c.x2, H); add(c.hdf1, H); add(c.hdf2, H); add(c.cc, H); add(c.co, NDF * 2 * ORD); add(c.dfp, NDF * 2 * ORD);
    add(c.d3, (size_t)d.fe[2] * C); add(c.d2, (size_t)d.fe[1] * C); add(c.d1, (size_t)d.fe[0] * C); add(c.m, d.fe[0]);
    add(c.spec_tc, (size_t)e.spec_tc_ld); add(c.yspec_tc, (size_t)e.yspec_tc_ld);
    size_t total = 0;
    for (auto& it : items) total += (it.second * sizeof(float) + 255) / 256 * 256;
    total += (Bm * sizeof(int) + 255) / 256 * 256;
    e.arena_bytes = total;
    if (cudaMalloc(&e.arena, total) != cudaSuccess)
      return bail(fail(DPDF_ERR_NOMEM, "cudaMalloc(%zu MB) for %d streams failed", total >> 20, max_streams));
    cudaMemset(e.arena, 0, total);
    char* cur = static_cast<char*>(e.arena);
    for (auto& it : items) {
      *it.first = reinterpret_cast<float*>(cur);
      cur += (it.second * sizeof(float) + 255) / 256 * 256;
    }
    s.pos = reinterpret_cast<int*>(cur);
  }
  if (cudaMalloc(&e.io_lanes, Engine::MAX_LANES * sizeof(IoDesc)) != cudaSuccess || cudaMalloc(&e.slots_dev, max_streams * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&e.flags_dev, max_streams * sizeof(int)) != cudaSuccess)
    return bail(fail(DPDF_ERR_NOMEM, "cudaMalloc(io) failed"));
  e.io_dev = e.io_lanes;
  e.progress_tiles = (max_streams + 31) / 32 + 1;            // sweep tiles are 128 / D streams, D <= 4
  if (cudaMalloc(&e.progress_dev, (size_t)Engine::MAX_LANES * 4 * e.progress_tiles * sizeof(int)) != cudaSuccess)
    return bail(fail(DPDF_ERR_NOMEM, "cudaMalloc(progress) failed"));
  cudaMemset(e.progress_dev, 0, (size_t)Engine::MAX_LANES * 4 * e.progress_tiles * sizeof(int));
  if (cudaMalloc(&e.post_ctr_dev, Engine::MAX_LANES * 4 * sizeof(int)) != cudaSuccess) return bail(fail(DPDF_ERR_NOMEM, "cudaMalloc(post counters) failed"));
  cudaMemset(e.post_ctr_dev, 0, Engine::MAX_LANES * 4 * sizeof(int));
  if (cudaHostAlloc(&e.err_host, 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(&e.err_dev, e.err_host, 0) != cudaSuccess)
    return bail(fail(DPDF_ERR_NOMEM, "cudaHostAlloc(error words) failed"));
  memset(e.err_host, 0, 4 * sizeof(int));
  if (cudaStreamCreateWithFlags(&e.own_stream, cudaStreamDefault) != cudaSuccess) return bail(fail(DPDF_ERR_CUDA, "stream creation failed"));
  for (int l = 0; l < Engine::MAX_LANES; ++l)
    if (cudaStreamCreateWithFlags(&e.lane_stream[l], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e.br_stream[l], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.br_fork[l], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.br_join[l], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.enc_fork[l], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.enc_join[l], cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e.dfp_stream[l], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.dfp_fork[l], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.dfp_join[l], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e.lane_done[l], cudaEventDisableTiming) != cudaSuccess)
      return bail(fail(DPDF_ERR_CUDA, "lane stream creation failed"));
  if (cudaEventCreateWithFlags(&e.lane_fork, cudaEventDisableTiming) != cudaSuccess) return bail(fail(DPDF_ERR_CUDA, "event creation failed"));
  init_frontend_kernels();
  init_conv_kernels();
  init_dprnn_kernels();
  init_dense_kernels();
  init_dprnn_tc_kernels();
  init_dprnn_intra_tc_kernels();
  init_dft_tc_kernels();
  init_conv_tc_kernels();
  init_conv_tma_kernels();
  init_gru_tc_kernels();
  launch_reset(e, nullptr, max_streams, e.own_stream);
  if (cudaStreamSynchronize(e.own_stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
    return bail(fail(DPDF_ERR_CUDA, "engine initialisation kernels failed: %s", cudaGetErrorString(cudaGetLastError())));
  *out = h;
  return 0;
}

extern "C" int dpdf_destroy(dpdf_engine* h) {
  if (!h) return 0;
  Engine& e = h->e;
  cudaSetDevice(e.device);
  cudaDeviceSynchronize();
  for (auto& g : e.graphs) cudaGraphExecDestroy(g.second);
  for (auto ev : e.tev) cudaEventDestroy(ev);
  for (auto& g : e.lane_graphs) cudaGraphExecDestroy(g.second);
  for (int l = 0; l < Engine::MAX_LANES; ++l) {
    if (e.lane_stream[l]) cudaStreamDestroy(e.lane_stream[l]);
    if (e.br_stream[l]) cudaStreamDestroy(e.br_stream[l]);
    if (e.br_fork[l]) cudaEventDestroy(e.br_fork[l]);
    if (e.br_join[l]) cudaEventDestroy(e.br_join[l]);
    if (e.enc_fork[l]) cudaEventDestroy(e.enc_fork[l]);
    if (e.enc_join[l]) cudaEventDestroy(e.enc_join[l]);
    if (e.dfp_stream[l]) cudaStreamDestroy(e.dfp_stream[l]);
    if (e.dfp_fork[l]) cudaEventDestroy(e.dfp_fork[l]);
    if (e.dfp_join[l]) cudaEventDestroy(e.dfp_join[l]);
    if (e.lane_done[l]) cudaEventDestroy(e.lane_done[l]);
  }
  if (e.lane_fork) cudaEventDestroy(e.lane_fork);
  cudaFree(e.weights_dev); cudaFree(e.arena); cudaFree(e.aux_int); cudaFree(e.io_lanes); cudaFree(e.progress_dev); cudaFree(e.post_ctr_dev);
  cudaFree(e.slots_dev); cudaFree(e.flags_dev); cudaFree(e.stage_in); cudaFree(e.stage_out);
  for (int k = 0; k < 2; ++k) {
    cudaFree(e.pipe.in[k]); cudaFree(e.pipe.out[k]); cudaFree(e.pipe.slots[k]); cudaFree(e.pipe.flags[k]);
    if (e.pipe.h2d_done[k]) cudaEventDestroy(e.pipe.h2d_done[k]);
    if (e.pipe.step_done[k]) cudaEventDestroy(e.pipe.step_done[k]);
    if (e.pipe.d2h_done[k]) cudaEventDestroy(e.pipe.d2h_done[k]);
  }
  if (e.pipe.in_stream) cudaStreamDestroy(e.pipe.in_stream);
  if (e.pipe.out_stream) cudaStreamDestroy(e.pipe.out_stream);
  if (e.pinned) cudaFreeHost(e.pinned);
  if (e.err_host) cudaFreeHost(e.err_host);
  if (e.own_stream) cudaStreamDestroy(e.own_stream);
  delete h;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// one hop: kernel schedule
// ---------------------------------------------------------------------------------------------
namespace dpdf {

#define RUN(name, call)                                           \
  do {                                                            \
    if (e.stop_after > 0 && e.run_idx++ >= e.stop_after) {        \
    } else if (e.timing) {                                        \
      cudaEvent_t ev0, ev1;                                       \
      cudaEventCreate(&ev0); cudaEventCreate(&ev1);               \
      cudaEventRecord(ev0, st);                                   \
      call;                                                       \
      cudaEventRecord(ev1, st);                                   \
      e.tev.push_back(ev0); e.tev.push_back(ev1);                 \
      e.tnames.push_back(name);                                   \
    } else {                                                      \
      call;                                                       \
    }                                                             \
  } while (0)

void enqueue_step(Engine& e, int B, cudaStream_t st) {
  const Dims& d = e.d;
  const Weights& w = e.w;
  Scratch& c = e.sc;
  int n = 0;
  e.run_idx = 0;
  const bool seg_pdl = e.pdl == 2 && !e.timing;            // PDL chains everywhere but across the DPRNN stack
  e.pdl_now = e.pdl && !e.timing;
  e.pdl_first = true;                                      // the analysis kernel follows a copy / an event, not a kernel
  const bool use_sep_tc = e.sep_tc == 1 || (e.sep_tc == 2 && std::max(B, e.total_B) >= e.sep_tc_min);
  auto sepconv = [&](Engine& en, const SepProblem* probs, int nprob, int Bn, cudaStream_t s_) {
    if (use_sep_tc && en.sep_tma && sepconv_tma_available()) launch_sepconv_tma(en, probs, nprob, Bn, s_);
    else if (use_sep_tc) launch_sepconv_tc(en, probs, nprob, Bn, s_);
    else launch_sepconv(en, probs, nprob, Bn, s_);
  };
  const bool use_gru_tc = e.gru_tc == 1 || (e.gru_tc == 2 && std::max(B, e.total_B) >= e.gru_tc_min);
  auto gru_cells = [&](Engine& en, const GRUProblem* probs, int nprob, int Bn, cudaStream_t s_) {
    if (use_gru_tc) launch_gru_tc(en, probs, nprob, Bn, s_);
    else launch_gru(en, probs, nprob, Bn, s_);
  };
  if (dft_on_tc(e, B)) { RUN("dft_tc", launch_dft_tc(e, B, st)); ++n; }
  RUN("analysis", launch_analysis(e, B, st)); ++n;
  e.pdl_first = false;
  auto sepp = [&](const SepW& sw, const float* in1, const float* in2, int pidx, float* out, int Fin, int Fout, int stride, int up) {
    SepProblem q{};
    q.mode = 0; q.in1 = in1; q.in2 = in2;
    q.pa = in2 ? w.convp_a[pidx] : nullptr; q.pb = in2 ? w.convp_b[pidx] : nullptr;
    q.dw = sw.dw; q.pw = sw.pw; q.tc_pw = sw.tc_pw; q.bias = sw.b; q.out = out; q.Fin = Fin; q.Fout = Fout; q.stride = stride; q.up = up;
    return q;
  };
  // latency batch sizes only: at throughput sizes the forked launch just competes for HBM (6.43 -> 6.45 ms at 16 384 streams)
  const bool dfp_early = e.dfp_early && e.encoder_fork && !e.dfp_ps && std::max(B, e.total_B) < e.overlap_max;
  SepProblem dfc0{};
  dfc0.mode = 1; dfc0.dw = w.df_conv0_w; dfc0.pw = w.df_conv0_pw; dfc0.tc_pw = w.df_conv0_tc_pw; dfc0.bias = w.df_conv0_b;
  dfc0.Fin = NDF; dfc0.Fout = NDF; dfc0.stride = 1; dfc0.up = 1; dfc0.out = c.c0;
  if (e.encoder_fork) {
    // The two encoder branches share nothing between the analysis kernel and the DPRNN stack: the df chain (df_conv0 ->
    // df_conv1, the bigger problems) runs on a forked stream beside erb_conv0 -> erb_conv1..3.  In the real chain of a
    // 1024-stream hop the three merged launches added 41 + 22 + 10 us after erb_conv0 (profiles/r2L_chain_B1024.txt),
    // of which the erb problems need a fraction each.
    cudaStream_t sd = st;
    const bool fork = !e.timing;
    if (fork) {
      cudaEventRecord(e.enc_fork[e.cur_lane], st);
      sd = e.br_stream[e.cur_lane];
      cudaStreamWaitEvent(sd, e.enc_fork[e.cur_lane], 0);
      e.pdl_first = true;                                    // first kernel of the forked chain: its predecessor is an event
    }
    RUN("sepconv", sepconv(e, &dfc0, 1, B, sd)); ++n;
    if (dfp_early) {
      // The pathway term of the deep-filter coefficients needs nothing but the c0 ring df_conv0 has just pushed: it runs
      // on a stream of its own beside the DPRNN stack (most SMs idle at latency batch sizes) instead of on the decoder
      // tail, where it was 40 of the coefficient tail's 50 us at 1024 streams (profiles/r3f_chain_B1024.txt).
      cudaStream_t sp = st;
      if (fork) {
        cudaEventRecord(e.dfp_fork[e.cur_lane], sd);
        sp = e.dfp_stream[e.cur_lane];
        cudaStreamWaitEvent(sp, e.dfp_fork[e.cur_lane], 0);
        e.pdl_first = true;
      } else {
        sp = sd;
      }
      RUN("df_pathway", launch_df_pathway_early(e, B, sp)); ++n;
      if (fork) cudaEventRecord(e.dfp_join[e.cur_lane], sp);
    }
    e.pdl_first = false;
    SepProblem q = sepp(w.df_conv1, c.c0, nullptr, 0, c.c1, NDF, NDF / 2, 2, 1);
    RUN("sepconv", sepconv(e, &q, 1, B, sd)); ++n;
    if (fork) cudaEventRecord(e.enc_join[e.cur_lane], sd);
    RUN("erb_conv0", launch_erb_conv0(e, B, st)); ++n;
    for (int i = 0; i < 3; ++i) {
      const float* in = i == 0 ? c.e0 : (i == 1 ? c.e1 : c.e2);
      float* out = i == 0 ? c.e1 : (i == 1 ? c.e2 : c.e3);
      q = sepp(w.erb_conv[i], in, nullptr, 0, out, d.fe[i], d.fe[i + 1], d.stride[i], 1);
      RUN("sepconv", sepconv(e, &q, 1, B, st)); ++n;
    }
    if (fork) cudaStreamWaitEvent(st, e.enc_join[e.cur_lane], 0);
  } else {
    RUN("erb_conv0", launch_erb_conv0(e, B, st)); ++n;
    SepProblem pr[2];
    pr[0] = dfc0;
    pr[1] = sepp(w.erb_conv[0], c.e0, nullptr, 0, c.e1, d.fe[0], d.fe[1], d.stride[0], 1);
    RUN("sepconv", sepconv(e, pr, 2, B, st)); ++n;
    pr[0] = sepp(w.df_conv1, c.c0, nullptr, 0, c.c1, NDF, NDF / 2, 2, 1);
    pr[1] = sepp(w.erb_conv[1], c.e1, nullptr, 0, c.e2, d.fe[1], d.fe[2], d.stride[1], 1);
    RUN("sepconv", sepconv(e, pr, 2, B, st)); ++n;
    pr[0] = sepp(w.erb_conv[2], c.e2, nullptr, 0, c.e3, d.fe[2], d.fe[3], d.stride[2], 1);
    RUN("sepconv", sepconv(e, pr, 1, B, st)); ++n;
  }
  const bool intra_on_tc = e.intra_tc == 1 || (e.intra_tc == 2 && std::max(B, e.total_B) >= e.intra_tc_min);
  e.overlap_now = e.overlap_now && intra_on_tc && !e.timing;        // requested by the caller (enqueue_lanes / run_hops_free)
  if (seg_pdl) e.pdl_now = false;
  for (int i = 0; i < d.N; ++i) {
    if (intra_on_tc) {
      // Sweep of block i >= 1 as a programmatic dependent of the post kernel of block i - 1 (Engine::intra_pdl): its CTAs
      // take the SMs the post kernel's tail leaves idle, set up barriers and tensor memory and pull their 96 KB of weight
      // images, then wait (griddepcontrol.wait) for the post grid to complete and flush.  They cannot crowd the post
      // kernel: dependents launch only once every post CTA has started.
      const bool chain = e.intra_pdl && i > 0 && e.post_tc && !e.timing && !e.pdl_now;
      if (chain) { e.pdl_now = true; e.pdl_first = false; }
      RUN("dprnn_intra", launch_dprnn_intra_tc(e, i, B, st));
      if (chain) e.pdl_now = false;
    } else { RUN("dprnn_intra", launch_dprnn_intra(e, i, B, st)); }
    ++n;
    if (e.post_tc) { RUN("dprnn_post", launch_dprnn_post_tc(e, i, B, st)); }
    else { RUN("dprnn_post", launch_dprnn_post(e, i, B, st)); }
    ++n;
  }
  const float* xe_final = d.N > 0 ? c.xe : c.e3;
  const bool pdl_saved = e.pdl_now;
  const bool tail_chain = (e.tail_pdl || seg_pdl) && !e.timing && !e.pdl_now;
  if (tail_chain) { e.pdl_now = true; e.pdl_first = true; }    // dense tail as a PDL chain (Engine::tail_pdl); its first kernel is a plain
                                                               // launch: it must see the whole DPRNN stack (overlapped post kernels included)
  auto glp = [&](const GLW& gw, const float* in0, int ld0, float* out, int ldo, int act) {
    GLProblem q{};
    q.in0 = in0; q.ld0 = ld0; q.in1 = nullptr; q.ld1 = 0; q.split = 1 << 30; q.w = gw; q.out = out; q.ldo = ldo; q.col0 = 0;
    q.addend = nullptr; q.lda = 0; q.act = act;
    return q;
  };
  {
    GLProblem pr[2];
    int np = 0;
    pr[np++] = glp(w.df_fc_emb, c.c1, (NDF / 2) * C, c.cemb, 512, 1);
    if (d.hr48) pr[np++] = glp(w.erb_fc_emb, xe_final, d.fe[3] * C, c.emb_e, 512, 1);
    RUN("gl", launch_gl(e, pr, np, B, st)); ++n;
    if (tail_chain) e.pdl_first = false;
  }
  {
    GLProblem q = glp(w.enc_in, d.hr48 ? c.emb_e : xe_final, 512, c.g0, H, 1);
    q.in1 = c.cemb; q.ld1 = 512; q.split = 512;
    RUN("gl", launch_gl(e, &q, 1, B, st)); ++n;
  }
  {
    GRUProblem g{c.g0, e.st.h_enc, H, w.enc_gru, c.henc};
    RUN("gru", gru_cells(e, &g, 1, B, st)); ++n;
  }
  {
    GLProblem q = glp(w.enc_out, c.henc, H, c.emb, 512, 1);
    RUN("gl", launch_gl(e, &q, 1, B, st)); ++n;
  }
  {
    GLProblem pr[2] = {glp(w.erbdec_in, c.emb, 512, c.x1, H, 1), glp(w.dfdec_in, c.emb, 512, c.x2, H, 1)};
    RUN("gl", launch_gl(e, pr, 2, B, st)); ++n;
  }
  {
    GRUProblem g[2] = {{c.x1, e.st.h_erb, 2 * H, w.erb_gru[0], c.herb1}, {c.x2, e.st.h_df, 2 * H, w.df_gru[0], c.hdf1}};
    RUN("gru", gru_cells(e, g, 2, B, st)); ++n;
    GRUProblem g2[2] = {{c.herb1, e.st.h_erb + H, 2 * H, w.erb_gru[1], c.herb2}, {c.hdf1, e.st.h_df + H, 2 * H, w.df_gru[1], c.hdf2}};
    RUN("gru", gru_cells(e, g2, 2, B, st)); ++n;
  }
  {
    GLProblem pr[2] = {glp(w.erbdec_out, c.herb2, H, c.ed, 512, 1), glp(w.df_skip, c.emb, 512, c.cc, H, 0)};
    pr[1].addend = c.hdf2; pr[1].lda = H;
    RUN("gl", launch_gl(e, pr, 2, B, st)); ++n;
  }
  e.pdl_now = pdl_saved || seg_pdl;
  // The two decoder tails are independent until the synthesis kernel: the deep-filter coefficients (df_out linear +
  // pathway conv) run on a forked stream beside the ERB decoder's transposed-conv stack (graph capture turns the events
  // into dependencies); sequential while timing with events.
  cudaStream_t sb = st;
  const bool fork = e.decoder_fork && !e.timing;
  if (fork) {
    cudaEventRecord(e.br_fork[e.cur_lane], st);
    sb = e.br_stream[e.cur_lane];
    cudaStreamWaitEvent(sb, e.br_fork[e.cur_lane], 0);
  }
  GRUProblem all[5] = {{c.g0, e.st.h_enc, H, w.enc_gru, c.henc},
                       {c.x1, e.st.h_erb, 2 * H, w.erb_gru[0], c.herb1}, {c.x2, e.st.h_df, 2 * H, w.df_gru[0], c.hdf1},
                       {c.herb1, e.st.h_erb + H, 2 * H, w.erb_gru[1], c.herb2}, {c.hdf1, e.st.h_df + H, 2 * H, w.df_gru[1], c.hdf2}};
  {
    // The new GRU states are committed to the slot arena beside the decoder tails: nothing reads them before the next hop.
    // In the real chain of a 1024-stream hop (tools/chain_profile.py, profiles/r2L_chain_B1024.txt) the coefficient tail
    // - df_out linear 9 us + pathway conv 44 us - is the LONGER of the two (the three transposed convs + conv0_out add
    // 34 us), so the commit (7 us) rides on the main chain.
    // ... unless the pathway conv has left it (dfp_early): then the coefficient tail is a linear and an add, and takes the commit
    if (!(dfp_early && fork)) { RUN("gru_commit", launch_gru_commit(e, all, 5, B, st)); ++n; }
    if (fork) e.pdl_first = true;                            // first kernel of the forked chain: its predecessor is an event, not a kernel
    GLProblem q = glp(w.df_out, c.cc, H, c.co, NDF * 2 * ORD, 2);
    RUN("gl", launch_gl(e, &q, 1, B, sb)); ++n;
    e.pdl_first = false;
  }
  if (dfp_early) {
    if (!e.timing) cudaStreamWaitEvent(sb, e.dfp_join[e.cur_lane], 0);
    RUN("df_combine", launch_df_combine(e, B, sb));
    if (fork) { RUN("gru_commit", launch_gru_commit(e, all, 5, B, sb)); ++n; }
  } else if (e.dfp_ps) { RUN("df_pathway", launch_df_pathway_ps(e, B, sb)); }
  else { RUN("df_pathway", launch_df_pathway(e, B, sb)); }
  ++n;
  if (fork) cudaEventRecord(e.br_join[e.cur_lane], sb);
  if (d.hr48) {
    GLProblem q = glp(w.erbdec_fc, c.ed, 512, c.ed2, d.fe[3] * C, 1);
    RUN("gl", launch_gl(e, &q, 1, B, st)); ++n;
  }
  {
    const float* edv = d.hr48 ? c.ed2 : c.ed;
    SepProblem q = sepp(w.convt[0], edv, c.e3, 0, c.d3, d.fe[3], d.fe[3] * d.up[0], 1, d.up[0]);
    RUN("sepconv", sepconv(e, &q, 1, B, st)); ++n;
    q = sepp(w.convt[1], c.d3, c.e2, 1, c.d2, d.fe[2], d.fe[2] * d.up[1], 1, d.up[1]);
    RUN("sepconv", sepconv(e, &q, 1, B, st)); ++n;
    q = sepp(w.convt[2], c.d2, c.e1, 2, c.d1, d.fe[1], d.fe[1] * d.up[2], 1, d.up[2]);
    RUN("sepconv", sepconv(e, &q, 1, B, st)); ++n;
  }
  RUN("conv0_out", launch_conv0_out(e, B, st)); ++n;
  if (fork) cudaStreamWaitEvent(st, e.br_join[e.cur_lane], 0);
  RUN("synthesis", launch_synthesis(e, B, st)); ++n;
  if (dft_on_tc(e, B)) { RUN("dft_tc", launch_idft_tc(e, B, st)); ++n; }
  e.pdl_now = false;
  e.launches = n;
}

}  // namespace dpdf

// ---------------------------------------------------------------------------------------------
// step entry points
// ---------------------------------------------------------------------------------------------
static int check_batch(Engine& e, int B) {
  if (B <= 0) return fail(DPDF_ERR_INVALID, "B must be positive");
  if (B > e.max_streams) return fail(DPDF_ERR_INVALID, "B=%d exceeds max_streams=%d", B, e.max_streams);
  return 0;
}

// Errors raised by kernels of earlier hops (Engine::err_host, mapped memory the kernels store to): the overlapped post
// kernel gave up waiting for its sweep, or an activation left the FP16 operand range.  Either way the recurrent state
// of the affected streams is no longer the reference's; the words are cleared once reported.  Called after the stream
// synchronisation of the *_host entry points (errors of THIS hop) and at the start of every device-pointer entry
// point (errors of hops enqueued earlier and finished since).
static int check_device_errors(Engine& e) {
  volatile int* w = e.err_host;
  if (!w) return 0;
  const int overlap = w[DPDF_ERRW_OVERLAP], range = w[DPDF_ERRW_RANGE];
  if (!overlap && !range) return 0;
  w[DPDF_ERRW_OVERLAP] = 0;
  w[DPDF_ERRW_RANGE] = 0;
  if (overlap)
    return fail(DPDF_ERR_CUDA, "a DPRNN post tile timed out waiting for the intra-frame sweep (GPU preempted or shared?): the "
                               "hop is invalid, reset the streams of that batch; set option overlap=0 to use full grid dependencies");
  return fail(DPDF_ERR_CUDA, "an activation exceeded the FP16 operand range (|x| > 65504) or was non-finite in a tensor-core "
                             "converter: outputs of that hop are invalid; reset the affected streams (option *_tc=0 selects the FP32 kernels)");
}

extern "C" int dpdf_poll_error(dpdf_engine* h, int32_t synchronize) {
  if (!h) return fail(DPDF_ERR_INVALID, "NULL engine");
  Engine& e = h->e;
  if (synchronize) {
    CU(cudaSetDevice(e.device));
    CU(cudaDeviceSynchronize());
  }
  return check_device_errors(e);
}

static void drop_graphs(Engine& e) {
  for (auto& g : e.graphs) cudaGraphExecDestroy(g.second);
  for (auto& g : e.lane_graphs) cudaGraphExecDestroy(g.second);
  e.graphs.clear();
  e.lane_graphs.clear();
}

// Number of lanes of a step over B streams: explicit option, else enough streams per lane to keep the tensor-core
// kernels' 128-stream tiles full.
// The overlapped post kernel and lanes exclude each other: post CTAs waiting for their sweep would
// hold the SMs another lane's sweep needs.  Measured (profiles/r01O_lanes.log): one chain + overlap wins below 4096
// streams, eight free-running lanes without overlap above.
// Round 2: with the fragment-form sweep (k_dprnn_intra_tc.cu:intra_sweep_f, 32-stream CTAs, 70 us per launch) the
// overlapped post kernel cannot keep up with its sweep any more and its spinning tiles only hold SMs: 128-stream lanes
// without overlap win wherever that form runs (profiles/r3p_*, r3q_*, r3r_*: 1024 streams 0.743 -> 0.653 ms per hop, lock-step
// p50 0.748 -> 0.717 ms; Engine::overlap = 2 forces the overlap there too).
static bool overlap_applies(const Engine& e, int B) {
  const bool tc = e.post_tc && (e.intra_tc == 1 || (e.intra_tc == 2 && B >= e.intra_tc_min));
  if (!e.overlap || !tc || B >= e.overlap_max || e.lanes > 1) return false;
  if (e.overlap == 1 && e.lanes != 1 && e.intra_frag && intra_tc_dup(e, B) == 4) return false;
  return true;
}
static int lanes_for(const Engine& e, int B) {
  if (overlap_applies(e, B)) return 1;
  // measured: profiles/r01B_lanes.log, r3q_sweep_overlap_1024.log, r3v_sweep_frag_lanes_big.log
  // r4b_sweep_intra_tc_tiny.log (256 / 384 / 512 / 640 streams: one lane 0.58 / 0.60 / 0.65 / 0.68 ms per hop, 128-stream lanes 0.51 / 0.54 / 0.56 / 0.59)
  // r4O / r4P / r4Q: beyond the fragment form's range two lanes beat eight (7168 / 8192 / 16384 / 24576 streams 3.03 / 3.36 / 6.44 /
  // 9.62 -> 2.95 / 3.25 / 6.27 / 9.35 ms per hop)
  const bool frag = e.intra_frag && intra_tc_dup(e, B) == 4;
  int L = e.lanes > 0 ? e.lanes : (B < 256 ? 1 : (frag ? (B >= 2560 ? 4 : 8) : 2));
  L = std::min(L, Engine::MAX_LANES);
  while (L > 1 && B / L < e.lane_min) --L;
  return std::max(L, 1);
}
static void lane_range(const Engine& e, int B, int L, int l, int* r0, int* n) {
  const int g = e.lane_min;
  const int per = ((B + L - 1) / L + g - 1) / g * g;              // lane sizes are multiples of the 128-stream MMA tile (Engine::lane_min)
  *r0 = std::min(B, l * per);
  *n = std::min(B, (l + 1) * per) - *r0;
}

// Enqueue one hop for B streams: every lane's kernel chain with its row range of the scratch arena and its own
// I/O descriptor; lanes > 0 run on forked streams (graph capture turns the events into graph dependencies).
static void enqueue_lanes(Engine& e, int B, cudaStream_t st, bool fork) {
  const int L = lanes_for(e, B);
  std::vector<float*> base(e.sc_items.size());
  for (size_t i = 0; i < base.size(); ++i) base[i] = *e.sc_items[i].first;
  e.total_B = B;
  e.overlap_now = L == 1 && overlap_applies(e, B);
  int launches = 0;
  if (fork && L > 1) cudaEventRecord(e.lane_fork, st);
  for (int l = 0; l < L; ++l) {
    int r0, n;
    lane_range(e, B, L, l, &r0, &n);
    if (n <= 0) continue;
    for (size_t i = 0; i < base.size(); ++i) *e.sc_items[i].first = base[i] + (size_t)r0 * e.sc_items[i].second;
    e.io_dev = e.io_lanes + l;
    e.cur_lane = l;
    cudaStream_t ls = (fork && l > 0) ? e.lane_stream[l] : st;
    if (fork && l > 0) cudaStreamWaitEvent(ls, e.lane_fork, 0);
    enqueue_step(e, n, ls);
    launches += e.launches;
    if (fork && l > 0) {
      cudaEventRecord(e.lane_done[l], ls);
      cudaStreamWaitEvent(st, e.lane_done[l], 0);
    }
  }
  for (size_t i = 0; i < base.size(); ++i) *e.sc_items[i].first = base[i];
  e.io_dev = e.io_lanes;
  e.total_B = 0;
  e.overlap_now = false;
  e.launches = launches;
}

// Launch the kernels of one hop on `st`, through a cached CUDA graph when enabled.
static int run_step(Engine& e, int B, cudaStream_t st) {
  if (!e.use_graph || e.timing) {
    enqueue_lanes(e, B, st, false);
    CU(cudaGetLastError());
    return 0;
  }
  auto it = e.graphs.find(B);
  if (it == e.graphs.end()) {
    // capture on the engine's own stream: the caller's stream may be the legacy default stream,
    // which cannot be captured; the instantiated graph is then launched on the caller's stream.
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(e.own_stream, cudaStreamCaptureModeRelaxed));
    enqueue_lanes(e, B, e.own_stream, true);
    cudaError_t err = cudaStreamEndCapture(e.own_stream, &graph);
    if (err != cudaSuccess) return fail(DPDF_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(err));
    cudaGraphExec_t exec = nullptr;
    err = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) return fail(DPDF_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(err));
    it = e.graphs.emplace(B, exec).first;
  }
  CU(cudaGraphLaunch(it->second, st));
  return 0;
}

// T hops with free-running lanes: the lanes of a batch share nothing inside a hop and nothing across hops but their
// own slots, so every lane replays its own one-hop graph T times on its own stream and the caller's stream joins them
// once at the end.  The lanes drift out of lock step, which is the point: the latency-bound recurrence kernels of one
// lane overlap with the throughput kernels of the others for the whole run, not just inside one hop.
static int run_hops_free(Engine& e, int B, int T, cudaStream_t st) {
  const int L = lanes_for(e, B);
  std::vector<float*> base(e.sc_items.size());
  for (size_t i = 0; i < base.size(); ++i) base[i] = *e.sc_items[i].first;
  cudaGraphExec_t exec[Engine::MAX_LANES] = {};
  int launches = 0, rc = 0;
  e.total_B = B;
  for (int l = 0; l < L && rc == 0; ++l) {
    int r0, n;
    lane_range(e, B, L, l, &r0, &n);
    if (n <= 0) continue;
    auto it = e.lane_graphs.find(B * Engine::MAX_LANES + l);
    if (it == e.lane_graphs.end()) {
      for (size_t i = 0; i < base.size(); ++i) *e.sc_items[i].first = base[i] + (size_t)r0 * e.sc_items[i].second;
      e.io_dev = e.io_lanes + l;
      e.cur_lane = l;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(e.own_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { rc = fail(DPDF_ERR_CUDA, "graph capture failed"); break; }
      enqueue_step(e, n, e.own_stream);
      cudaError_t err = cudaStreamEndCapture(e.own_stream, &graph);
      cudaGraphExec_t ex = nullptr;
      if (err == cudaSuccess) err = cudaGraphInstantiate(&ex, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (err != cudaSuccess) { rc = fail(DPDF_ERR_CUDA, "lane graph failed: %s", cudaGetErrorString(err)); break; }
      it = e.lane_graphs.emplace(B * Engine::MAX_LANES + l, ex).first;
      e.launches_per_lane = e.launches;
    }
    exec[l] = it->second;
    launches += e.launches_per_lane;
  }
  for (size_t i = 0; i < base.size(); ++i) *e.sc_items[i].first = base[i];
  e.io_dev = e.io_lanes;
  e.total_B = 0;
  if (rc) return rc;
  e.launches = launches;
  CU(cudaEventRecord(e.lane_fork, st));
  for (int l = 0; l < L; ++l)
    if (exec[l]) CU(cudaStreamWaitEvent(e.lane_stream[l], e.lane_fork, 0));
  for (int t = 0; t < T; ++t)
    for (int l = 0; l < L; ++l)
      if (exec[l]) CU(cudaGraphLaunch(exec[l], e.lane_stream[l]));
  for (int l = 0; l < L; ++l)
    if (exec[l]) {
      CU(cudaEventRecord(e.lane_done[l], e.lane_stream[l]));
      CU(cudaStreamWaitEvent(st, e.lane_done[l], 0));
    }
  return 0;
}

static int set_io(Engine& e, const float* in, long long in_stride, float* out, long long out_stride,
                  const int32_t* slot_ids, const int32_t* flags, int mode, int B, cudaStream_t st) {
  IoDesc io[Engine::MAX_LANES] = {};
  const int L = lanes_for(e, B);
  const long long rin = mode == 1 ? (long long)e.d.F * 2 : in_stride, rout = mode == 1 ? (long long)e.d.F * 2 : out_stride;
  for (int l = 0; l < L; ++l) {
    int r0, n;
    lane_range(e, B, L, l, &r0, &n);
    io[l].in = in + (size_t)r0 * rin; io[l].out = out + (size_t)r0 * rout;
    io[l].in_stride = in_stride; io[l].out_stride = out_stride;
    io[l].slot_ids = slot_ids ? slot_ids + r0 : nullptr; io[l].flags = flags ? flags + r0 : nullptr;
    io[l].t_in = 0; io[l].t_out = 0; io[l].mode = mode; io[l].slot_base = r0;
    io[l].err = e.err_dev;
  }
  CU(cudaMemcpyAsync(e.io_lanes, io, sizeof(IoDesc) * L, cudaMemcpyHostToDevice, st));   // pageable source: staged before return
  return 0;
}

extern "C" int dpdf_step_spec(dpdf_engine* h, const float* spec_in, float* spec_out, const int32_t* slot_ids,
                              const int32_t* flags, int32_t B, void* cuda_stream) {
  if (!h || !spec_in || !spec_out) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  CU(cudaSetDevice(e.device));
  if (int rc = check_device_errors(e)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (int rc = set_io(e, spec_in, 0, spec_out, 0, slot_ids, flags, 1, B, st)) return rc;
  e.last_B = B;
  return run_step(e, B, st);
}

extern "C" int dpdf_run_pcm(dpdf_engine* h, const float* pcm_in, int64_t in_stride, float* pcm_out, int64_t out_stride,
                            const int32_t* slot_ids, const int32_t* flags, int32_t B, int32_t T, void* cuda_stream) {
  if (!h || !pcm_in || !pcm_out) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  if (T <= 0) return fail(DPDF_ERR_INVALID, "T must be positive");
  if (in_stride < (int64_t)T * e.d.hop || out_stride < (int64_t)T * e.d.hop)
    return fail(DPDF_ERR_INVALID, "row stride smaller than T*hop");
  CU(cudaSetDevice(e.device));
  if (int rc = check_device_errors(e)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (int rc = set_io(e, pcm_in, in_stride, pcm_out, out_stride, slot_ids, flags, 0, B, st)) return rc;
  e.last_B = B;
  if (T > 1 && e.use_graph && !e.timing && e.free_lanes && lanes_for(e, B) > 1) return run_hops_free(e, B, T, st);
  for (int t = 0; t < T; ++t)
    if (int rc = run_step(e, B, st)) return rc;
  return 0;
}

extern "C" int dpdf_step_pcm(dpdf_engine* h, const float* pcm_in, int64_t in_stride, float* pcm_out, int64_t out_stride,
                             const int32_t* slot_ids, const int32_t* flags, int32_t B, void* cuda_stream) {
  return dpdf_run_pcm(h, pcm_in, in_stride, pcm_out, out_stride, slot_ids, flags, B, 1, cuda_stream);
}

extern "C" int dpdf_prime_pcm(dpdf_engine* h, const float* pcm_in, int64_t in_stride, const int32_t* slot_ids, int32_t B,
                              void* cuda_stream) {
  if (!h || !pcm_in) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  CU(cudaSetDevice(e.device));
  launch_prime(e, pcm_in, in_stride, slot_ids, B, static_cast<cudaStream_t>(cuda_stream));
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dpdf_reset(dpdf_engine* h, const int32_t* slots_host, int32_t n, void* cuda_stream) {
  if (!h) return fail(DPDF_ERR_INVALID, "NULL engine");
  Engine& e = h->e;
  CU(cudaSetDevice(e.device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (!slots_host || n <= 0) {
    launch_reset(e, nullptr, e.max_streams, st);
  } else {
    if (n > e.max_streams) return fail(DPDF_ERR_INVALID, "n=%d exceeds max_streams", n);
    for (int i = 0; i < n; ++i)
      if (slots_host[i] < 0 || slots_host[i] >= e.max_streams) return fail(DPDF_ERR_INVALID, "slot %d out of range", slots_host[i]);
    CU(cudaMemcpyAsync(e.slots_dev, slots_host, n * sizeof(int), cudaMemcpyHostToDevice, st));
    launch_reset(e, e.slots_dev, n, st);
  }
  CU(cudaGetLastError());
  return 0;
}

// ---- host-buffer variants ---------------------------------------------------------------------
static int ensure_stage(Engine& e, size_t in_floats, size_t out_floats) {
  if (in_floats > e.stage_in_floats) {
    cudaFree(e.stage_in);
    e.stage_in = nullptr;
    if (cudaMalloc(&e.stage_in, in_floats * sizeof(float)) != cudaSuccess) return fail(DPDF_ERR_NOMEM, "staging alloc failed");
    e.stage_in_floats = in_floats;
  }
  if (out_floats > e.stage_out_floats) {
    cudaFree(e.stage_out);
    e.stage_out = nullptr;
    if (cudaMalloc(&e.stage_out, out_floats * sizeof(float)) != cudaSuccess) return fail(DPDF_ERR_NOMEM, "staging alloc failed");
    e.stage_out_floats = out_floats;
  }
  return 0;
}

static int host_ids(Engine& e, const int32_t* slot_ids, const int32_t* flags, int B, const int32_t** s_dev,
                    const int32_t** f_dev, cudaStream_t st) {
  *s_dev = nullptr;
  *f_dev = nullptr;
  if (slot_ids) {
    for (int i = 0; i < B; ++i)
      if (slot_ids[i] < 0 || slot_ids[i] >= e.max_streams) return fail(DPDF_ERR_INVALID, "slot %d out of range", slot_ids[i]);
    CU(cudaMemcpyAsync(e.slots_dev, slot_ids, B * sizeof(int), cudaMemcpyHostToDevice, st));
    *s_dev = e.slots_dev;
  }
  if (flags) {
    CU(cudaMemcpyAsync(e.flags_dev, flags, B * sizeof(int), cudaMemcpyHostToDevice, st));
    *f_dev = e.flags_dev;
  }
  return 0;
}

extern "C" int dpdf_step_spec_host(dpdf_engine* h, const float* spec_in, float* spec_out, const int32_t* slot_ids,
                                   const int32_t* flags, int32_t B) {
  if (!h || !spec_in || !spec_out) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  CU(cudaSetDevice(e.device));
  const size_t n = (size_t)B * e.d.F * 2;
  if (int rc = ensure_stage(e, n, n)) return rc;
  cudaStream_t st = e.own_stream;
  const int32_t *s_dev, *f_dev;
  if (int rc = host_ids(e, slot_ids, flags, B, &s_dev, &f_dev, st)) return rc;
  CU(cudaMemcpyAsync(e.stage_in, spec_in, n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = dpdf_step_spec(h, e.stage_in, e.stage_out, s_dev, f_dev, B, st)) return rc;
  CU(cudaMemcpyAsync(spec_out, e.stage_out, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return check_device_errors(e);
}

extern "C" int dpdf_run_pcm_host(dpdf_engine* h, const float* pcm_in, float* pcm_out, const int32_t* slot_ids, int32_t B,
                                 int32_t T) {
  if (!h || !pcm_in || !pcm_out) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  if (T <= 0) return fail(DPDF_ERR_INVALID, "T must be positive");
  CU(cudaSetDevice(e.device));
  const size_t n = (size_t)B * T * e.d.hop;
  if (int rc = ensure_stage(e, n, n)) return rc;
  cudaStream_t st = e.own_stream;
  const int32_t *s_dev, *f_dev;
  if (int rc = host_ids(e, slot_ids, nullptr, B, &s_dev, &f_dev, st)) return rc;
  CU(cudaMemcpyAsync(e.stage_in, pcm_in, n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = dpdf_run_pcm(h, e.stage_in, (int64_t)T * e.d.hop, e.stage_out, (int64_t)T * e.d.hop, s_dev, nullptr, B, T, st)) return rc;
  CU(cudaMemcpyAsync(pcm_out, e.stage_out, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return check_device_errors(e);
}

extern "C" int dpdf_step_pcm_host(dpdf_engine* h, const float* pcm_in, float* pcm_out, const int32_t* slot_ids,
                                  const int32_t* flags, int32_t B) {
  if (!h || !pcm_in || !pcm_out) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  CU(cudaSetDevice(e.device));
  const size_t n = (size_t)B * e.d.hop;
  if (int rc = ensure_stage(e, n, n)) return rc;
  cudaStream_t st = e.own_stream;
  const int32_t *s_dev, *f_dev;
  if (int rc = host_ids(e, slot_ids, flags, B, &s_dev, &f_dev, st)) return rc;
  CU(cudaMemcpyAsync(e.stage_in, pcm_in, n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = dpdf_run_pcm(h, e.stage_in, e.d.hop, e.stage_out, e.d.hop, s_dev, f_dev, B, 1, st)) return rc;
  CU(cudaMemcpyAsync(pcm_out, e.stage_out, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return check_device_errors(e);
}

// ---- pipelined host entry ----------------------------------------------------------------------------------------------
// dpdf_step_pcm_host is synchronous: copy in, hop, copy out, wait - what a caller with one hop of audio in hand does.  A
// server that always has the next hop ready submits it before collecting the previous one: ticket t's host->device copy
// (in_stream) and ticket t-1's device->host copy (out_stream) then overlap ticket t-1's / t's kernels (own_stream), and
// the host never idles in a stream synchronisation between hops.  Two staging sets, so at most two tickets are in flight;
// submit blocks until ticket t-2 has been delivered.  Host buffers must stay valid (and should be pinned) until dpdf_wait.
static int ensure_pipe(Engine& e, size_t floats, int B) {
  Engine::HostPipe& hp = e.pipe;
  if (!hp.in_stream) {
    if (cudaStreamCreateWithFlags(&hp.in_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&hp.out_stream, cudaStreamNonBlocking) != cudaSuccess)
      return fail(DPDF_ERR_CUDA, "pipeline stream creation failed");
    for (int k = 0; k < 2; ++k)
      if (cudaEventCreateWithFlags(&hp.h2d_done[k], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&hp.step_done[k], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&hp.d2h_done[k], cudaEventDisableTiming) != cudaSuccess)
        return fail(DPDF_ERR_CUDA, "pipeline event creation failed");
  }
  if (floats > hp.cap || B > hp.cap_ids) {
    CU(cudaDeviceSynchronize());
    for (int k = 0; k < 2; ++k) {
      cudaFree(hp.in[k]); cudaFree(hp.out[k]); cudaFree(hp.slots[k]); cudaFree(hp.flags[k]);
      hp.in[k] = hp.out[k] = nullptr; hp.slots[k] = hp.flags[k] = nullptr;
      if (cudaMalloc(&hp.in[k], floats * sizeof(float)) != cudaSuccess || cudaMalloc(&hp.out[k], floats * sizeof(float)) != cudaSuccess ||
          cudaMalloc(&hp.slots[k], B * sizeof(int)) != cudaSuccess || cudaMalloc(&hp.flags[k], B * sizeof(int)) != cudaSuccess)
        return fail(DPDF_ERR_NOMEM, "pipeline staging alloc failed");
    }
    hp.cap = floats;
    hp.cap_ids = B;
  }
  return 0;
}

extern "C" int dpdf_submit_pcm_host(dpdf_engine* h, const float* pcm_in, float* pcm_out, const int32_t* slot_ids,
                                    const int32_t* flags, int32_t B, int64_t* ticket) {
  if (!h || !pcm_in || !pcm_out || !ticket) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  CU(cudaSetDevice(e.device));
  const size_t n = (size_t)B * e.d.hop;
  if (int rc = ensure_pipe(e, n, B)) return rc;
  Engine::HostPipe& hp = e.pipe;
  const long long t = hp.submitted;
  const int k = (int)(t & 1);
  if (t >= 2) CU(cudaEventSynchronize(hp.d2h_done[k]));            // ticket t-2 delivered: staging set k is free
  if (slot_ids)
    for (int i = 0; i < B; ++i)
      if (slot_ids[i] < 0 || slot_ids[i] >= e.max_streams) return fail(DPDF_ERR_INVALID, "slot %d out of range", slot_ids[i]);
  CU(cudaMemcpyAsync(hp.in[k], pcm_in, n * sizeof(float), cudaMemcpyHostToDevice, hp.in_stream));
  if (slot_ids) CU(cudaMemcpyAsync(hp.slots[k], slot_ids, B * sizeof(int), cudaMemcpyHostToDevice, hp.in_stream));
  if (flags) CU(cudaMemcpyAsync(hp.flags[k], flags, B * sizeof(int), cudaMemcpyHostToDevice, hp.in_stream));
  CU(cudaEventRecord(hp.h2d_done[k], hp.in_stream));
  CU(cudaStreamWaitEvent(e.own_stream, hp.h2d_done[k], 0));
  if (int rc = dpdf_run_pcm(h, hp.in[k], e.d.hop, hp.out[k], e.d.hop, slot_ids ? hp.slots[k] : nullptr, flags ? hp.flags[k] : nullptr, B, 1,
                            e.own_stream))
    return rc;
  CU(cudaEventRecord(hp.step_done[k], e.own_stream));
  CU(cudaStreamWaitEvent(hp.out_stream, hp.step_done[k], 0));
  CU(cudaMemcpyAsync(pcm_out, hp.out[k], n * sizeof(float), cudaMemcpyDeviceToHost, hp.out_stream));
  CU(cudaEventRecord(hp.d2h_done[k], hp.out_stream));
  *ticket = t;
  ++hp.submitted;
  return 0;
}

extern "C" int dpdf_wait(dpdf_engine* h, int64_t ticket) {
  if (!h) return fail(DPDF_ERR_INVALID, "NULL engine");
  Engine& e = h->e;
  Engine::HostPipe& hp = e.pipe;
  if (ticket < 0 || ticket >= hp.submitted) return fail(DPDF_ERR_INVALID, "unknown ticket %lld", (long long)ticket);
  if (ticket + 2 < hp.submitted) return check_device_errors(e);     // already overwritten by a later ticket: delivered long ago
  CU(cudaSetDevice(e.device));
  CU(cudaEventSynchronize(hp.d2h_done[ticket & 1]));
  return check_device_errors(e);
}

// ---------------------------------------------------------------------------------------------
// reference-layout state (onnx_model/dpdfnet.py:737-746)
// ---------------------------------------------------------------------------------------------
extern "C" int dpdf_state_size(const dpdf_engine* h) { return h ? h->e.spec.state_size : 0; }

namespace {
struct Seg { float* base; size_t per; };
}

static int fetch(Engine& e, const float* base, size_t per, int slot, std::vector<float>& out) {
  out.resize(per);
  CU(cudaMemcpy(out.data(), base + (size_t)slot * per, per * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
static int store(Engine& e, float* base, size_t per, int slot, const std::vector<float>& in) {
  CU(cudaMemcpy(base + (size_t)slot * per, in.data(), per * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int dpdf_state_export(dpdf_engine* h, int32_t slot, float* flat) {
  if (!h || !flat) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (slot < 0 || slot >= e.max_streams) return fail(DPDF_ERR_INVALID, "slot %d out of range", slot);
  CU(cudaSetDevice(e.device));
  CU(cudaDeviceSynchronize());
  const Dims& d = e.d;
  int pos = 0;
  CU(cudaMemcpy(&pos, e.st.pos + slot, sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<float> buf;
  float* o = flat;
  auto plain = [&](const float* base, size_t per) -> int {
    if (int rc = fetch(e, base, per, slot, buf)) return rc;
    memcpy(o, buf.data(), per * sizeof(float));
    o += per;
    return 0;
  };
  // ring with L frames of `frame` floats: logical frame k (oldest first) = physical (pos + k) % L
  auto ring = [&](const float* base, int L, size_t frame) -> int {
    if (int rc = fetch(e, base, L * frame, slot, buf)) return rc;
    for (int k = 0; k < L; ++k) memcpy(o + k * frame, buf.data() + ((pos + k) % L) * frame, frame * sizeof(float));
    o += L * frame;
    return 0;
  };
  int rc = 0;
  if ((rc = plain(e.st.mu, d.fe_feat))) return rc;
  if ((rc = plain(e.st.s, NDF))) return rc;
  if ((rc = ring(e.st.erb_ring, 3, d.fe_feat))) return rc;
  if (d.N > 0 && (rc = plain(e.st.inter_erb, (size_t)d.N * d.fe[3] * C))) return rc;
  if ((rc = ring(e.st.df_ring, 3, 2 * NDF))) return rc;
  if (d.N > 0 && (rc = plain(e.st.inter_df, (size_t)d.N * (NDF / 2) * C))) return rc;
  if ((rc = plain(e.st.h_enc, H))) return rc;
  if ((rc = plain(e.st.h_erb, 2 * H))) return rc;
  if ((rc = plain(e.st.h_df, 2 * H))) return rc;
  {   // c0 ring: engine [5][96][64] -> reference [5][64][96]
    if ((rc = fetch(e, e.st.c0_ring, (size_t)ORD * NDF * C, slot, buf))) return rc;
    if (e.st.c0_fp16) {                            // compact FP16 frames at the start of the region -> FP32 (via a copy)
      std::vector<float> wide(buf.size());
      const __half* hsrc = reinterpret_cast<const __half*>(buf.data());
      for (size_t i = 0; i < wide.size(); ++i) wide[i] = __half2float(hsrc[i]);
      buf.swap(wide);
    }
    for (int k = 0; k < ORD; ++k) {
      const float* src = buf.data() + (size_t)((pos + k) % ORD) * NDF * C;
      for (int c = 0; c < C; ++c)
        for (int f = 0; f < NDF; ++f) o[((size_t)k * C + c) * NDF + f] = src[f * C + c];
    }
    o += (size_t)ORD * NDF * C;
  }
  if ((rc = ring(e.st.mask_ring, 3, (size_t)d.F * 2))) return rc;
  {   // coef ring: engine [3][96][5][2] -> reference [3][5][96][2]
    if ((rc = fetch(e, e.st.coef_ring, (size_t)3 * NDF * 2 * ORD, slot, buf))) return rc;
    for (int k = 0; k < 3; ++k) {
      const float* src = buf.data() + (size_t)((pos + k) % 3) * NDF * 2 * ORD;
      for (int n = 0; n < ORD; ++n)
        for (int f = 0; f < NDF; ++f)
          for (int r = 0; r < 2; ++r) o[(((size_t)k * ORD + n) * NDF + f) * 2 + r] = src[(f * ORD + n) * 2 + r];
    }
    o += (size_t)3 * NDF * 2 * ORD;
  }
  if ((rc = ring(e.st.dfspec_ring, ORD, (size_t)d.F * 2))) return rc;
  if ((o - flat) != e.spec.state_size) return fail(DPDF_ERR_INVALID, "internal: exported %ld floats, expected %d", (long)(o - flat), e.spec.state_size);
  return 0;
}

extern "C" int dpdf_state_import(dpdf_engine* h, int32_t slot, const float* flat) {
  if (!h || !flat) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (slot < 0 || slot >= e.max_streams) return fail(DPDF_ERR_INVALID, "slot %d out of range", slot);
  CU(cudaSetDevice(e.device));
  CU(cudaDeviceSynchronize());
  const Dims& d = e.d;
  const float* o = flat;
  std::vector<float> buf;
  int rc = 0;
  auto plain = [&](float* base, size_t per) -> int {
    buf.assign(o, o + per);
    o += per;
    return store(e, base, per, slot, buf);
  };
  if ((rc = plain(e.st.mu, d.fe_feat))) return rc;
  if ((rc = plain(e.st.s, NDF))) return rc;
  if ((rc = plain(e.st.erb_ring, 3 * d.fe_feat))) return rc;            // pos := 0 => logical == physical
  if (d.N > 0 && (rc = plain(e.st.inter_erb, (size_t)d.N * d.fe[3] * C))) return rc;
  if ((rc = plain(e.st.df_ring, 3 * 2 * NDF))) return rc;
  if (d.N > 0 && (rc = plain(e.st.inter_df, (size_t)d.N * (NDF / 2) * C))) return rc;
  if ((rc = plain(e.st.h_enc, H))) return rc;
  if ((rc = plain(e.st.h_erb, 2 * H))) return rc;
  if ((rc = plain(e.st.h_df, 2 * H))) return rc;
  {
    buf.resize((size_t)ORD * NDF * C);
    for (int k = 0; k < ORD; ++k)
      for (int c = 0; c < C; ++c)
        for (int f = 0; f < NDF; ++f) buf[((size_t)k * NDF + f) * C + c] = o[((size_t)k * C + c) * NDF + f];
    o += buf.size();
    if (e.st.c0_fp16) {                            // the pending sums below use the FP32 values; the device ring gets the rounded frames
      std::vector<float> packed(buf.size(), 0.f);
      __half* hdst = reinterpret_cast<__half*>(packed.data());
      for (size_t i = 0; i < buf.size(); ++i) hdst[i] = __float2half_rn(buf[i]);
      if ((rc = store(e, e.st.c0_ring, packed.size(), slot, packed))) return rc;
    } else if ((rc = store(e, e.st.c0_ring, buf.size(), slot, buf))) return rc;
    // pending sums of the df pathway conv, rebuilt from the five imported frames (pos := 0, logical == physical):
    // the output m hops ahead already has the taps kt = 0 .. 3 - m of the frames kt + 1 + m
    std::vector<float> acc((size_t)ORD * NDF * 10, 0.f);
    if (e.dfp_w_host.size() == (size_t)10 * ORD * 32) {
      for (int m = 0; m < ORD - 1; ++m)
        for (int f = 0; f < NDF; ++f)
          for (int oo = 0; oo < 10; ++oo) {
            double sum = 0.0;
            for (int kt = 0; kt + m < ORD - 1; ++kt)
              for (int ci = 0; ci < 32; ++ci)
                sum += (double)e.dfp_w_host[((size_t)oo * ORD + kt) * 32 + ci] *
                       buf[((size_t)(kt + 1 + m) * NDF + f) * C + (oo / 5) * 32 + ci];
            acc[((size_t)f * ORD + m) * 10 + oo] = (float)sum;
          }
    }
    if ((rc = store(e, e.st.dfp_acc, acc.size(), slot, acc))) return rc;
  }
  if ((rc = plain(e.st.mask_ring, (size_t)3 * d.F * 2))) return rc;
  {
    buf.resize((size_t)3 * NDF * 2 * ORD);
    for (int k = 0; k < 3; ++k)
      for (int n = 0; n < ORD; ++n)
        for (int f = 0; f < NDF; ++f)
          for (int r = 0; r < 2; ++r) buf[(((size_t)k * NDF + f) * ORD + n) * 2 + r] = o[(((size_t)k * ORD + n) * NDF + f) * 2 + r];
    o += buf.size();
    if ((rc = store(e, e.st.coef_ring, buf.size(), slot, buf))) return rc;
  }
  if ((rc = plain(e.st.dfspec_ring, (size_t)ORD * d.F * 2))) return rc;
  const int zero = 0;
  CU(cudaMemcpy(e.st.pos + slot, &zero, sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
extern "C" int dpdf_debug_tensor(dpdf_engine* h, const char* name, float* out, size_t max_floats, size_t* numel) {
  if (!h || !name) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  const Dims& d = e.d;
  const Scratch& c = e.sc;
  struct Ent { const char* n; const float* p; size_t per; };
  const Ent table[] = {
      {"e0", c.e0, (size_t)d.fe[0] * C}, {"e1", c.e1, (size_t)d.fe[1] * C}, {"e2", c.e2, (size_t)d.fe[2] * C},
      {"e3", c.e3, (size_t)d.fe[3] * C}, {"c0", c.c0, (size_t)NDF * C}, {"xd", c.c1, (size_t)(NDF / 2) * C},
      {"xe", c.xe, (size_t)d.fe[3] * C}, {"hcat_e", c.hcat_e, (size_t)d.fe[3] * 2 * C}, {"hcat_d", c.hcat_d, (size_t)(NDF / 2) * 2 * C},
      {"cemb", c.cemb, 512}, {"emb", c.emb, 512}, {"ed", d.hr48 ? c.ed2 : c.ed, d.hr48 ? (size_t)d.fe[3] * C : 512},
      {"d3", c.d3, (size_t)d.fe[2] * C}, {"d2", c.d2, (size_t)d.fe[1] * C}, {"d1", c.d1, (size_t)d.fe[0] * C},
      {"m", c.m, (size_t)d.fe[0]}, {"co", c.co, (size_t)NDF * 2 * ORD}, {"g0", c.g0, H}, {"henc", c.henc, H},
      {"hdf2", c.hdf2, H}, {"herb2", c.herb2, H}, {"cc", c.cc, H}};
  for (const Ent& t : table) {
    if (strcmp(t.n, name) == 0) {
      if (numel) *numel = t.per;
      if (out) {
        const size_t n = t.per * (size_t)e.last_B;
        if (n > max_floats) return fail(DPDF_ERR_INVALID, "buffer too small for %s: need %zu floats", name, n);
        CU(cudaSetDevice(e.device));
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(out, t.p, n * sizeof(float), cudaMemcpyDeviceToHost));
      }
      return 0;
    }
  }
  return fail(DPDF_ERR_INVALID, "unknown debug tensor '%s'", name);
}

extern "C" int dpdf_kernel_launches(const dpdf_engine* h) { return h ? h->e.launches : 0; }

extern "C" int dpdf_set_option(dpdf_engine* h, const char* key, int32_t value) {
  if (!h || !key) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (strcmp(key, "graph") == 0) e.use_graph = value ? 1 : 0;
  else if (strcmp(key, "intra_bt") == 0) {
    if (value != 0 && value != 8 && value != 16 && value != 32) return fail(DPDF_ERR_INVALID, "intra_bt must be 0, 8, 16 or 32");
    e.intra_bt = value;
    drop_graphs(e);
  } else if (strcmp(key, "intra_tc") == 0 || strcmp(key, "intra_tc_min") == 0) {
    if (key[8] == 0) {
      if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "intra_tc must be 0 (FFMA2), 1 (tcgen05) or 2 (by batch size)");
      e.intra_tc = value;
    } else {
      e.intra_tc_min = value;
    }
    drop_graphs(e);
  } else if (strcmp(key, "dft_tc") == 0 || strcmp(key, "dft_tc_min") == 0) {
    if (key[6] == 0) {
      if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "dft_tc must be 0 (FFMA2), 1 (tcgen05) or 2 (by batch size)");
      if (value && !e.w.dft_fwd_tc) return fail(DPDF_ERR_INVALID, "the weight blob has no tensor-core DFT images (const.dft_fwd_tc / const.dft_inv_tc)");
      e.dft_tc = value;
    } else {
      e.dft_tc_min = value;
    }
    drop_graphs(e);
  } else if (strcmp(key, "encoder_fork") == 0) {
    e.encoder_fork = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "stop_after") == 0) {
    e.stop_after = value;                                    // profiling only (tools/chain_profile.py): the streams' state is garbage afterwards
    drop_graphs(e);
  } else if (strcmp(key, "post_dual") == 0) {
    if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "post_dual must be 0 (one tile per CTA), 1 (two) or 2 (two when not overlapped with the sweep)");
    e.post_dual = value;
    drop_graphs(e);
  } else if (strcmp(key, "post_res") == 0) {
    e.post_res = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "gru_uc") == 0) {
    if (value != 0 && value != 32 && value != 64) return fail(DPDF_ERR_INVALID, "gru_uc must be 0 (auto), 32 or 64");
    e.gru_uc = value;
    drop_graphs(e);
  } else if (strcmp(key, "post_pair") == 0) {
    if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "post_pair must be 0 (single CTAs), 1 (CTA pairs) or 2 (pairs when not overlapped with the sweep)");
    e.post_pair = value;
    drop_graphs(e);
  } else if (strcmp(key, "intra_pdl") == 0) {
    e.intra_pdl = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "dfp_early") == 0) {
    e.dfp_early = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "sweep_prio") == 0) {
    e.sweep_prio = value;
    drop_graphs(e);
  } else if (strcmp(key, "lane_min") == 0) {
    if (value < 32 || value % 32) return fail(DPDF_ERR_INVALID, "lane_min must be a positive multiple of 32");
    e.lane_min = value;
    drop_graphs(e);
  } else if (strcmp(key, "frag_max") == 0) {
    e.frag_max = value;
    drop_graphs(e);
  } else if (strcmp(key, "intra_frag_erb") == 0) {
    e.intra_frag_erb = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "intra_frag") == 0) {
    e.intra_frag = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "intra_sr") == 0) {
    if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "intra_sr must be 0 (off), 1 (whenever rows are duplicated) or 2 (with intra_dup = 4 only)");
    e.intra_sr = value;
    drop_graphs(e);
  } else if (strcmp(key, "intra_dup") == 0) {
    if (value != 0 && value != 1 && value != 2 && value != 4) return fail(DPDF_ERR_INVALID, "intra_dup must be 0 (auto), 1, 2 or 4");
    e.intra_dup = value;
    drop_graphs(e);
  } else if (strcmp(key, "gru_tc") == 0) {
    if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "gru_tc must be 0 (FFMA2), 1 (tcgen05) or 2 (by batch size)");
    e.gru_tc = value;
    drop_graphs(e);
  } else if (strcmp(key, "sep_tc") == 0) {
    if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "sep_tc must be 0 (FFMA2), 1 (tcgen05) or 2 (by batch size)");
    e.sep_tc = value;
    drop_graphs(e);
  } else if (strcmp(key, "sep_tma") == 0) {
    e.sep_tma = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "decoder_fork") == 0) {
    e.decoder_fork = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "c0_fp16") == 0) {
    e.st.c0_fp16 = value ? 1 : 0;                  // switch only on freshly reset streams: the ring changes its storage format
    drop_graphs(e);
  } else if (strcmp(key, "dfp_ps") == 0) {
    e.dfp_ps = value ? 1 : 0;                      // switch only on freshly reset streams: the two forms keep different state
    drop_graphs(e);
  } else if (strcmp(key, "tail_pdl") == 0) {
    e.tail_pdl = value ? 1 : 0;
    drop_graphs(e);
  } else if (strcmp(key, "pdl") == 0) {
    if (value < 0 || value > 2) return fail(DPDF_ERR_INVALID, "pdl must be 0 (off), 1 (whole hop) or 2 (all segments but the DPRNN stack)");
    e.pdl = value;
    drop_graphs(e);
  } else if (strcmp(key, "overlap") == 0 || strcmp(key, "overlap_max") == 0) {
    if (key[7] == 0) e.overlap = value < 0 ? 0 : (value > 2 ? 2 : value);
    else e.overlap_max = value;
    drop_graphs(e);
  } else if (strcmp(key, "free_lanes") == 0) {
    e.free_lanes = value ? 1 : 0;
  } else if (strcmp(key, "lanes") == 0) {
    if (value < 0 || value > Engine::MAX_LANES) return fail(DPDF_ERR_INVALID, "lanes must be 0 (auto) .. %d", Engine::MAX_LANES);
    e.lanes = value;
    drop_graphs(e);
  } else if (strcmp(key, "ana_force") == 0 || strcmp(key, "syn_force") == 0) {
    (key[0] == 'a' ? e.ana_force : e.syn_force) = value;
    drop_graphs(e);
  } else if (strcmp(key, "ana_nb") == 0 || strcmp(key, "syn_sb") == 0) {
    (key[0] == 'a' ? e.ana_nb : e.syn_sb) = value;
    drop_graphs(e);
  } else if (strcmp(key, "post_pf") == 0) {
    if (value < 0 || value > 16) return fail(DPDF_ERR_INVALID, "post_pf must be 0 (off) .. 16");
    e.post_pf = value;
    drop_graphs(e);
  } else if (strcmp(key, "post_tc") == 0) {
    e.post_tc = value ? 1 : 0;
    drop_graphs(e);
  } else return fail(DPDF_ERR_INVALID, "unknown option '%s'", key);
  return 0;
}

// Per-kernel device time of a hop at batch B: runs `iters` un-graphed hops on the engine's staging
// buffers with a CUDA-event pair around every launch.  Advances the state of slots 0..B-1.
extern "C" int dpdf_time_kernels(dpdf_engine* h, int32_t B, int32_t iters, float* ms_out, const char** names_out,
                                 int32_t max_entries, int32_t* n_entries) {
  if (!h || !ms_out || !names_out || !n_entries) return fail(DPDF_ERR_INVALID, "NULL argument");
  Engine& e = h->e;
  if (int rc = check_batch(e, B)) return rc;
  CU(cudaSetDevice(e.device));
  const size_t n = (size_t)B * e.d.hop;
  if (int rc = ensure_stage(e, n, n)) return rc;
  cudaStream_t st = e.own_stream;
  CU(cudaMemsetAsync(e.stage_in, 0, n * sizeof(float), st));
  struct OneLane {                           // the whole batch as a single kernel chain: per-kernel times of B streams
    Engine& e; int saved;
    explicit OneLane(Engine& e_) : e(e_), saved(e_.lanes) { e.lanes = 1; }
    ~OneLane() { e.lanes = saved; }
  } one_lane(e);
  static std::vector<std::string> keep;     // storage for the returned names
  std::vector<double> acc;
  keep.clear();
  e.last_B = B;
  for (int it = 0; it < iters + 1; ++it) {
    if (int rc = set_io(e, e.stage_in, e.d.hop, e.stage_out, e.d.hop, nullptr, nullptr, 0, B, st)) return rc;
    e.timing = true;
    e.tev.clear();
    e.tnames.clear();
    enqueue_step(e, B, st);
    e.timing = false;
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (acc.empty()) { acc.assign(e.tnames.size(), 0.0); keep = e.tnames; }
    for (size_t i = 0; i < e.tnames.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e.tev[2 * i], e.tev[2 * i + 1]);
      if (it > 0) acc[i] += ms;         // first iteration is warm-up
      cudaEventDestroy(e.tev[2 * i]);
      cudaEventDestroy(e.tev[2 * i + 1]);
    }
    e.tev.clear();
  }
  const int cnt = (int)std::min<size_t>(keep.size(), (size_t)max_entries);
  for (int i = 0; i < cnt; ++i) {
    ms_out[i] = (float)(acc[i] / std::max(1, iters));
    names_out[i] = keep[i].c_str();
  }
  *n_entries = cnt;
  return 0;
}
