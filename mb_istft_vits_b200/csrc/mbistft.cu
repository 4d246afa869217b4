// mbistft.cu -- the C ABI of include/mbistft.h: handle, weight packer, workspace planner and the launch
// sequences for flow-reverse (models.py:207-214) and the three iSTFT decoders (models.py:248-474).
//
// Nothing here computes on the host: without a CUDA device every compute entry returns MBV_ERR_CUDA.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/mbistft.h"
#include "common.cuh"
#include "kernels.h"

using namespace mbv;


namespace {

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- host float -> operand conversions (RNE)
inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
// fp32 -> fp16 bits, round to nearest even, overflow saturates to +-65504 (like the device-side stores)
inline uint16_t f32_to_f16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t a = u & 0x7fffffffu;
  if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);          // NaN
  if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);         // >= 65520 (or inf) -> max finite
  if (a < 0x33000001u) return (uint16_t)sign;                      // < 2^-25 -> 0
  int e = (int)(a >> 23) - 127;
  uint32_t m = (a & 0x7fffffu) | 0x800000u;
  int shift = (e < -14) ? (13 + (-14 - e)) : 13;                   // subnormal halves lose extra bits
  uint32_t half_m = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_m & 1u))) half_m++;
  uint32_t out;
  if (e < -14) out = half_m;                                       // subnormal (a carry makes it the smallest normal)
  else out = ((uint32_t)(e + 15) << 10) + (half_m - 0x400u);       // a mantissa carry bumps the exponent
  return (uint16_t)(sign | out);
}
inline float f32_to_tf32(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return f;
  u += 0xfffu + ((u >> 13) & 1u);
  u &= 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

struct ConvLayer {
  // geometry
  int Cp_in = 0, N_total = 0, taps = 0, dil = 1, n_phases = 1, gate = 0;
  int shift0[kMaxPhases] = {0};
  int n_valid = 0;      // real output channels (unpadded)
  double macs_per_row = 0;  // real MACs per computed row (all phases), for FLOP accounting
  // device data
  void* w = nullptr;    // packed operand weights [phase][tap][N_total][Cp_in]
  float* bias = nullptr;  // [N_total] (gate: [tanh Hp | sigmoid Hp])
};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

}  // namespace

struct mbv_handle {
  mbv_config cfg;
  int prec = 0;
  int esize = 4;
  int res_half = 0;  // decoder residual stream in fp16 (MBV_FLAG_RESIDUAL_FP16 with bf16; always with fp16)
  int single = 0;    // fp16 precision on the tensor-core path: residual streams live in the operand tensors (EpiParams::res_half 2)
  int rsize = 4;
  int num_sms = 148;
  bool weights_loaded = false;
  std::string err;
  int last_launches = 0;
  std::vector<void*> dev_allocs;

  // derived geometry
  int Cz = 0;    // inter channels (192), must be a multiple of 64
  int H = 0, Hp = 0, Hw = 0;  // flow hidden width: real, padded to 64 (activation pitch), padded to 128 (weight rows)
  int n_stage = 0;
  int stage_C[MBV_MAX_UPS] = {0};
  int n_logit = 0;  // subbands * 18
  int spf = 0;      // samples per latent frame

  // decoder layers
  ConvLayer conv_pre, conv_post;
  ConvLayer ups[MBV_MAX_UPS];
  // resblocks[stage][kernel]: convs1[p], convs2[p] (type 1) or convs[p] in c1 (type 2)
  ConvLayer rb_c1[MBV_MAX_UPS][MBV_MAX_KERNELS][MBV_MAX_DILATIONS];
  ConvLayer rb_c2[MBV_MAX_UPS][MBV_MAX_KERNELS][MBV_MAX_DILATIONS];
  float* rb_cond_w[MBV_MAX_UPS][MBV_MAX_KERNELS] = {{nullptr}};  // [C][gin]
  float* rb_cond_b[MBV_MAX_UPS][MBV_MAX_KERNELS] = {{nullptr}};  // [C]
  float post_bias[72];
  float tail_coef[4][64];
  float tail_mod[8][4];
  float tail_g2[4][16];
  int tail_fast = 0;

  // flow layers, indexed by coupling layer 0..3 (reference order)
  ConvLayer fl_pre[4], fl_post[4], fl_in[4][4], fl_rs[4][4];
  float* fl_cond_w[4] = {nullptr};  // [L*2Hp][gin] packed in the gate row order
  float* fl_cond_b[4] = {nullptr};  // [L*2Hp]

  // posterior encoder (models.py:217-246; optional: loaded when the state-dict carries enc_q.*): pre 1x1 -> 16-layer WN ->
  // proj 1x1, the WN skip path collapsed into proj like the flow's into post
  bool has_enc_q = false;
  int eq_layers = 16;        // hard-coded at models.py:646 (kernel 5, dilation_rate 1, 16 layers)
  int eq_spec = 0, eq_spec_p = 0;   // spectrogram channels (513) and their 64-padded pitch
  ConvLayer eq_pre, eq_proj, eq_in[MBV_MAX_ENCQ_LAYERS], eq_rs[MBV_MAX_ENCQ_LAYERS];
  float* eq_cond_w = nullptr;
  float* eq_cond_b = nullptr;

  // text encoder (models.py:140-181; optional: loaded when the state-dict carries enc_p.*).  Geometry is read off the
  // tensors: layers = number of attn_layers, heads = hidden / emb_rel_k.shape[2], window = (emb_rel_k.shape[1] - 1) / 2
  bool has_enc_p = false;
  int tp_layers = 0, tp_heads = 0, tp_window = 0, tp_filter = 0, tp_vocab = 0;
  float* tp_emb = nullptr;
  ConvLayer tp_qkv[MBV_MAX_TEXT_LAYERS], tp_o[MBV_MAX_TEXT_LAYERS], tp_f1[MBV_MAX_TEXT_LAYERS], tp_f2[MBV_MAX_TEXT_LAYERS], tp_proj;
  float* tp_relk[MBV_MAX_TEXT_LAYERS] = {nullptr};
  float* tp_relv[MBV_MAX_TEXT_LAYERS] = {nullptr};
  float* tp_ln[MBV_MAX_TEXT_LAYERS][4] = {{nullptr}};  // gamma1, beta1, gamma2, beta2

  // tensor-map cache: valid while (B, T, ws) stay the same
  struct PlanKey { int B, T; void* ws; int kind; bool operator<(const PlanKey& o) const {
    if (B != o.B) return B < o.B; if (T != o.T) return T < o.T; if (ws != o.ws) return ws < o.ws; return kind < o.kind; } };
  std::map<PlanKey, std::vector<TcPlan>> plan_cache;
  std::map<PlanKey, std::vector<TcPairPlan>> pair_cache;
  int use_branches = 0;  // the parallel ResBlocks of a stage run on separate streams (run_decode)
  cudaStream_t br_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  int use_pair = 0;  // fused ResBlock conv pairs on 128-channel stages (conv_pair_kernel)
  int pair_max_taps = 17;

  // per-launch device timing (mbv_set_profiling)
  bool profiling = false;
  struct ProfRec { int kind; cudaEvent_t e0, e1; char desc[56]; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

namespace {

int fail(mbv_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

// Every entry point runs on the handle's device and leaves the caller's current device as it found it (PyTorch reads it
// with cudaGetDevice: an Engine on device k must not silently move the caller's allocations and default stream there).
struct DeviceGuard {
  int prev = -1, dev;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int d) : dev(d) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != d) { err = cudaSetDevice(d); switched = (err == cudaSuccess); }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_GUARD(h)                                                                                            \
  DeviceGuard _dev_guard((h)->cfg.device);                                                                         \
  if (_dev_guard.err != cudaSuccess) return fail(h, MBV_ERR_CUDA, "cudaSetDevice(%d): %s", (h)->cfg.device, cudaGetErrorString(_dev_guard.err))

#define CUDA_TRY(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(h, MBV_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)

// upload packed weights (host fp32, layout [phase][tap][N_total][Cp_in]) in the operand type
int upload_weights(mbv_handle* h, ConvLayer& L, const std::vector<float>& packed, const std::vector<float>& bias) {
  const size_t n = packed.size();
  void* d = nullptr;
  if (h->prec == MBV_PREC_BF16 || h->prec == MBV_PREC_FP16) {
    std::vector<uint16_t> tmp(n);
    if (h->prec == MBV_PREC_BF16) for (size_t i = 0; i < n; ++i) tmp[i] = f32_to_bf16(packed[i]);
    else for (size_t i = 0; i < n; ++i) tmp[i] = f32_to_f16(packed[i]);
    CUDA_TRY(h, cudaMalloc(&d, n * 2));
    CUDA_TRY(h, cudaMemcpy(d, tmp.data(), n * 2, cudaMemcpyHostToDevice));
  } else {
    std::vector<float> tmp(packed);
    if (h->prec == MBV_PREC_TF32)
      for (size_t i = 0; i < n; ++i) tmp[i] = f32_to_tf32(tmp[i]);
    CUDA_TRY(h, cudaMalloc(&d, n * 4));
    CUDA_TRY(h, cudaMemcpy(d, tmp.data(), n * 4, cudaMemcpyHostToDevice));
  }
  h->dev_allocs.push_back(d);
  L.w = d;
  float* db = nullptr;
  CUDA_TRY(h, cudaMalloc(&db, bias.size() * 4));
  CUDA_TRY(h, cudaMemcpy(db, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  h->dev_allocs.push_back(db);
  L.bias = db;
  return MBV_OK;
}

int upload_f32(mbv_handle* h, const std::vector<float>& v, float** out) {
  float* d = nullptr;
  CUDA_TRY(h, cudaMalloc(&d, v.size() * 4));
  CUDA_TRY(h, cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  h->dev_allocs.push_back(d);
  *out = d;
  return MBV_OK;
}

typedef std::map<std::string, HostTensor> TensorMap;

const HostTensor* find_tensor(mbv_handle* h, const TensorMap& m, const std::string& name, int rank, const int64_t* shape) {
  auto it = m.find(name);
  if (it == m.end()) {
    fail(h, MBV_ERR_WEIGHTS, "missing tensor %s", name.c_str());
    return nullptr;
  }
  const HostTensor& t = it->second;
  bool ok = (int)t.shape.size() == rank;
  for (int i = 0; ok && i < rank; ++i) ok = t.shape[i] == shape[i];
  if (!ok) {
    fail(h, MBV_ERR_WEIGHTS, "tensor %s has the wrong shape", name.c_str());
    return nullptr;
  }
  return &t;
}

// Conv1d [O][I][K] 'same', dilation d  ->  taps K, shift0 = -d(K-1)/2.  Output rows may be permuted / padded:
// out_map[n] = source output channel of packed row n (or -1 = zero row); in_map[c] = source input channel of
// packed column c (or -1).
int pack_conv1d(mbv_handle* h, const TensorMap& m, const std::string& prefix, int O, int I, int K, int dil,
                const std::vector<int>& out_map, const std::vector<int>& in_map, bool has_bias, int gate, ConvLayer* L) {
  const int64_t wshape[3] = {O, I, K};
  const HostTensor* w = find_tensor(h, m, prefix + ".weight", 3, wshape);
  if (!w) return MBV_ERR_WEIGHTS;
  const HostTensor* b = nullptr;
  if (has_bias) {
    const int64_t bshape[1] = {O};
    b = find_tensor(h, m, prefix + ".bias", 1, bshape);
    if (!b) return MBV_ERR_WEIGHTS;
  }
  // packed rows: a multiple of 128 (one UMMA M tile); a gate conv passes [tanh half | sigmoid half], each padded
  const int n_map = (int)out_map.size(), Cp = (int)in_map.size();
  const int N = round_up(n_map, 128);
  if (gate && (n_map % 128) != 0) return fail(h, MBV_ERR_INVALID, "%s: gate rows must come in 128-row tiles", prefix.c_str());
  L->Cp_in = Cp; L->N_total = N; L->taps = K; L->dil = dil; L->n_phases = 1; L->gate = gate;
  L->shift0[0] = -dil * (K - 1) / 2;
  L->n_valid = O;
  L->macs_per_row = (double)O * I * K;
  std::vector<float> packed((size_t)K * N * Cp, 0.f), bias(N, 0.f);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < n_map; ++n) {
      const int o = out_map[n];
      if (o < 0) continue;
      float* dst = &packed[((size_t)k * N + n) * Cp];
      for (int c = 0; c < Cp; ++c) {
        const int i = in_map[c];
        if (i >= 0) dst[c] = w->data[((size_t)o * I + i) * K + k];
      }
    }
  if (b)
    for (int n = 0; n < n_map; ++n)
      if (out_map[n] >= 0) bias[n] = b->data[out_map[n]];
  return upload_weights(h, *L, packed, bias);
}

std::vector<int> iota_pad(int n, int padded) {
  std::vector<int> v(padded, -1);
  for (int i = 0; i < n; ++i) v[i] = i;
  return v;
}

// ConvTranspose1d [I][O][K], stride S, padding (K-S)/2 as S polyphase branches of K/S taps
int pack_convT(mbv_handle* h, const TensorMap& m, const std::string& prefix, int I, int O, int K, int S, ConvLayer* L) {
  const int64_t wshape[3] = {I, O, K};
  const int64_t bshape[1] = {O};
  const HostTensor* w = find_tensor(h, m, prefix + ".weight", 3, wshape);
  const HostTensor* b = w ? find_tensor(h, m, prefix + ".bias", 1, bshape) : nullptr;
  if (!w || !b) return MBV_ERR_WEIGHTS;
  if (K % S != 0 || (K - S) % 2 != 0 || S > kMaxPhases)
    return fail(h, MBV_ERR_UNSUPPORTED, "%s: upsampler kernel %d / stride %d not supported (need K %% S == 0, K-S even, S <= %d)",
                prefix.c_str(), K, S, kMaxPhases);
  const int P = (K - S) / 2, TP = K / S;
  const int Cp = round_up(I, 64), N = round_up(O, 128);
  L->Cp_in = Cp; L->N_total = N; L->taps = TP; L->dil = 1; L->n_phases = S; L->gate = 0;
  L->n_valid = O;
  L->macs_per_row = (double)I * O * K;  // per input row, all S phases together
  std::vector<float> packed((size_t)S * TP * N * Cp, 0.f), bias(N, 0.f);
  for (int r = 0; r < S; ++r) {
    const int base = (r + P) / S, k0 = (r + P) % S;
    L->shift0[r] = base - (TP - 1);
    for (int j = 0; j < TP; ++j) {
      const int k = k0 + S * (TP - 1 - j);  // tap j reads input row q + shift0 + j
      for (int o = 0; o < O; ++o) {
        float* dst = &packed[(((size_t)r * TP + j) * N + o) * Cp];
        for (int i = 0; i < I; ++i) dst[i] = w->data[((size_t)i * O + o) * K + k];
      }
    }
  }
  for (int o = 0; o < O; ++o) bias[o] = b->data[o];
  return upload_weights(h, *L, packed, bias);
}

// modified Bessel I0 (power series), for the Kaiser window of pqmf.py:40
double bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 200; ++k) {
    term *= q / ((double)k * k);
    sum += term;
    if (term < 1e-20 * sum) break;
  }
  return sum;
}

// pqmf.py:15-43 + 64-79: 63-tap Kaiser(beta 9) prototype at cutoff 0.15, cosine-modulated synthesis bank (float64 -> float32)
void design_pqmf_synthesis(float hs[4][63], double proto[63]) {
  const int taps = 62;
  const double cutoff = 0.15, beta = 9.0, pi = 3.14159265358979323846;
  for (int n = 0; n <= taps; ++n) {
    const double x = n - 0.5 * taps;
    const double hi = (n == taps / 2) ? cutoff : sin(pi * cutoff * x) / (pi * x);
    const double r = 2.0 * n / taps - 1.0;
    const double wk = bessel_i0(beta * sqrt(1.0 - r * r)) / bessel_i0(beta);
    proto[n] = hi * wk;
  }
  for (int k = 0; k < 4; ++k)
    for (int n = 0; n <= taps; ++n) {
      const double v = 2.0 * proto[n] * cos((2 * k + 1) * (pi / 8.0) * (n - (taps - 1) / 2.0) - ((k & 1) ? -1.0 : 1.0) * pi / 4.0);
      hs[k][n] = (float)v;
    }
}

// polyphase table of the synthesis FIR: G[c][r*16 + (d+7)] = 4 * h[c][4d + 31 - r]  (tail.cu phase 3)
void fill_tail_coef(mbv_handle* h, const float hs[4][63]) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      for (int d = -7; d <= 8; ++d) {
        const int k = 4 * d + 31 - r;
        h->tail_coef[c][r * 16 + d + 7] = (k >= 0 && k <= 62) ? 4.f * hs[c][k] : 0.f;
      }
}


// One WaveNet stack (modules.WN, modules.py:111-176) feeding a 1x1 projection `post` (ResidualCouplingLayer.post /
// PosteriorEncoder.proj).  Packs, for l < NL: the gate conv in_layers[l] as 128-row tiles of [64 tanh | 64 sigmoid] rows and,
// for l < NL-1, the RESIDUAL half of res_skip_layers[l].  The skip halves of all layers and `post` are linear in the gate
// outputs, so they collapse into ONE conv over the concatenated gate outputs:
//   post(sum_l skip_l) = sum_l (W_post W_skip,l) acts_l + (W_post sum_l b_skip,l + b_post)      (modules.py:169-176)
// post_map[n] = source row of `post` for packed output row n.  The running skip sum never exists in memory and the last
// layer has no res/skip conv at all.
int pack_wn(mbv_handle* h, const TensorMap& m, const std::string& enc, int NL, int K, const std::string& post_name, int post_out,
            const std::vector<int>& post_map, ConvLayer* in_layers, ConvLayer* rs_layers, ConvLayer* post, float** cond_w,
            float** cond_b) {
  const int H = h->H, Hp = h->Hp, gin = h->cfg.gin_channels;
  int rc;
  char pfx[128];
  std::vector<int> hin = iota_pad(H, Hp);
  for (int l = 0; l < NL; ++l) {
    // gate rows: 128-row MMA tiles of [64 tanh rows | 64 sigmoid rows] for 64 consecutive channels, so ONE
    // accumulator holds both halves of a channel (lanes i and i+64) and H = 192 = 3 x 64 needs no padding
    std::vector<int> gmap(2 * Hp, -1);
    for (int o = 0; o < H; ++o) { gmap[128 * (o / 64) + (o % 64)] = o; gmap[128 * (o / 64) + 64 + (o % 64)] = H + o; }
    snprintf(pfx, sizeof(pfx), "%s.in_layers.%d", enc.c_str(), l);
    rc = pack_conv1d(h, m, pfx, 2 * H, H, K, 1, gmap, hin, true, 1, &in_layers[l]);
    if (rc) return rc;
    snprintf(pfx, sizeof(pfx), "%s.res_skip_layers.%d", enc.c_str(), l);
    if (l < NL - 1) {
      rc = pack_conv1d(h, m, pfx, 2 * H, H, 1, 1, iota_pad(H, Hp), hin, true, 0, &rs_layers[l]);
      if (rc) return rc;
      rs_layers[l].n_valid = H;
    }
    rs_layers[l].macs_per_row = (double)(l < NL - 1 ? 2 * H : H) * H;  // algorithmic MACs of the reference layer
  }
  {
    const int64_t pws[3] = {post_out, H, 1}, pbs[1] = {post_out};
    const HostTensor* pw = find_tensor(h, m, post_name + ".weight", 3, pws);
    const HostTensor* pb = pw ? find_tensor(h, m, post_name + ".bias", 1, pbs) : nullptr;
    if (!pw || !pb) return MBV_ERR_WEIGHTS;
    HostTensor fw, fb;
    fw.shape = {post_out, (int64_t)NL * H, 1};
    fw.data.assign((size_t)post_out * NL * H, 0.f);
    fb.shape = {post_out};
    fb.data.assign(post_out, 0.f);
    std::vector<double> bsum(H, 0.0);
    for (int l = 0; l < NL; ++l) {
      const bool lastl = (l == NL - 1);
      const int64_t ws[3] = {lastl ? H : 2 * H, H, 1}, bs[1] = {lastl ? H : 2 * H};
      snprintf(pfx, sizeof(pfx), "%s.res_skip_layers.%d", enc.c_str(), l);
      const HostTensor* w = find_tensor(h, m, std::string(pfx) + ".weight", 3, ws);
      const HostTensor* b = w ? find_tensor(h, m, std::string(pfx) + ".bias", 1, bs) : nullptr;
      if (!w || !b) return MBV_ERR_WEIGHTS;
      const int r0 = lastl ? 0 : H;  // first skip row
      for (int j2 = 0; j2 < H; ++j2) bsum[j2] += b->data[r0 + j2];
      std::vector<double> acc(H);
      for (int o = 0; o < post_out; ++o) {
        for (int cidx = 0; cidx < H; ++cidx) acc[cidx] = 0.0;
        for (int j2 = 0; j2 < H; ++j2) {
          const double pv = pw->data[(size_t)o * H + j2];
          const float* wr = &w->data[(size_t)(r0 + j2) * H];
          for (int cidx = 0; cidx < H; ++cidx) acc[cidx] += pv * (double)wr[cidx];
        }
        for (int cidx = 0; cidx < H; ++cidx) fw.data[(size_t)o * NL * H + (size_t)l * H + cidx] = (float)acc[cidx];
      }
    }
    for (int o = 0; o < post_out; ++o) {
      double acc = pb->data[o];
      for (int j2 = 0; j2 < H; ++j2) acc += (double)pw->data[(size_t)o * H + j2] * bsum[j2];
      fb.data[o] = (float)acc;
    }
    TensorMap fused;
    fused["post_fused.weight"] = std::move(fw);
    fused["post_fused.bias"] = std::move(fb);
    std::vector<int> fin((size_t)NL * Hp, -1);
    for (int l = 0; l < NL; ++l)
      for (int cidx = 0; cidx < H; ++cidx) fin[(size_t)l * Hp + cidx] = l * H + cidx;
    rc = pack_conv1d(h, fused, "post_fused", post_out, NL * H, 1, 1, post_map, fin, true, 0, post);
    if (rc) return rc;
    post->macs_per_row = (double)post_out * H;  // algorithmic MACs of the reference projection
  }
  if (gin) {
    snprintf(pfx, sizeof(pfx), "%s.cond_layer", enc.c_str());
    const int64_t ws[3] = {2 * H * NL, gin, 1}, bs[1] = {2 * H * NL};
    const HostTensor* w = find_tensor(h, m, std::string(pfx) + ".weight", 3, ws);
    const HostTensor* b = w ? find_tensor(h, m, std::string(pfx) + ".bias", 1, bs) : nullptr;
    if (!w || !b) return MBV_ERR_WEIGHTS;
    std::vector<float> pw((size_t)NL * 2 * Hp * gin, 0.f), pb((size_t)NL * 2 * Hp, 0.f);
    for (int l = 0; l < NL; ++l)
      for (int o = 0; o < 2 * H; ++o) {
        const int ch = o < H ? o : o - H;  // same row order as the packed in_layers weights
        const int dst = l * 2 * Hp + 128 * (ch / 64) + (ch % 64) + (o < H ? 0 : 64);
        const int src = l * 2 * H + o;
        memcpy(&pw[(size_t)dst * gin], &w->data[(size_t)src * gin], sizeof(float) * gin);
        pb[dst] = b->data[src];
      }
    if ((rc = upload_f32(h, pw, cond_w))) return rc;
    if ((rc = upload_f32(h, pb, cond_b))) return rc;
  }
  return MBV_OK;
}

}  // namespace

// =================================================================================================
// ABI
// =================================================================================================
extern "C" int mbv_abi_version(void) { return MBV_ABI_VERSION; }

extern "C" const char* mbv_last_error(mbv_handle* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int mbv_create(const mbv_config* cfg, mbv_handle** out) {
  if (!cfg || !out) return MBV_ERR_INVALID;
  *out = nullptr;
  mbv_handle* h = new mbv_handle();
  h->cfg = *cfg;
  *out = h;  // returned even on failure so the caller can read the message; destroy it either way
  const mbv_config& c = h->cfg;
  if (c.variant < 0 || c.variant > 2) return fail(h, MBV_ERR_INVALID, "variant must be 0 (istft), 1 (mb) or 2 (ms)");
  if (c.precision < 0 || c.precision > 3) return fail(h, MBV_ERR_INVALID, "precision must be 0 (fp32), 1 (tf32), 2 (bf16) or 3 (fp16)");
  h->prec = c.precision;
  h->esize = c.precision >= MBV_PREC_BF16 ? 2 : 4;
  if ((c.flags & MBV_FLAG_RESIDUAL_FP16) || c.precision == MBV_PREC_FP16) {
    if (c.precision < MBV_PREC_BF16) return fail(h, MBV_ERR_INVALID, "MBV_FLAG_RESIDUAL_FP16 requires a 16-bit precision (the fp32/tf32 paths keep an fp32 residual stream)");
    h->res_half = 1;
    h->rsize = 2;
    h->single = (c.precision == MBV_PREC_FP16 && !(c.flags & MBV_FLAG_FORCE_SIMT)) ? 1 : 0;
  }
  if (c.inter_channels <= 0 || c.inter_channels % 64 != 0 || (c.inter_channels / 2) % 16 != 0)
    return fail(h, MBV_ERR_UNSUPPORTED, "inter_channels must be a multiple of 64 (got %d)", c.inter_channels);
  if (c.hidden_channels <= 0 || c.hidden_channels % 16 != 0)
    return fail(h, MBV_ERR_UNSUPPORTED, "hidden_channels must be a multiple of 16 (got %d)", c.hidden_channels);
  if (c.n_ups < 1 || c.n_ups > MBV_MAX_UPS) return fail(h, MBV_ERR_UNSUPPORTED, "n_ups must be 1..%d", MBV_MAX_UPS);
  if (c.n_kernels < 1 || c.n_kernels > 3) return fail(h, MBV_ERR_UNSUPPORTED, "1..3 resblock kernels supported (got %d)", c.n_kernels);
  if (c.resblock_type != 1 && c.resblock_type != 2) return fail(h, MBV_ERR_INVALID, "resblock_type must be 1 or 2");
  if (c.n_dilations < 1 || c.n_dilations > MBV_MAX_DILATIONS) return fail(h, MBV_ERR_UNSUPPORTED, "1..3 dilations per resblock supported");
  if (c.n_fft != 16 || c.hop != 4) return fail(h, MBV_ERR_UNSUPPORTED, "the fused tail implements gen_istft_n_fft=16 / hop=4 only (got %d/%d)", c.n_fft, c.hop);
  if (c.variant == MBV_VARIANT_ISTFT ? c.subbands != 1 : c.subbands != 4)
    return fail(h, MBV_ERR_UNSUPPORTED, "subbands must be 4 for mb/ms and 1 for istft (got %d)", c.subbands);
  if (c.flow_kernel % 2 != 1 || c.flow_layers < 1 || c.flow_layers > 4 || c.flow_n != 4 || c.flow_dilation_rate != 1)
    return fail(h, MBV_ERR_UNSUPPORTED, "flow geometry: odd kernel, 1..4 layers, dilation_rate 1, n_flows 4");
  if (c.gin_channels < 0) return fail(h, MBV_ERR_INVALID, "gin_channels < 0");
  h->Cz = c.inter_channels;
  h->H = c.hidden_channels;
  h->Hp = round_up(c.hidden_channels, 64);
  h->Hw = round_up(c.hidden_channels, 128);
  h->n_stage = c.n_ups;
  int ch = c.upsample_initial_channel;
  if (ch % 64 != 0 || ch > 1024) return fail(h, MBV_ERR_UNSUPPORTED, "upsample_initial_channel must be a multiple of 64 and <= 1024");
  int up = 1;
  for (int i = 0; i < c.n_ups; ++i) {
    if (ch % 2) return fail(h, MBV_ERR_UNSUPPORTED, "channel halving hit an odd width");
    ch /= 2;
    if (ch % 64 != 0) return fail(h, MBV_ERR_UNSUPPORTED, "stage %d width %d is not a multiple of 64", i, ch);
    h->stage_C[i] = ch;
    up *= c.upsample_rates[i];
    for (int j = 0; j < c.n_kernels; ++j) {
      const int k = c.resblock_kernel_sizes[j];
      if (k % 2 != 1) return fail(h, MBV_ERR_UNSUPPORTED, "resblock kernel sizes must be odd");
      for (int p = 0; p < c.n_dilations; ++p)
        if ((k - 1) * c.resblock_dilations[j][p] > 128)
          return fail(h, MBV_ERR_UNSUPPORTED, "resblock receptive field (k-1)*d = %d exceeds the 128-row halo", (k - 1) * c.resblock_dilations[j][p]);
    }
  }
  h->n_logit = c.subbands * (c.n_fft + 2);
  h->spf = up * c.hop * c.subbands;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    h->err = "no CUDA device: handle created for inspection only, compute entries will fail";
    return MBV_OK;  // geometry validated; compute entries fail loudly later
  }
  if (c.device < 0 || c.device >= ndev) return fail(h, MBV_ERR_INVALID, "device %d out of range", c.device);
  DEVICE_GUARD(h);
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, c.device));
  h->num_sms = prop.multiProcessorCount;
  if (prop.major != 10) return fail(h, MBV_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", c.device, prop.major, prop.minor);
  if (h->prec != MBV_PREC_FP32) CUDA_TRY(h, tc_set_attributes());
  if (h->prec >= MBV_PREC_BF16) CUDA_TRY(h, tc_pair_set_attributes());
  h->use_branches = (c.flags & MBV_FLAG_BRANCHES) ? 1 : 0;
  if (h->use_branches && h->prec != MBV_PREC_FP32) {
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(h, cudaStreamCreateWithFlags(&h->br_stream[i], cudaStreamNonBlocking));
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
    }
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  } else {
    h->use_branches = 0;
  }
  h->use_pair = (c.flags & MBV_FLAG_FUSED_PAIR) ? 1 : 0;
  if (const char* e = getenv("MBV_PAIR_MAX_TAPS")) h->pair_max_taps = atoi(e);  // A/B measurements only
  return MBV_OK;
}

extern "C" void mbv_destroy(mbv_handle* h) {
  if (!h) return;
  for (void* p : h->dev_allocs) cudaFree(p);
  for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->br_stream[i]) cudaStreamDestroy(h->br_stream[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  delete h;
}

extern "C" int mbv_load_weights(mbv_handle* h, const mbv_tensor* tensors, int32_t n) {
  if (!h || !tensors || n <= 0) return MBV_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(h, MBV_ERR_CUDA, "mbv_load_weights: no CUDA device (there is no CPU fallback)");
  }
  DEVICE_GUARD(h);
  const mbv_config& c = h->cfg;
  TensorMap m;
  for (int i = 0; i < n; ++i) {
    const mbv_tensor& t = tensors[i];
    if (!t.name || !t.data || t.rank < 1 || t.rank > 4) return fail(h, MBV_ERR_WEIGHTS, "tensor %d is malformed", i);
    HostTensor ht;
    size_t cnt = 1;
    for (int r = 0; r < t.rank; ++r) { ht.shape.push_back(t.shape[r]); cnt *= (size_t)t.shape[r]; }
    ht.data.assign(t.data, t.data + cnt);
    m[t.name] = std::move(ht);
  }
  int rc;
  const int gin = c.gin_channels;
  // ---------------- decoder
  {
    const int C0 = c.upsample_initial_channel;
    rc = pack_conv1d(h, m, "dec.conv_pre", C0, h->Cz, 7, 1, iota_pad(C0, C0), iota_pad(h->Cz, h->Cz), true, 0, &h->conv_pre);
    if (rc) return rc;
    int cin = C0;
    for (int i = 0; i < c.n_ups; ++i) {
      const int C = h->stage_C[i];
      char name[64];
      snprintf(name, sizeof(name), "dec.ups.%d", i);
      rc = pack_convT(h, m, name, cin, C, c.upsample_kernel_sizes[i], c.upsample_rates[i], &h->ups[i]);
      if (rc) return rc;
      for (int j = 0; j < c.n_kernels; ++j) {
        const int k = c.resblock_kernel_sizes[j];
        char pfx[96];
        for (int p = 0; p < c.n_dilations; ++p) {
          const int d = c.resblock_dilations[j][p];
          if (c.resblock_type == 1) {
            snprintf(pfx, sizeof(pfx), "dec.resblocks.%d.convs1.%d", i * c.n_kernels + j, p);
            rc = pack_conv1d(h, m, pfx, C, C, k, d, iota_pad(C, C), iota_pad(C, C), true, 0, &h->rb_c1[i][j][p]);
            if (rc) return rc;
            snprintf(pfx, sizeof(pfx), "dec.resblocks.%d.convs2.%d", i * c.n_kernels + j, p);
            rc = pack_conv1d(h, m, pfx, C, C, k, 1, iota_pad(C, C), iota_pad(C, C), true, 0, &h->rb_c2[i][j][p]);
            if (rc) return rc;
          } else {
            snprintf(pfx, sizeof(pfx), "dec.resblocks.%d.convs.%d", i * c.n_kernels + j, p);
            rc = pack_conv1d(h, m, pfx, C, C, k, d, iota_pad(C, C), iota_pad(C, C), true, 0, &h->rb_c1[i][j][p]);
            if (rc) return rc;
          }
        }
        if (gin) {
          snprintf(pfx, sizeof(pfx), "dec.resblocks.%d.cond", i * c.n_kernels + j);
          const int64_t ws[3] = {C, gin, 1}, bs[1] = {C};
          const HostTensor* w = find_tensor(h, m, std::string(pfx) + ".weight", 3, ws);
          const HostTensor* b = w ? find_tensor(h, m, std::string(pfx) + ".bias", 1, bs) : nullptr;
          if (!w || !b) return MBV_ERR_WEIGHTS;
          if ((rc = upload_f32(h, w->data, &h->rb_cond_w[i][j]))) return rc;
          if ((rc = upload_f32(h, b->data, &h->rb_cond_b[i][j]))) return rc;
        }
      }
      cin = C;
    }
    const char* post = c.variant == MBV_VARIANT_ISTFT ? "dec.conv_post" : "dec.subband_conv_post";
    rc = pack_conv1d(h, m, post, h->n_logit, cin, 7, 1, iota_pad(h->n_logit, h->n_logit), iota_pad(cin, cin), true, 0, &h->conv_post);
    if (rc) return rc;
    {  // host copy of the bias: the fused conv_post + tail kernel takes it by value (constant-bank operands)
      const HostTensor& pb = m[std::string(post) + ".bias"];
      memset(h->post_bias, 0, sizeof(h->post_bias));
      for (int i = 0; i < h->n_logit && i < 72; ++i) h->post_bias[i] = pb.data[i];
    }
    float hs[4][63];
    memset(hs, 0, sizeof(hs));
    if (c.variant == MBV_VARIANT_MB) {
      // fast path tables: cos(theta_c(k)) = (-1)^floor(k/8) cos(theta_c(k mod 8)) (pqmf.py:72-75)
      double proto[63];
      design_pqmf_synthesis(hs, proto);
      const double pi = 3.14159265358979323846;
      for (int mI = 0; mI < 8; ++mI)
        for (int cc = 0; cc < 4; ++cc)
          h->tail_mod[mI][cc] = (float)(2.0 * cos((2 * cc + 1) * (pi / 8.0) * (mI - 30.5) - ((cc & 1) ? -1.0 : 1.0) * pi / 4.0));
      for (int r = 0; r < 4; ++r)
        for (int d = -7; d <= 8; ++d) {
          const int k = 4 * d + 31 - r;
          h->tail_g2[r][d + 7] = (k >= 0 && k <= 62) ? (float)(4.0 * proto[k] * (((k / 8) & 1) ? -1.0 : 1.0)) : 0.f;
        }
      h->tail_fast = 1;
    } else if (c.variant == MBV_VARIANT_MS) {
      const int64_t ws[3] = {1, 4, 63};
      const HostTensor* w = find_tensor(h, m, "dec.multistream_conv_post.weight", 3, ws);
      if (!w) return MBV_ERR_WEIGHTS;
      for (int ch4 = 0; ch4 < 4; ++ch4)
        for (int k = 0; k < 63; ++k) hs[ch4][k] = w->data[ch4 * 63 + k];
    }
    fill_tail_coef(h, hs);
  }
  // ---------------- flow (Flips folded into pre/post)
  {
    const int half = h->Cz / 2, H = h->H, Hp = h->Hp, K = c.flow_kernel, NL = c.flow_layers;
    for (int f = 0; f < 4; ++f) {
      // coupling layer f runs after (4 - f) flips: odd -> reads rev(z[half:]) and updates rev(z[:half])
      const bool odd = ((4 - f) & 1) != 0;
      char pfx[96];
      std::vector<int> in_map(h->Cz, -1), out_map(Hp, -1);
      for (int j = 0; j < h->Cz; ++j) {
        if (!odd && j < half) in_map[j] = j;
        if (odd && j >= half) in_map[j] = h->Cz - 1 - j;
      }
      for (int o = 0; o < H; ++o) out_map[o] = o;
      snprintf(pfx, sizeof(pfx), "flow.flows.%d.pre", 2 * f);
      rc = pack_conv1d(h, m, pfx, H, half, 1, 1, out_map, in_map, true, 0, &h->fl_pre[f]);
      if (rc) return rc;
      std::vector<int> pmap(half, -1);
      for (int j2 = 0; j2 < half; ++j2) pmap[j2] = odd ? half - 1 - j2 : j2;
      snprintf(pfx, sizeof(pfx), "flow.flows.%d.enc", 2 * f);
      char post_name[96];
      snprintf(post_name, sizeof(post_name), "flow.flows.%d.post", 2 * f);
      rc = pack_wn(h, m, pfx, NL, K, post_name, half, pmap, h->fl_in[f], h->fl_rs[f], &h->fl_post[f], &h->fl_cond_w[f], &h->fl_cond_b[f]);
      if (rc) return rc;
    }
  }
  // ---------------- posterior encoder (optional)
  if (m.count("enc_q.pre.weight")) {
    const HostTensor& pw = m["enc_q.pre.weight"];
    if (pw.shape.size() != 3 || pw.shape[0] != h->H || pw.shape[2] != 1) return fail(h, MBV_ERR_WEIGHTS, "enc_q.pre.weight has the wrong shape");
    int NL = 0;
    char pfx[96];
    for (;; ++NL) { snprintf(pfx, sizeof(pfx), "enc_q.enc.in_layers.%d.weight", NL); if (!m.count(pfx)) break; }
    if (NL < 1 || NL > MBV_MAX_ENCQ_LAYERS) return fail(h, MBV_ERR_UNSUPPORTED, "enc_q: %d WN layers (1..%d supported)", NL, MBV_MAX_ENCQ_LAYERS);
    h->eq_layers = NL;
    h->eq_spec = (int)pw.shape[1];
    h->eq_spec_p = round_up(h->eq_spec, 64);
    rc = pack_conv1d(h, m, "enc_q.pre", h->H, h->eq_spec, 1, 1, iota_pad(h->H, h->Hp), iota_pad(h->eq_spec, h->eq_spec_p), true, 0, &h->eq_pre);
    if (rc) return rc;
    // proj rows [m (inter) | logs (inter)] in their own order (models.py:243-244)
    rc = pack_wn(h, m, "enc_q.enc", NL, c.flow_kernel, "enc_q.proj", 2 * h->Cz, iota_pad(2 * h->Cz, 2 * h->Cz), h->eq_in, h->eq_rs, &h->eq_proj,
                 &h->eq_cond_w, &h->eq_cond_b);
    if (rc) return rc;
    h->has_enc_q = true;
  }
  // ---------------- text encoder (optional)
  if (m.count("enc_p.emb.weight")) {
    const HostTensor& ew = m["enc_p.emb.weight"];
    if (ew.shape.size() != 2 || ew.shape[1] != h->H) return fail(h, MBV_ERR_WEIGHTS, "enc_p.emb.weight has the wrong shape");
    const int H = h->H, Hp = h->Hp;
    h->tp_vocab = (int)ew.shape[0];
    if ((rc = upload_f32(h, ew.data, &h->tp_emb))) return rc;
    int NL = 0;
    char pfx[128];
    for (;; ++NL) { snprintf(pfx, sizeof(pfx), "enc_p.encoder.attn_layers.%d.conv_q.weight", NL); if (!m.count(pfx)) break; }
    if (NL < 1 || NL > MBV_MAX_TEXT_LAYERS) return fail(h, MBV_ERR_UNSUPPORTED, "enc_p: %d encoder layers (1..%d supported)", NL, MBV_MAX_TEXT_LAYERS);
    h->tp_layers = NL;
    for (int l = 0; l < NL; ++l) {
      snprintf(pfx, sizeof(pfx), "enc_p.encoder.attn_layers.%d", l);
      const std::string a(pfx);
      auto rk = m.find(a + ".emb_rel_k"), rv = m.find(a + ".emb_rel_v");
      if (rk == m.end() || rv == m.end() || rk->second.shape.size() != 3 || rk->second.shape[0] != 1 || rk->second.shape != rv->second.shape)
        return fail(h, MBV_ERR_UNSUPPORTED, "%s: shared relative embeddings [1, 2w+1, dk] expected", pfx);
      const int dk = (int)rk->second.shape[2], win = ((int)rk->second.shape[1] - 1) / 2;
      if (dk < 4 || dk > 128 || (dk & 3) || H % dk != 0) return fail(h, MBV_ERR_UNSUPPORTED, "%s: head width %d not supported", pfx, dk);
      if (l == 0) { h->tp_heads = H / dk; h->tp_window = win; }
      else if (h->tp_heads != H / dk || h->tp_window != win) return fail(h, MBV_ERR_WEIGHTS, "%s: inconsistent attention geometry", pfx);
      if ((rc = upload_f32(h, rk->second.data, &h->tp_relk[l]))) return rc;
      if ((rc = upload_f32(h, rv->second.data, &h->tp_relv[l]))) return rc;
      // q | k | v as ONE 1x1 conv with 3H output rows
      HostTensor qw, qb;
      qw.shape = {3 * H, H, 1};
      qb.shape = {3 * H};
      const char* names[3] = {".conv_q", ".conv_k", ".conv_v"};
      for (int j = 0; j < 3; ++j) {
        const int64_t ws[3] = {H, H, 1}, bs[1] = {H};
        const HostTensor* w = find_tensor(h, m, a + names[j] + ".weight", 3, ws);
        const HostTensor* b = w ? find_tensor(h, m, a + names[j] + ".bias", 1, bs) : nullptr;
        if (!w || !b) return MBV_ERR_WEIGHTS;
        qw.data.insert(qw.data.end(), w->data.begin(), w->data.end());
        qb.data.insert(qb.data.end(), b->data.begin(), b->data.end());
      }
      TensorMap fused;
      fused["qkv.weight"] = std::move(qw);
      fused["qkv.bias"] = std::move(qb);
      rc = pack_conv1d(h, fused, "qkv", 3 * H, H, 1, 1, iota_pad(3 * H, round_up(3 * H, 128)), iota_pad(H, Hp), true, 0, &h->tp_qkv[l]);
      if (rc) return rc;
      rc = pack_conv1d(h, m, a + ".conv_o", H, H, 1, 1, iota_pad(H, round_up(H, 128)), iota_pad(H, Hp), true, 0, &h->tp_o[l]);
      if (rc) return rc;
      snprintf(pfx, sizeof(pfx), "enc_p.encoder.ffn_layers.%d.conv_1.weight", l);
      auto f1 = m.find(pfx);
      if (f1 == m.end() || f1->second.shape.size() != 3) return fail(h, MBV_ERR_WEIGHTS, "missing tensor %s", pfx);
      const int F = (int)f1->second.shape[0], ks = (int)f1->second.shape[2];
      if (F % 64 != 0 || (ks & 1) == 0) return fail(h, MBV_ERR_UNSUPPORTED, "enc_p FFN: filter_channels %d must be a multiple of 64, kernel %d odd", F, ks);
      h->tp_filter = F;
      snprintf(pfx, sizeof(pfx), "enc_p.encoder.ffn_layers.%d", l);
      rc = pack_conv1d(h, m, std::string(pfx) + ".conv_1", F, H, ks, 1, iota_pad(F, round_up(F, 128)), iota_pad(H, Hp), true, 0, &h->tp_f1[l]);
      if (rc) return rc;
      rc = pack_conv1d(h, m, std::string(pfx) + ".conv_2", H, F, ks, 1, iota_pad(H, round_up(H, 128)), iota_pad(F, F), true, 0, &h->tp_f2[l]);
      if (rc) return rc;
      const char* ln[4] = {"norm_layers_1.%d.gamma", "norm_layers_1.%d.beta", "norm_layers_2.%d.gamma", "norm_layers_2.%d.beta"};
      for (int j = 0; j < 4; ++j) {
        char nm[128];
        snprintf(nm, sizeof(nm), ln[j], l);
        const int64_t vs[1] = {H};
        const HostTensor* v = find_tensor(h, m, std::string("enc_p.encoder.") + nm, 1, vs);
        if (!v) return MBV_ERR_WEIGHTS;
        if ((rc = upload_f32(h, v->data, &h->tp_ln[l][j]))) return rc;
      }
    }
    rc = pack_conv1d(h, m, "enc_p.proj", 2 * h->Cz, H, 1, 1, iota_pad(2 * h->Cz, round_up(2 * h->Cz, 128)), iota_pad(H, Hp), true, 0, &h->tp_proj);
    if (rc) return rc;
    h->has_enc_p = true;
  }
  h->weights_loaded = true;
  return MBV_OK;
}

// =================================================================================================
// workspace planning
// =================================================================================================
namespace {

struct Arena {
  uint8_t* base;
  size_t off = 0;
  explicit Arena(void* p) : base((uint8_t*)p) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

struct DecBufs {
  void* zin_op; void* pre_act;
  struct Stage { void* x; void* a[3]; void* xr[3]; void* ar[3]; void* hop[3]; void* xs; void* next; } st[MBV_MAX_UPS];
  float* logits;
  float* cond;  // [n_stage][n_kernels][2][B][C]: (cond, bias2+cond) per resblock
};
struct FlowBufs {
  float* z; void* zop; float* h; void* hop; void* acts; float* gcond;  // acts [B][T][NL*Hp]; gcond [4][B][NL*2Hp]
};

void layout_dec(mbv_handle* h, Arena& A, int B, int T, DecBufs* d) {
  const int es = h->esize;
  d->zin_op = A.take((size_t)B * T * h->Cz * es);
  d->pre_act = A.take((size_t)B * T * h->cfg.upsample_initial_channel * es);
  int L = T;
  for (int i = 0; i < h->n_stage; ++i) {
    L *= h->cfg.upsample_rates[i];
    const int C = h->stage_C[i];
    const size_t n = (size_t)B * L * C;
    auto& s = d->st[i];
    s.x = A.take(n * h->rsize);
    const int na = h->cfg.gin_channels ? h->cfg.n_kernels : 1;
    for (int j = 0; j < 3; ++j) s.a[j] = j < na ? A.take(n * es) : nullptr;
    const int nb = (h->use_branches && h->cfg.n_kernels > 1) ? h->cfg.n_kernels : 1;  // one buffer set per ResBlock branch
    for (int j = 0; j < 3; ++j) {
      s.xr[j] = j < nb ? A.take(n * h->rsize) : nullptr;
      s.ar[j] = j < nb ? A.take(n * es) : nullptr;
      s.hop[j] = j < nb ? A.take(n * es) : nullptr;
    }
    s.xs = A.take(n * h->rsize);
    const bool last = (i == h->n_stage - 1);
    s.next = A.take((size_t)B * (L + (last ? 1 : 0)) * C * es);
  }
  d->logits = (float*)A.take((size_t)B * (L + 1) * h->n_logit * 4);
  d->cond = (float*)A.take((size_t)h->n_stage * h->cfg.n_kernels * 2 * B * 512 * 4);
}

void layout_flow(mbv_handle* h, Arena& A, int B, int T, FlowBufs* f) {
  const int es = h->esize;
  const size_t nz = (size_t)B * T * h->Cz, nh = (size_t)B * T * h->Hp;
  f->z = (float*)A.take(nz * 4);
  f->zop = A.take(nz * es);
  f->h = (float*)A.take(nh * 4);
  f->hop = A.take(nh * es);
  f->acts = A.take(nh * h->cfg.flow_layers * es);
  f->gcond = (float*)A.take((size_t)4 * B * h->cfg.flow_layers * 2 * h->Hp * 4);
}

// Launch context: counts launches, caches tensor maps
struct Ctx {
  mbv_handle* h;
  cudaStream_t st;
  std::vector<TcPlan>* plans;
  std::vector<TcPairPlan>* pair_plans = nullptr;
  bool plans_valid;
  size_t plan_idx = 0, pair_idx = 0;
  int launches = 0;
  bool pdl = true;   // launch convs with programmatic stream serialization (off inside ResBlock branches)
};

// RAII bracket: two events around one launch when profiling is on
struct ProfScope {
  mbv_handle* h; cudaStream_t st; int kind; cudaEvent_t e0; char desc[56];
  ProfScope(Ctx& cx, int kind_, const char* d = "") : h(cx.h), st(cx.st), kind(kind_), e0(nullptr) {
    snprintf(desc, sizeof(desc), "%s", d);
    if (h->profiling) { e0 = h->get_event(); cudaEventRecord(e0, st); }
  }
  ~ProfScope() {
    if (e0) {
      cudaEvent_t e1 = h->get_event();
      cudaEventRecord(e1, st);
      mbv_handle::ProfRec r;
      r.kind = kind; r.e0 = e0; r.e1 = e1;
      memcpy(r.desc, desc, sizeof(desc));
      h->prof.push_back(r);
    }
  }
};

int run_conv(Ctx& cx, const ConvLayer& L, const void* x, int B, int L_in, int L_out, const EpiParams& epi, int x_ld = 0) {
  mbv_handle* h = cx.h;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.x_ld = x_ld > 0 ? x_ld : L.Cp_in; a.w = L.w; a.B = B; a.L_in = L_in; a.L_out = L_out; a.Cp_in = L.Cp_in; a.N_total = L.N_total;
  a.taps = L.taps; a.dil = L.dil; a.n_phases = L.n_phases; a.gate = L.gate;
  for (int i = 0; i < kMaxPhases; ++i) a.shift0[i] = L.shift0[i];
  a.epi = epi;
  if (a.epi.bias == nullptr) { a.epi.bias = L.bias; a.epi.bias_bs = 0; }
  if (h->prec == MBV_PREC_FP32 || (h->cfg.flags & MBV_FLAG_FORCE_SIMT)) {
    ProfScope prof(cx, 0, "conv_simt");
    CUDA_TRY(h, launch_conv_simt(h->prec, a, cx.st));
  } else {
    TcPlan plan;
    if (cx.plans_valid && cx.plan_idx < cx.plans->size()) {
      plan = (*cx.plans)[cx.plan_idx];
    } else {
      const char* msg = tc_make_plan(h->prec, a, h->cfg.flags, h->num_sms, &plan);
      if (msg) return fail(h, MBV_ERR_CUDA, "%s", msg);
      cx.plans->push_back(plan);
    }
    cx.plan_idx++;
    char desc[56];
    snprintf(desc, sizeof(desc), "conv m%d Ci%d N%d k%d d%d ph%d L%d nt%d", epi.mode, L.Cp_in, L.N_total, L.taps, L.dil,
             L.n_phases, L_out, plan.n_time);
    ProfScope prof(cx, 0, h->profiling ? desc : "");
    CUDA_TRY(h, launch_conv_tc(h->prec, a, plan, cx.st, cx.pdl ? 1 : 0));
  }
  cx.launches++;
  return MBV_OK;
}

// fused ResBlock1 conv pair: x' = xin + c2(lrelu(c1(a_in))) (conv_pair_kernel); epi is c2's RES epilogue
bool pair_ok(const mbv_handle* h, const ConvLayer& c1, const ConvLayer& c2) {
  return h->use_pair && h->prec >= MBV_PREC_BF16 && !(h->cfg.flags & MBV_FLAG_FORCE_SIMT) && c1.Cp_in == 128 && c1.N_total == 128 &&
         c2.Cp_in == 128 && c2.N_total == 128 && c1.taps == c2.taps && c2.dil == 1 && (c1.taps & 1) && (c1.taps - 1) / 2 <= 8 &&
         c1.taps <= h->pair_max_taps && c1.n_phases == 1 && c2.n_phases == 1;
}

// k = 3 conv pairs of a 128-channel stage whose residual add is plain (sum_mode 0) or starts the running ResBlock sum (1):
// pair_tm_kernel (default on the bf16 path)
bool pairtm_ok(const mbv_handle* h, const ConvLayer& c1, const ConvLayer& c2, int sum_mode) {
  return (sum_mode == 0 || sum_mode == 1) && h->prec == MBV_PREC_BF16 && h->res_half && !h->single && h->num_sms >= 2 &&
         !(h->cfg.flags & (MBV_FLAG_FORCE_SIMT | MBV_FLAG_NO_PW | MBV_FLAG_NO_PAIR_TM)) && c1.Cp_in == 128 && c1.N_total == 128 &&
         c2.Cp_in == 128 && c2.N_total == 128 && c1.taps == 3 && c2.taps == 3 && c2.dil == 1 && c1.n_phases == 1 && c2.n_phases == 1;
}

int run_pair(Ctx& cx, const ConvLayer& c1, const ConvLayer& c2, const void* x, int B, int L, const EpiParams& epi, float slope_h) {
  mbv_handle* h = cx.h;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.x_ld = c1.Cp_in; a.w = c1.w; a.w2 = c2.w; a.B = B; a.L_in = L; a.L_out = L; a.Cp_in = c1.Cp_in; a.N_total = c1.N_total;
  a.taps = c1.taps; a.dil = c1.dil; a.n_phases = 1; a.gate = 0;
  a.shift0[0] = c1.shift0[0];
  a.bias_h = c1.bias; a.slope_h = slope_h;
  a.epi = epi;
  if (a.epi.bias == nullptr) { a.epi.bias = c2.bias; a.epi.bias_bs = 0; }
  TcPairPlan plan;
  if (cx.plans_valid && cx.pair_idx < cx.pair_plans->size()) {
    plan = (*cx.pair_plans)[cx.pair_idx];
  } else {
    memset(&plan, 0, sizeof(plan));
    // k = 3 pairs: pair_tm_kernel (time on the accumulator lane, CTA pairs, resident weights); else the experimental conv_pair_kernel
    const bool tm = ptm_eligible(h->prec, a, h->cfg.flags, h->num_sms);
    const char* msg = tm ? ptm_make_plan(h->prec, a, h->num_sms, &plan) : tc_make_pair_plan(h->prec, a, h->num_sms, &plan);
    if (msg) return fail(h, MBV_ERR_CUDA, "%s", msg);
    cx.pair_plans->push_back(plan);
  }
  cx.pair_idx++;
  char desc[56];
  snprintf(desc, sizeof(desc), "pair m1 Ci%d N%d k%d d%d ph1 L%d nt%d", c1.Cp_in, c1.N_total, c1.taps, c1.dil, L, plan.tm ? plan.tm_out_rows : 160);
  ProfScope prof(cx, 0, h->profiling ? desc : "");
  if (plan.tm) CUDA_TRY(h, launch_ptm(h->prec, a, plan, cx.st, cx.pdl ? 1 : 0));
  else CUDA_TRY(h, launch_conv_pair(h->prec, a, plan, cx.st, cx.pdl ? 1 : 0));
  cx.launches++;
  return MBV_OK;
}

// ld = channel pitch of the destination buffers; by default every channel of that pitch is written (pad channels
// come out as exact zeros because their packed weight rows and biases are zero)
EpiParams epi_base(int mode, int ld, int rows) {
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.mode = mode; e.ld = ld; e.rows_out = rows; e.rows_res = rows; e.row_mul = 1; e.row_add = 0;
  e.dup_src = -1; e.dup_dst = 0; e.slope = 1.f; e.scale = 1.f; e.n_valid = ld; e.post_sign = 1.f;
  return e;
}

// modules.WN.forward (modules.py:148-176) on h (fp32 / fp16 stream `hres` + operand copy `hop`): per layer the gate conv
// writes acts_l = tanh(.) * sigmoid(.) into slot l of the gate-output buffer `acts` [B][T][NL*Hp] and, for l < NL-1, the
// residual 1x1 conv updates h = (h + res_l(acts_l)) * mask.  The skip path is applied by the caller's fused projection
// over `acts` (pack_wn).  gc: per-utterance cond_layer(g) rows [B][NL*2Hp] in the packed gate row order, or null.
int run_wn(Ctx& cx, const ConvLayer* in_layers, const ConvLayer* rs_layers, int NL, float* hres, void* hop, void* acts,
           const float* mask, const float* gc, int B, int T) {
  mbv_handle* h = cx.h;
  const int Hp = h->Hp;
  const bool tc_path = h->prec >= MBV_PREC_BF16 && !(h->cfg.flags & MBV_FLAG_FORCE_SIMT);  // (the CUDA-core epilogues keep h in fp32)
  const int Ha = NL * Hp;  // channel pitch of the gate-output buffer: one Hp-wide slot per WN layer
  int rc;
  for (int l = 0; l < NL; ++l) {
    {  // acts_l = tanh(.) * sigmoid(.) of in_layer_l(h) [+ cond_l(g)]
      EpiParams e = epi_base(EPI_GATE, Ha, T);
      e.n_valid = Hp; e.ch_off = l * Hp;
      e.act[0] = acts; e.n_act = 1;
      if (gc) { e.add2 = gc + (size_t)l * 2 * Hp; e.add2_bs = NL * 2 * Hp; }
      if ((rc = run_conv(cx, in_layers[l], hop, B, T, T, e))) return rc;
    }
    if (l < NL - 1) {  // h = (h + res_l(acts_l)) * mask; the skip halves are applied by the fused projection
      EpiParams e = epi_base(EPI_RS, Hp, T);
      e.mask = mask; e.n_split = rs_layers[l].N_total; e.xin = hres; e.xout = hres; e.act[0] = hop; e.n_act = 1;
      if (h->single) { e.xin = hop; e.xout = nullptr; e.res_half = 2; e.inv_slope = 1.f; }
      else if (h->res_half && tc_path) e.res_half = 1;
      const char* ax = (const char*)acts + (size_t)l * Hp * h->esize;
      if ((rc = run_conv(cx, rs_layers[l], ax, B, T, T, e, Ha))) return rc;
    }
  }
  return MBV_OK;
}

// forward = false: ResidualCouplingBlock.forward(reverse=True) (inference); true: the forward direction (voice conversion).
// With the Flips folded into the weights a coupling layer sees the same flip parity in both directions (4 - f and f flips
// before layer f), so the packed weights are shared: forward runs the layers in ascending order and adds m.
int run_flow(Ctx& cx, const FlowBufs& f, const float* z_p, const float* mask, const float* g, float* z_out, int B, int T,
             bool forward = false) {
  mbv_handle* h = cx.h;
  const mbv_config& c = h->cfg;
  const int Hp = h->Hp, NL = c.flow_layers;
  const bool tc_path = h->prec >= MBV_PREC_BF16 && !(c.flags & MBV_FLAG_FORCE_SIMT);  // (the CUDA-core epilogues keep h in fp32)
  {
    ProfScope prof(cx, 2);
    CUDA_TRY(h, launch_pack_input(h->prec, z_p, nullptr, f.zop, f.z, B, h->Cz, T, h->Cz, cx.st));
  }
  cx.launches++;
  for (int step = 0; step < 4; ++step) {
    const int f_i = forward ? step : 3 - step;
    float* gc = nullptr;
    if (g) {
      gc = f.gcond + (size_t)f_i * B * NL * 2 * Hp;
      ProfScope prof(cx, 2);
      CUDA_TRY(h, launch_cond_gemv(g, h->fl_cond_w[f_i], h->fl_cond_b[f_i], nullptr, gc, B, c.gin_channels, NL * 2 * Hp, NL * 2 * Hp, cx.st));
      cx.launches++;
    }
    int rc;
    {  // h = pre(x0) * mask
      EpiParams e = epi_base(EPI_ACT, Hp, T);
      e.mask = mask; e.xout = h->single ? nullptr : f.h; e.act[0] = f.hop; e.n_act = 1;  // single: the fp16 operand copy IS h
      if (!h->single && h->res_half && tc_path) e.res_half = 1;  // bf16 + fp16 streams: h is carried in fp16 next to its bf16 operand copy
      if ((rc = run_conv(cx, h->fl_pre[f_i], f.zop, B, T, T, e))) return rc;
    }
    if ((rc = run_wn(cx, h->fl_in[f_i], h->fl_rs[f_i], NL, f.h, f.hop, f.acts, mask, gc, B, T))) return rc;
    {  // x1 = (x1 - m * mask) * mask,  m = post(sum_l skip_l) as one conv over all gate outputs
      EpiParams e = epi_base(EPI_POST, h->Cz, T);
      e.mask = mask; e.xin = f.z; e.xout = f.z; e.act[0] = f.zop; e.n_act = 1;
      e.n_valid = h->Cz / 2;
      e.ch_off = ((4 - f_i) & 1) ? 0 : h->Cz / 2;
      e.post_sign = forward ? -1.f : 1.f;
      if ((rc = run_conv(cx, h->fl_post[f_i], f.acts, B, T, T, e))) return rc;
    }
  }
  if (z_out) {
    ProfScope prof(cx, 2);
    CUDA_TRY(h, launch_unpack_output(f.z, z_out, B, h->Cz, T, h->Cz, cx.st));
    cx.launches++;
  }
  return MBV_OK;
}

struct PostBufs { void* spec_op; float* h; void* hop; void* acts; float* gcond; float* stats; };

void layout_posterior(mbv_handle* h, Arena& A, int B, int T, PostBufs* p) {
  const int es = h->esize, NL = h->eq_layers;
  const size_t nh = (size_t)B * T * h->Hp;
  p->spec_op = A.take((size_t)B * T * h->eq_spec_p * es);
  p->h = (float*)A.take(nh * 4);
  p->hop = A.take(nh * es);
  p->acts = A.take(nh * NL * es);
  p->gcond = (float*)A.take((size_t)B * NL * 2 * h->Hp * 4);
  p->stats = (float*)A.take((size_t)B * T * 2 * h->Cz * 4);
}

// PosteriorEncoder.forward (models.py:236-246): x = pre(spec) * mask; x = WN(x, mask, g); stats = proj(x) * mask;
// m, logs = split(stats); z = (m + noise * exp(logs)) * mask.  stats_out is [B, 2*inter, T] (m | logs), NCT like z.
int run_posterior(Ctx& cx, const PostBufs& p, const float* spec, const float* mask, const float* g, const float* noise,
                  float* z_out, float* stats_out, int B, int T) {
  mbv_handle* h = cx.h;
  const int Hp = h->Hp, NL = h->eq_layers;
  const bool tc_path = h->prec >= MBV_PREC_BF16 && !(h->cfg.flags & MBV_FLAG_FORCE_SIMT);
  int rc;
  {
    ProfScope prof(cx, 2);
    CUDA_TRY(h, launch_pack_input(h->prec, spec, nullptr, p.spec_op, nullptr, B, h->eq_spec, T, h->eq_spec_p, cx.st));
    cx.launches++;
  }
  float* gc = nullptr;
  if (g) {
    gc = p.gcond;
    ProfScope prof(cx, 2);
    CUDA_TRY(h, launch_cond_gemv(g, h->eq_cond_w, h->eq_cond_b, nullptr, gc, B, h->cfg.gin_channels, NL * 2 * Hp, NL * 2 * Hp, cx.st));
    cx.launches++;
  }
  {  // x = pre(spec) * mask
    EpiParams e = epi_base(EPI_ACT, Hp, T);
    e.mask = mask; e.xout = h->single ? nullptr : p.h; e.act[0] = p.hop; e.n_act = 1;
    if (!h->single && h->res_half && tc_path) e.res_half = 1;
    if ((rc = run_conv(cx, h->eq_pre, p.spec_op, B, T, T, e))) return rc;
  }
  if ((rc = run_wn(cx, h->eq_in, h->eq_rs, NL, p.h, p.hop, p.acts, mask, gc, B, T))) return rc;
  {  // stats = proj(sum_l skip_l) * mask as one conv over all gate outputs, fp32 channels-last
    EpiParams e = epi_base(EPI_ACT, 2 * h->Cz, T);
    e.mask = mask; e.xout = p.stats; e.n_act = 0;
    if ((rc = run_conv(cx, h->eq_proj, p.acts, B, T, T, e))) return rc;
  }
  {
    ProfScope prof(cx, 2);
    CUDA_TRY(h, launch_unpack_output(p.stats, stats_out, B, 2 * h->Cz, T, 2 * h->Cz, cx.st));
    CUDA_TRY(h, launch_posterior_sample(stats_out, noise, mask, z_out, B, h->Cz, T, cx.st));
    cx.launches += 2;
  }
  return MBV_OK;
}

struct TextBufs { float* x; void* xop; float* qkv; void* att; float* y; void* hid; float* stats; };

void layout_text(mbv_handle* h, Arena& A, int B, int T, TextBufs* t) {
  const int es = h->esize;
  const size_t n = (size_t)B * T;
  t->x = (float*)A.take(n * h->H * 4);
  t->xop = A.take(n * h->Hp * es);
  t->qkv = (float*)A.take(n * 3 * h->H * 4);
  t->att = A.take(n * h->Hp * es);
  t->y = (float*)A.take(n * h->H * 4);
  t->hid = A.take(n * h->tp_filter * es);
  t->stats = (float*)A.take(n * 2 * h->Cz * 4);
}

// TextEncoder.forward (models.py:172-181) + attentions.Encoder.forward (attentions.py:35-47)
int run_text(Ctx& cx, const TextBufs& t, const long long* tokens, const float* mask, float* x_out, float* stats_out, int B, int T) {
  mbv_handle* h = cx.h;
  const int H = h->H, Hp = h->Hp, F = h->tp_filter, rows = B * T;
  int rc;
  {
    ProfScope prof(cx, 2);
    // pad channels [H, Hp) of the operand copies are never written below: keep them zero (their weight columns are zero,
    // but 0 x NaN is not)
    if (Hp != H) {
      CUDA_TRY(h, cudaMemsetAsync(t.xop, 0, (size_t)rows * Hp * h->esize, cx.st));
      CUDA_TRY(h, cudaMemsetAsync(t.att, 0, (size_t)rows * Hp * h->esize, cx.st));
    }
    CUDA_TRY(h, launch_text_embed(tokens, h->tp_emb, mask, t.x, t.xop, rows, H, Hp, h->tp_vocab, h->prec, cx.st));
    cx.launches++;
  }
  for (int l = 0; l < h->tp_layers; ++l) {
    {  // q | k | v = 1x1 convs of x -> fp32 [B][T][3H]
      EpiParams e = epi_base(EPI_ACT, 3 * H, T);
      e.xout = t.qkv; e.n_act = 0;
      if ((rc = run_conv(cx, h->tp_qkv[l], t.xop, B, T, T, e))) return rc;
    }
    {
      ProfScope prof(cx, 2);
      CUDA_TRY(h, launch_text_attention(t.qkv, mask, h->tp_relk[l], h->tp_relv[l], t.att, B, T, H, Hp, h->tp_heads, h->tp_window, h->prec, cx.st));
      cx.launches++;
    }
    {  // y = conv_o(attention)
      EpiParams e = epi_base(EPI_ACT, H, T);
      e.xout = t.y; e.n_act = 0;
      if ((rc = run_conv(cx, h->tp_o[l], t.att, B, T, T, e))) return rc;
    }
    {
      ProfScope prof(cx, 2);
      CUDA_TRY(h, launch_text_ln(t.x, t.y, H, h->tp_ln[l][0], h->tp_ln[l][1], mask, t.x, t.xop, rows, H, Hp, 0, h->prec, cx.st));
      cx.launches++;
    }
    {  // FFN: relu(conv_1(x * mask)) * mask -> operand
      EpiParams e = epi_base(EPI_ACT, F, T);
      e.mask = mask; e.slope = 0.f; e.act[0] = t.hid; e.n_act = 1;
      if ((rc = run_conv(cx, h->tp_f1[l], t.xop, B, T, T, e))) return rc;
    }
    {  // y = conv_2(.) * mask
      EpiParams e = epi_base(EPI_ACT, H, T);
      e.mask = mask; e.xout = t.y; e.n_act = 0;
      if ((rc = run_conv(cx, h->tp_f2[l], t.hid, B, T, T, e))) return rc;
    }
    {
      ProfScope prof(cx, 2);
      const int last = (l == h->tp_layers - 1) ? 1 : 0;  // x = x * x_mask after the last layer (attentions.py:46)
      CUDA_TRY(h, launch_text_ln(t.x, t.y, H, h->tp_ln[l][2], h->tp_ln[l][3], mask, t.x, t.xop, rows, H, Hp, last, h->prec, cx.st));
      cx.launches++;
    }
  }
  {  // stats = proj(x) * mask
    EpiParams e = epi_base(EPI_ACT, 2 * h->Cz, T);
    e.mask = mask; e.xout = t.stats; e.n_act = 0;
    if ((rc = run_conv(cx, h->tp_proj, t.xop, B, T, T, e))) return rc;
  }
  {
    ProfScope prof(cx, 2);
    CUDA_TRY(h, launch_unpack_output(t.x, x_out, B, H, T, H, cx.st));
    CUDA_TRY(h, launch_unpack_output(t.stats, stats_out, B, 2 * h->Cz, T, 2 * h->Cz, cx.st));
    cx.launches += 2;
  }
  return MBV_OK;
}

// conv_post inside the tail kernel (tail_fused_kernel): 16-bit tensor-core paths, the 4-band decoders (72 logit channels),
// conv_post k = 7 on 64 or 128 channels.  Everything else (fp32 / tf32 paths, single-band decoder) keeps conv_post as a
// conv launch that writes fp32 logits for the stand-alone tail kernel.
bool fused_tail_ok(const mbv_handle* h) {
  return h->prec >= MBV_PREC_BF16 && !(h->cfg.flags & (MBV_FLAG_FORCE_SIMT | MBV_FLAG_SPLIT_TAIL)) && h->cfg.variant != MBV_VARIANT_ISTFT &&
         h->n_logit == 72 && h->conv_post.taps == 7 && h->conv_post.dil == 1 && (h->conv_post.Cp_in == 64 || h->conv_post.Cp_in == 128);
}

void fill_tail_args(mbv_handle* h, TailArgs* ta, float* wav, float* o_mb, float* spec, float* phase, int B, int Lfr) {
  memset(ta, 0, sizeof(*ta));
  ta->wav = wav; ta->o_mb = o_mb; ta->spec = spec; ta->phase = phase;
  ta->B = B; ta->L = Lfr; ta->n_ch = h->n_logit; ta->variant = h->cfg.variant;
  memcpy(ta->coef, h->tail_coef, sizeof(ta->coef));
  memcpy(ta->mod, h->tail_mod, sizeof(ta->mod));
  memcpy(ta->g2, h->tail_g2, sizeof(ta->g2));
  ta->fast_pqmf = h->tail_fast;
}

// act: the operand tensor [B][Lfr + 1][C] conv_post consumes (reflect-padded lrelu_0.01 of the last stage)
int run_tail_fused(Ctx& cx, const void* act, float* wav, float* o_mb, float* spec, float* phase, int B, int Lfr) {
  mbv_handle* h = cx.h;
  TailArgs ta;
  fill_tail_args(h, &ta, wav, o_mb, spec, phase, B, Lfr);
  ProfScope prof(cx, 1, "tail (conv_post fused)");
  CUDA_TRY(h, launch_tail_fused(ta, act, h->conv_post.w, h->post_bias, h->conv_post.Cp_in, h->prec == MBV_PREC_FP16 ? 1 : 0,
                                h->num_sms, cx.st));
  cx.launches++;
  return MBV_OK;
}

int run_tail(Ctx& cx, const float* logits, float* wav, float* o_mb, float* spec, float* phase, int B, int Lfr) {
  mbv_handle* h = cx.h;
  TailArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.logits = logits; ta.wav = wav; ta.o_mb = o_mb; ta.spec = spec; ta.phase = phase;
  ta.B = B; ta.L = Lfr; ta.n_ch = h->n_logit; ta.variant = h->cfg.variant;
  memcpy(ta.coef, h->tail_coef, sizeof(ta.coef));
  memcpy(ta.mod, h->tail_mod, sizeof(ta.mod));
  memcpy(ta.g2, h->tail_g2, sizeof(ta.g2));
  ta.fast_pqmf = h->tail_fast;
  ProfScope prof(cx, 1, "tail");
  CUDA_TRY(h, launch_tail(ta, h->prec == MBV_PREC_FP32 ? 1 : 0, h->num_sms, cx.st));
  cx.launches++;
  return MBV_OK;
}

// z_op_ready: the operand copy of z already sits in d.zin_op (fused flow+decode path)
int run_decode(Ctx& cx, const DecBufs& d, const float* z, const float* z_mask, const void* z_op_ready, const float* g,
               float* wav, float* o_mb, float* spec, float* phase, int B, int T) {
  mbv_handle* h = cx.h;
  const mbv_config& c = h->cfg;
  int rc;
  const void* zin = z_op_ready;
  if (!zin) {
    ProfScope prof(cx, 2);
    CUDA_TRY(h, launch_pack_input(h->prec, z, z_mask, d.zin_op, nullptr, B, h->Cz, T, h->Cz, cx.st));
    cx.launches++;
    zin = d.zin_op;
  }
  const int C0 = c.upsample_initial_channel;
  {  // x = conv_pre(z); the only consumer is leaky_relu(x, 0.1) -> ups[0]
    EpiParams e = epi_base(EPI_ACT, C0, T);
    e.slope = 0.1f; e.act[0] = d.pre_act; e.n_act = 1;
    if ((rc = run_conv(cx, h->conv_pre, zin, B, T, T, e))) return rc;
  }
  const void* cur = d.pre_act;
  int L = T;
  for (int i = 0; i < h->n_stage; ++i) {
    const int S = c.upsample_rates[i], C = h->stage_C[i], Lin = L;
    L *= S;
    const auto& s = d.st[i];
    const bool last = (i == h->n_stage - 1);
    const int nk = c.n_kernels;
    // per-resblock conditioning vectors cond_j(g) [B][C] and (bias2_first + cond_j)
    float* cond[3] = {nullptr, nullptr, nullptr};
    float* bias2c[3] = {nullptr, nullptr, nullptr};
    if (g) {
      for (int j = 0; j < nk; ++j) {
        cond[j] = d.cond + ((size_t)(i * nk + j) * 2 + 0) * B * 512;
        bias2c[j] = d.cond + ((size_t)(i * nk + j) * 2 + 1) * B * 512;
        const ConvLayer& first_res = (c.resblock_type == 1) ? h->rb_c2[i][j][0] : h->rb_c1[i][j][0];
        ProfScope prof(cx, 2);
        CUDA_TRY(h, launch_cond_gemv(g, h->rb_cond_w[i][j], h->rb_cond_b[i][j], nullptr, cond[j], B, c.gin_channels, C, C, cx.st));
        CUDA_TRY(h, launch_cond_gemv(g, h->rb_cond_w[i][j], h->rb_cond_b[i][j], first_res.bias, bias2c[j], B, c.gin_channels, C, C, cx.st));
        cx.launches += 2;
      }
    }
    {  // x = ups[i](lrelu(x)): S polyphase branches; emits the fp32 residual stream and lrelu(x [+ cond_j]) operand copies
      EpiParams e = epi_base(EPI_ACT, C, L);
      e.rows_res = Lin; e.row_mul = S; e.slope = 0.1f; e.xout = s.x; e.res_half = h->res_half;
      if (h->single) { e.xout = nullptr; e.res_half = 0; }  // x lives in the operand copies lrelu(x [+ cond_j])
      e.n_act = g ? nk : 1;
      for (int j = 0; j < e.n_act; ++j) { e.act[j] = s.a[j]; e.act_add[j] = g ? cond[j] : nullptr; }
      e.act_add_bs = C;
      if ((rc = run_conv(cx, h->ups[i], cur, B, Lin, Lin, e))) return rc;
    }
    // The nk parallel ResBlocks of a stage are independent until their outputs are summed (models.py:355-361).  With
    // `branches` each runs on its own stream (its own residual / operand buffers): the persistent conv kernels own a whole
    // SM per CTA, so kernels of different branches do not share SMs -- but the CTAs of the next kernel move onto SMs as
    // soon as the CTAs of the running one retire, which hides every launch's fill / drain and partial last round behind
    // another branch's work.  Only the LAST launch of each ResBlock (the one that folds x into the running sum xs) stays
    // on the caller's stream, in ResBlock order.
    const bool branches = h->use_branches && nk > 1 && nk <= 3 && !h->profiling && h->prec != MBV_PREC_FP32 &&
                          !(c.flags & MBV_FLAG_FORCE_SIMT);
    cudaStream_t main_st = cx.st;
    const void* a_cur[3];
    const void* x_cur[3];
    const int np = c.n_dilations;
    // phase 0: everything but the last launch of ResBlock j;  phase 1: that last launch
    auto run_block = [&](int j, int phase) -> int {
      const int bj = branches ? j : 0;  // buffer set
      void* xr = s.xr[bj]; void* ar = s.ar[bj]; void* hop = s.hop[bj];
      if (phase == 0) { a_cur[j] = g ? s.a[j] : s.a[0]; x_cur[j] = s.x; }
      const void*& a_in = a_cur[j];
      const void*& x_in = x_cur[j];
      int rc2;
      for (int p = (phase == 0 ? 0 : np - 1); p < np; ++p) {
        const bool final_conv = (p == np - 1);
        EpiParams e = epi_base(EPI_RES, C, L);
        e.xin = x_in; e.slope = 0.1f; e.res_half = h->res_half;
        if (p == 0 && g && !h->single) { e.bias = bias2c[j]; e.bias_bs = C; }  // x + cond(g) folded into the first residual add
        if (h->single) { e.xin = a_in; e.res_half = 2; e.inv_slope = 10.f; }    // a_in = lrelu_0.1(x [+ cond_j]) -> x [+ cond_j]
        if (final_conv) {
          e.xs = s.xs;
          e.sum_mode = (nk == 1) ? 4 : (j == 0 ? 1 : (j == nk - 1 ? 3 : 2));
          if (e.sum_mode >= 3) {
            e.scale = 1.f / nk; e.act[0] = s.next; e.n_act = 1;
            e.slope = last ? 0.01f : 0.1f;
            if (last) { e.rows_out = L + 1; e.row_add = 1; e.dup_src = 2; e.dup_dst = 0; }
          }
        } else {
          e.xout = h->single ? nullptr : xr;
        }
        if (c.resblock_type == 1 && (pair_ok(h, h->rb_c1[i][j][p], h->rb_c2[i][j][p]) ||
                                     pairtm_ok(h, h->rb_c1[i][j][p], h->rb_c2[i][j][p], e.sum_mode))) {
          // one launch for the conv pair; the operand copies ping-pong (c1 of a neighbouring tile still reads a_in's halo
          // rows while this tile's epilogue writes its output)
          if (final_conv && phase == 0) break;
          void* a_out = (p & 1) ? hop : ar;
          if (!final_conv) { e.act[0] = a_out; e.n_act = 1; }
          if ((rc2 = run_pair(cx, h->rb_c1[i][j][p], h->rb_c2[i][j][p], a_in, B, L, e, 0.1f))) return rc2;
          a_in = a_out;
        } else if (c.resblock_type == 1) {
          // (after fused pairs the operand copy may live in `hop`: the intermediate and the next operand copy take the other buffers)
          void* hbuf = (a_in == hop) ? ar : hop;
          void* aout = (hbuf == ar) ? hop : ar;
          if (!(final_conv && phase == 1)) {  // c1 of the last pair still belongs to the branch
            EpiParams e1 = epi_base(EPI_ACT, C, L);
            e1.slope = 0.1f; e1.act[0] = hbuf; e1.n_act = 1;
            if ((rc2 = run_conv(cx, h->rb_c1[i][j][p], a_in, B, L, L, e1))) return rc2;
          }
          if (final_conv && phase == 0) break;
          if (!final_conv) { e.act[0] = aout; e.n_act = 1; }
          if ((rc2 = run_conv(cx, h->rb_c2[i][j][p], hbuf, B, L, L, e))) return rc2;
          a_in = aout;
        } else {
          // ResBlock2: x = x + c_p(lrelu(x)); operand copies ping-pong between ar and hop
          if (final_conv && phase == 0) break;
          void* a_out = (p & 1) ? hop : ar;
          if (!final_conv) { e.act[0] = a_out; e.n_act = 1; }
          if ((rc2 = run_conv(cx, h->rb_c1[i][j][p], a_in, B, L, L, e))) return rc2;
          a_in = a_out;
        }
        x_in = xr;
      }
      return MBV_OK;
    };
    if (branches) CUDA_TRY(h, cudaEventRecord(h->ev_fork, main_st));
    static const int branch_pdl = getenv("MBV_BRANCH_PDL") ? atoi(getenv("MBV_BRANCH_PDL")) : 0;  // A/B measurements only
    cx.pdl = !branches || branch_pdl;
    if (!branches) {  // one buffer set: each ResBlock runs to completion before the next starts
      for (int j = 0; j < nk; ++j)
        for (int phase = 0; phase < 2; ++phase)
          if ((rc = run_block(j, phase))) return rc;
    } else {
      for (int j = 0; j < nk; ++j) {
        if (j > 0) {
          cx.st = h->br_stream[j - 1];
          CUDA_TRY(h, cudaStreamWaitEvent(cx.st, h->ev_fork, 0));
        }
        if ((rc = run_block(j, 0))) { cx.st = main_st; cx.pdl = true; return rc; }
        if (j > 0) CUDA_TRY(h, cudaEventRecord(h->ev_join[j - 1], cx.st));
        cx.st = main_st;
      }
      for (int j = 0; j < nk; ++j) {
        if (j > 0) CUDA_TRY(h, cudaStreamWaitEvent(main_st, h->ev_join[j - 1], 0));
        if ((rc = run_block(j, 1))) { cx.pdl = true; return rc; }
      }
    }
    cx.pdl = true;
    cur = s.next;
  }
  if (fused_tail_ok(h)) return run_tail_fused(cx, cur, wav, o_mb, spec, phase, B, L);
  {  // conv_post on the reflect-padded L+1 frames -> fp32 logits, pitch n_logit
    EpiParams e = epi_base(EPI_F32, h->n_logit, L + 1);
    e.n_valid = h->n_logit; e.xout = d.logits;
    if ((rc = run_conv(cx, h->conv_post, cur, B, L + 1, L + 1, e))) return rc;
  }
  return run_tail(cx, d.logits, wav, o_mb, spec, phase, B, L);
}

int check_ready(mbv_handle* h, int B, int T) {
  if (!h) return MBV_ERR_INVALID;
  if (!h->weights_loaded) return fail(h, MBV_ERR_WEIGHTS, "mbv_load_weights has not completed");
  if (B < 1 || T < 1) return fail(h, MBV_ERR_INVALID, "B and T must be >= 1");
  if (B > 65535) return fail(h, MBV_ERR_UNSUPPORTED, "B > 65535");
  return MBV_OK;
}

int total_ws(mbv_handle* h, int B, int T, size_t* dec_off, size_t* total) {
  Arena A(nullptr);
  FlowBufs f;
  layout_flow(h, A, B, T, &f);
  *dec_off = (A.off + 1023) & ~(size_t)1023;
  DecBufs d;
  layout_dec(h, A, B, T, &d);
  *total = A.off + 1024;
  return 0;
}

Ctx make_ctx(mbv_handle* h, int B, int T, void* ws, int kind, void* stream) {
  Ctx cx;
  cx.h = h;
  cx.st = (cudaStream_t)stream;
  // (cached tensor maps bake in buffer addresses: the branch layout uses other ResBlock buffers than the sequential one)
  mbv_handle::PlanKey key{B, T, ws, kind + ((h->use_branches && !h->profiling) ? 16 : 0)};
  auto it = h->plan_cache.find(key);
  if (it == h->plan_cache.end()) {
    if (h->plan_cache.size() > 64) { h->plan_cache.clear(); h->pair_cache.clear(); }
    it = h->plan_cache.emplace(key, std::vector<TcPlan>()).first;
    cx.plans_valid = false;
  } else {
    cx.plans_valid = true;
  }
  cx.plans = &it->second;
  cx.pair_plans = &h->pair_cache[key];
  if (!cx.plans_valid) cx.pair_plans->clear();
  return cx;
}

int check_ws(mbv_handle* h, int B, int T, void* ws, size_t ws_bytes) {
  size_t dec_off, total;
  total_ws(h, B, T, &dec_off, &total);
  if (!ws || ((uintptr_t)ws & 1023) != 0) return fail(h, MBV_ERR_WORKSPACE, "workspace must be non-null and 1024-byte aligned");
  if (ws_bytes < total) return fail(h, MBV_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", total, ws_bytes);
  return MBV_OK;
}

}  // namespace

extern "C" int mbv_workspace_bytes(mbv_handle* h, int32_t B, int32_t T, size_t* bytes) {
  if (!h || !bytes) return MBV_ERR_INVALID;
  if (B < 1 || T < 1) return fail(h, MBV_ERR_INVALID, "B and T must be >= 1");
  size_t dec_off;
  return total_ws(h, B, T, &dec_off, bytes);
}

extern "C" int mbv_flow_reverse(mbv_handle* h, const float* z_p, const float* y_mask, const float* g, float* z_out,
                                int32_t B, int32_t T, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!z_p || !y_mask || !z_out) return fail(h, MBV_ERR_INVALID, "null tensor");
  if (g && h->cfg.gin_channels == 0) return fail(h, MBV_ERR_INVALID, "g given but gin_channels == 0");
  if ((rc = check_ws(h, B, T, ws, ws_bytes))) return rc;
  DEVICE_GUARD(h);
  Arena A(ws);
  FlowBufs f;
  layout_flow(h, A, B, T, &f);
  Ctx cx = make_ctx(h, B, T, ws, g ? 1 : 0, stream);
  rc = run_flow(cx, f, z_p, y_mask, g, z_out, B, T);
  if (rc) { cx.plans->clear(); return rc; }
  h->last_launches = cx.launches;
  return MBV_OK;
}

extern "C" int mbv_flow_forward(mbv_handle* h, const float* x, const float* y_mask, const float* g, float* z_out, int32_t B,
                                int32_t T, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!x || !y_mask || !z_out) return fail(h, MBV_ERR_INVALID, "null tensor");
  if (g && h->cfg.gin_channels == 0) return fail(h, MBV_ERR_INVALID, "g given but gin_channels == 0");
  if ((rc = check_ws(h, B, T, ws, ws_bytes))) return rc;
  DEVICE_GUARD(h);
  Arena A(ws);
  FlowBufs f;
  layout_flow(h, A, B, T, &f);
  Ctx cx = make_ctx(h, B, T, ws, g ? 7 : 6, stream);
  rc = run_flow(cx, f, x, y_mask, g, z_out, B, T, true);
  if (rc) { cx.plans->clear(); return rc; }
  h->last_launches = cx.launches;
  return MBV_OK;
}

extern "C" int mbv_decode(mbv_handle* h, const float* z, const float* z_mask, const float* g, float* wav, float* o_mb,
                          float* spec, float* phase, int32_t B, int32_t T, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!z || !wav) return fail(h, MBV_ERR_INVALID, "null tensor");
  if (g && h->cfg.gin_channels == 0) return fail(h, MBV_ERR_INVALID, "g given but gin_channels == 0");
  if ((spec == nullptr) != (phase == nullptr)) return fail(h, MBV_ERR_INVALID, "spec and phase must be requested together");
  if (o_mb && h->cfg.variant == MBV_VARIANT_ISTFT) return fail(h, MBV_ERR_INVALID, "the single-band decoder has no o_mb (models.py:297 returns None)");
  if ((rc = check_ws(h, B, T, ws, ws_bytes))) return rc;
  DEVICE_GUARD(h);
  size_t dec_off, total;
  total_ws(h, B, T, &dec_off, &total);
  Arena A(ws);
  A.off = dec_off;
  DecBufs d;
  layout_dec(h, A, B, T, &d);
  Ctx cx = make_ctx(h, B, T, ws, g ? 3 : 2, stream);
  rc = run_decode(cx, d, z, z_mask, nullptr, g, wav, o_mb, spec, phase, B, T);
  if (rc) { cx.plans->clear(); return rc; }
  h->last_launches = cx.launches;
  return MBV_OK;
}

extern "C" int mbv_flow_decode(mbv_handle* h, const float* z_p, const float* y_mask, const float* g, float* z_out,
                               float* wav, float* o_mb, float* spec, float* phase, int32_t B, int32_t T, void* ws,
                               size_t ws_bytes, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!z_p || !y_mask || !wav) return fail(h, MBV_ERR_INVALID, "null tensor");
  if (g && h->cfg.gin_channels == 0) return fail(h, MBV_ERR_INVALID, "g given but gin_channels == 0");
  if ((spec == nullptr) != (phase == nullptr)) return fail(h, MBV_ERR_INVALID, "spec and phase must be requested together");
  if (o_mb && h->cfg.variant == MBV_VARIANT_ISTFT) return fail(h, MBV_ERR_INVALID, "the single-band decoder has no o_mb");
  if ((rc = check_ws(h, B, T, ws, ws_bytes))) return rc;
  DEVICE_GUARD(h);
  size_t dec_off, total;
  total_ws(h, B, T, &dec_off, &total);
  Arena A(ws);
  FlowBufs f;
  layout_flow(h, A, B, T, &f);
  A.off = dec_off;
  DecBufs d;
  layout_dec(h, A, B, T, &d);
  Ctx cx = make_ctx(h, B, T, ws, g ? 5 : 4, stream);
  rc = run_flow(cx, f, z_p, y_mask, g, z_out, B, T);
  // after the four masked couplings z is already zero on padded frames, so z * y_mask is the
  // identity and the flow's operand copy feeds conv_pre directly
  if (!rc) rc = run_decode(cx, d, nullptr, nullptr, f.zop, g, wav, o_mb, spec, phase, B, T);
  if (rc) { cx.plans->clear(); return rc; }
  h->last_launches = cx.launches;
  return MBV_OK;
}

extern "C" int mbv_posterior_workspace_bytes(mbv_handle* h, int32_t B, int32_t T, size_t* bytes) {
  if (!h || !bytes) return MBV_ERR_INVALID;
  if (!h->has_enc_q) return fail(h, MBV_ERR_WEIGHTS, "no enc_q.* weights were loaded");
  if (B < 1 || T < 1) return fail(h, MBV_ERR_INVALID, "B and T must be >= 1");
  Arena A(nullptr);
  PostBufs p;
  layout_posterior(h, A, B, T, &p);
  *bytes = A.off + 1024;
  return MBV_OK;
}

extern "C" int mbv_posterior_encode(mbv_handle* h, const float* spec, const float* y_mask, const float* g, const float* noise,
                                    float* z, float* stats, int32_t B, int32_t T, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!h->has_enc_q) return fail(h, MBV_ERR_WEIGHTS, "no enc_q.* weights were loaded");
  if (!spec || !y_mask || !noise || !z || !stats) return fail(h, MBV_ERR_INVALID, "null tensor");
  if (g && h->cfg.gin_channels == 0) return fail(h, MBV_ERR_INVALID, "g given but gin_channels == 0");
  size_t need = 0;
  mbv_posterior_workspace_bytes(h, B, T, &need);
  if (!ws || ((uintptr_t)ws & 1023) != 0) return fail(h, MBV_ERR_WORKSPACE, "workspace must be non-null and 1024-byte aligned");
  if (ws_bytes < need) return fail(h, MBV_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  DEVICE_GUARD(h);
  Arena A(ws);
  PostBufs p;
  layout_posterior(h, A, B, T, &p);
  Ctx cx = make_ctx(h, B, T, ws, g ? 9 : 8, stream);
  rc = run_posterior(cx, p, spec, y_mask, g, noise, z, stats, B, T);
  if (rc) { cx.plans->clear(); return rc; }
  h->last_launches = cx.launches;
  return MBV_OK;
}

extern "C" int mbv_text_workspace_bytes(mbv_handle* h, int32_t B, int32_t T, size_t* bytes) {
  if (!h || !bytes) return MBV_ERR_INVALID;
  if (!h->has_enc_p) return fail(h, MBV_ERR_WEIGHTS, "no enc_p.* weights were loaded");
  if (B < 1 || T < 1) return fail(h, MBV_ERR_INVALID, "B and T must be >= 1");
  Arena A(nullptr);
  TextBufs t;
  layout_text(h, A, B, T, &t);
  *bytes = A.off + 1024;
  return MBV_OK;
}

extern "C" int mbv_text_encode(mbv_handle* h, const int64_t* tokens, const float* x_mask, float* x_out, float* stats, int32_t B,
                               int32_t T, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!h->has_enc_p) return fail(h, MBV_ERR_WEIGHTS, "no enc_p.* weights were loaded");
  if (!tokens || !x_mask || !x_out || !stats) return fail(h, MBV_ERR_INVALID, "null tensor");
  if (T > 2048) return fail(h, MBV_ERR_UNSUPPORTED, "mbv_text_encode: more than 2048 tokens per utterance");
  size_t need = 0;
  mbv_text_workspace_bytes(h, B, T, &need);
  if (!ws || ((uintptr_t)ws & 1023) != 0) return fail(h, MBV_ERR_WORKSPACE, "workspace must be non-null and 1024-byte aligned");
  if (ws_bytes < need) return fail(h, MBV_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  DEVICE_GUARD(h);
  Arena A(ws);
  TextBufs t;
  layout_text(h, A, B, T, &t);
  Ctx cx = make_ctx(h, B, T, ws, 12, stream);
  rc = run_text(cx, t, (const long long*)tokens, x_mask, x_out, stats, B, T);
  if (rc) { cx.plans->clear(); return rc; }
  h->last_launches = cx.launches;
  return MBV_OK;
}

extern "C" int mbv_tail(mbv_handle* h, const float* logits, float* wav, float* o_mb, float* spec, float* phase,
                        int32_t B, int32_t T, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!logits || !wav) return fail(h, MBV_ERR_INVALID, "null tensor");
  if ((spec == nullptr) != (phase == nullptr)) return fail(h, MBV_ERR_INVALID, "spec and phase must be requested together");
  DEVICE_GUARD(h);
  int L = T;
  for (int i = 0; i < h->n_stage; ++i) L *= h->cfg.upsample_rates[i];
  Ctx cx;
  cx.h = h; cx.st = (cudaStream_t)stream; cx.plans = nullptr; cx.plans_valid = false;
  rc = run_tail(cx, logits, wav, o_mb, spec, phase, B, L);
  h->last_launches = cx.launches;
  return rc;
}

extern "C" int mbv_tail_fused(mbv_handle* h, const void* act, float* wav, float* o_mb, float* spec, float* phase, int32_t B,
                              int32_t T, void* stream) {
  int rc = check_ready(h, B, T);
  if (rc) return rc;
  if (!act || !wav) return fail(h, MBV_ERR_INVALID, "null tensor");
  if ((spec == nullptr) != (phase == nullptr)) return fail(h, MBV_ERR_INVALID, "spec and phase must be requested together");
  if (!fused_tail_ok(h)) return fail(h, MBV_ERR_UNSUPPORTED, "the fused conv_post + tail kernel needs a 16-bit precision and a 4-band decoder");
  DEVICE_GUARD(h);
  int L = T;
  for (int i = 0; i < h->n_stage; ++i) L *= h->cfg.upsample_rates[i];
  Ctx cx;
  cx.h = h; cx.st = (cudaStream_t)stream; cx.plans = nullptr; cx.plans_valid = false;
  rc = run_tail_fused(cx, act, wav, o_mb, spec, phase, B, L);
  h->last_launches = cx.launches;
  return rc;
}

extern "C" int mbv_pcm16(mbv_handle* h, const float* wav, const int32_t* n_samples, int32_t B, int32_t stride,
                         int32_t auto_normalize, void* scratch, int16_t* pcm, void* stream) {
  if (!h) return MBV_ERR_INVALID;
  if (!wav || !pcm || !scratch || B < 1 || stride < 1) return fail(h, MBV_ERR_INVALID, "mbv_pcm16: bad argument");
  if (B > 65535) return fail(h, MBV_ERR_UNSUPPORTED, "B > 65535");
  DEVICE_GUARD(h);
  CUDA_TRY(h, launch_pcm16(wav, n_samples, B, stride, auto_normalize, (unsigned int*)scratch, pcm, (cudaStream_t)stream));
  h->last_launches = 2;
  return MBV_OK;
}

extern "C" int mbv_expand_prior(mbv_handle* h, const float* m_p, const float* logs_p, const float* w_ceil, const float* x_mask,
                                const float* noise, float noise_scale, int32_t B, int32_t C, int32_t Tx, int32_t Ty,
                                float* z_p, float* y_mask, float* m_exp, float* logs_exp, float* attn, int64_t* y_lengths,
                                void* stream) {
  if (!h) return MBV_ERR_INVALID;
  if (!m_p || !logs_p || !w_ceil || !noise || !z_p || !y_mask) return fail(h, MBV_ERR_INVALID, "mbv_expand_prior: null tensor");
  if (B < 1 || C < 1 || Tx < 1 || Ty < 1) return fail(h, MBV_ERR_INVALID, "mbv_expand_prior: sizes must be >= 1");
  if (B > 65535) return fail(h, MBV_ERR_UNSUPPORTED, "B > 65535");
  if (Tx > 12000) return fail(h, MBV_ERR_UNSUPPORTED, "mbv_expand_prior: more than 12000 tokens per utterance");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(h, MBV_ERR_CUDA, "mbv_expand_prior: no CUDA device (there is no CPU fallback)");
  }
  DEVICE_GUARD(h);
  CUDA_TRY(h, launch_expand_prior(m_p, logs_p, w_ceil, x_mask, noise, noise_scale, B, C, Tx, Ty, z_p, y_mask, m_exp, logs_exp,
                                  attn, (long long*)y_lengths, (cudaStream_t)stream));
  h->last_launches = 1;
  return MBV_OK;
}

// =================================================================================================
// streaming decode: exact chunked decoding with receptive-field halos carried in device state
// =================================================================================================
struct mbv_stream {
  mbv_handle* h = nullptr;
  int B = 0, max_chunk = 0, halo = 0, cap = 0;
  float* hist[2] = {nullptr, nullptr};  // [B][inter][cap] latent history (ping-pong for the shift), column 0 = frame hist_start
  float* win = nullptr;                 // [B][inter][W] contiguous decode window
  float* wav = nullptr;                 // [B][1][spf * cap] decoded window
  int cur = 0;
  long long received = 0, emitted = 0, hist_start = 0;
  bool finished = false;
};

namespace {
// Upper bound (latent frames, one side) of the decoder's receptive field; mirrors configs.receptive_field_frames
int receptive_field(const mbv_handle* h) {
  const mbv_config& c = h->cfg;
  double rf = 3.0, rate = 1.0;
  for (int i = 0; i < c.n_ups; ++i) {
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    rf += (double)((k + 2 * u - 1) / (2 * u)) / rate;
    rate *= u;
    int widest = 0;
    for (int j = 0; j < c.n_kernels; ++j) {
      const int rk = c.resblock_kernel_sizes[j];
      int r = 0;
      for (int p = 0; p < c.n_dilations; ++p) r += (rk - 1) / 2 * c.resblock_dilations[j][p];
      if (c.resblock_type == 1) r += (rk - 1) / 2 * c.n_dilations;
      if (r > widest) widest = r;
    }
    rf += widest / rate;
  }
  rf += (3 + 1 + 4 + (c.variant != MBV_VARIANT_ISTFT ? 3 : 0)) / rate;
  return (int)ceil(rf);
}
}  // namespace

extern "C" int mbv_receptive_field(mbv_handle* h) { return h ? receptive_field(h) : 0; }

extern "C" int mbv_stream_open(mbv_handle* h, int32_t B, int32_t max_chunk_frames, mbv_stream** out) {
  if (!h || !out) return MBV_ERR_INVALID;
  *out = nullptr;
  int rc = check_ready(h, B, max_chunk_frames);
  if (rc) return rc;
  DEVICE_GUARD(h);
  mbv_stream* s = new mbv_stream();
  s->h = h; s->B = B; s->max_chunk = max_chunk_frames;
  s->halo = receptive_field(h) + 1;
  s->cap = max_chunk_frames + 3 * s->halo;
  const size_t nz = (size_t)B * h->Cz * s->cap * sizeof(float);
  cudaError_t e = cudaMalloc(&s->hist[0], nz);
  if (e == cudaSuccess) e = cudaMalloc(&s->hist[1], nz);
  if (e == cudaSuccess) e = cudaMalloc(&s->win, nz);
  if (e == cudaSuccess) e = cudaMalloc(&s->wav, (size_t)B * h->spf * s->cap * sizeof(float));
  if (e != cudaSuccess) {
    for (float* p : {s->hist[0], s->hist[1], s->win, s->wav}) if (p) cudaFree(p);
    delete s;
    return fail(h, MBV_ERR_CUDA, "mbv_stream_open: %s", cudaGetErrorString(e));
  }
  *out = s;
  return MBV_OK;
}

extern "C" void mbv_stream_close(mbv_stream* s) {
  if (!s) return;
  for (float* p : {s->hist[0], s->hist[1], s->win, s->wav}) if (p) cudaFree(p);
  delete s;
}

extern "C" int mbv_stream_workspace_bytes(mbv_stream* s, size_t* bytes) {
  if (!s || !bytes) return MBV_ERR_INVALID;
  return mbv_workspace_bytes(s->h, s->B, s->cap, bytes);
}

extern "C" int mbv_stream_halo(mbv_stream* s) { return s ? s->halo : 0; }

extern "C" int mbv_stream_push(mbv_stream* s, const float* z_chunk, int32_t n_frames, int32_t last, const float* g,
                               float* wav_out, int32_t wav_capacity_frames, int64_t* first_frame, int32_t* n_frames_out,
                               void* ws, size_t ws_bytes, void* stream) {
  if (!s || !first_frame || !n_frames_out) return MBV_ERR_INVALID;
  mbv_handle* h = s->h;
  *n_frames_out = 0;
  *first_frame = s->emitted;
  if (s->finished) return fail(h, MBV_ERR_INVALID, "mbv_stream_push after the last chunk");
  if (n_frames < 0 || n_frames > s->max_chunk) return fail(h, MBV_ERR_INVALID, "chunk of %d frames (stream opened for <= %d)", n_frames, s->max_chunk);
  if (n_frames > 0 && !z_chunk) return fail(h, MBV_ERR_INVALID, "null chunk");
  if (g && h->cfg.gin_channels == 0) return fail(h, MBV_ERR_INVALID, "g given but gin_channels == 0");
  DEVICE_GUARD(h);
  cudaStream_t st = (cudaStream_t)stream;
  const int Cz = h->Cz, rows = s->B * Cz;
  const size_t cap_pitch = (size_t)s->cap * sizeof(float);
  if (n_frames > 0) {  // append the chunk to the history
    const long long col = s->received - s->hist_start;
    CUDA_TRY(h, cudaMemcpy2DAsync(s->hist[s->cur] + col, cap_pitch, z_chunk, (size_t)n_frames * sizeof(float),
                                  (size_t)n_frames * sizeof(float), rows, cudaMemcpyDeviceToDevice, st));
    s->received += n_frames;
  }
  // frames [emitted, emit_end) have their whole right context (or the stream ends here)
  long long emit_end = last ? s->received : s->received - s->halo;
  if (emit_end < s->emitted) emit_end = s->emitted;
  const int n_emit = (int)(emit_end - s->emitted);
  if (last) s->finished = true;
  if (n_emit == 0) return MBV_OK;
  if (!wav_out || wav_capacity_frames < n_emit) return fail(h, MBV_ERR_INVALID, "wav_out holds %d frames, %d are due", wav_capacity_frames, n_emit);
  const long long w0 = s->emitted - s->halo > 0 ? s->emitted - s->halo : 0;
  const long long w1 = emit_end + s->halo < s->received ? emit_end + s->halo : s->received;
  const int W = (int)(w1 - w0);
  int rc = check_ws(h, s->B, W, ws, ws_bytes);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpy2DAsync(s->win, (size_t)W * sizeof(float), s->hist[s->cur] + (w0 - s->hist_start), cap_pitch,
                                (size_t)W * sizeof(float), rows, cudaMemcpyDeviceToDevice, st));
  {
    size_t dec_off, total;
    total_ws(h, s->B, W, &dec_off, &total);
    Arena A(ws);
    A.off = dec_off;
    DecBufs d;
    layout_dec(h, A, s->B, W, &d);
    Ctx cx = make_ctx(h, s->B, W, ws, g ? 11 : 10, stream);
    rc = run_decode(cx, d, s->win, nullptr, nullptr, g, s->wav, nullptr, nullptr, nullptr, s->B, W);
    if (rc) { cx.plans->clear(); return rc; }
    h->last_launches = cx.launches + 3;
  }
  const int spf = h->spf;
  CUDA_TRY(h, cudaMemcpy2DAsync(wav_out, (size_t)n_emit * spf * sizeof(float), s->wav + (size_t)(s->emitted - w0) * spf,
                                (size_t)W * spf * sizeof(float), (size_t)n_emit * spf * sizeof(float), s->B,
                                cudaMemcpyDeviceToDevice, st));
  *n_frames_out = n_emit;
  s->emitted = emit_end;
  // keep [emitted - halo, received) for the next push
  const long long keep0 = s->emitted - s->halo > 0 ? s->emitted - s->halo : 0;
  if (!last && keep0 > s->hist_start) {
    const long long n_keep = s->received - keep0;
    if (n_keep > 0)
      CUDA_TRY(h, cudaMemcpy2DAsync(s->hist[s->cur ^ 1], cap_pitch, s->hist[s->cur] + (keep0 - s->hist_start), cap_pitch,
                                    (size_t)n_keep * sizeof(float), rows, cudaMemcpyDeviceToDevice, st));
    s->cur ^= 1;
    s->hist_start = keep0;
  }
  return MBV_OK;
}

extern "C" int mbv_last_launch_count(mbv_handle* h) { return h ? h->last_launches : 0; }

extern "C" int mbv_set_profiling(mbv_handle* h, int32_t on) {
  if (!h) return MBV_ERR_INVALID;
  h->profiling = on != 0;
  return MBV_OK;
}

extern "C" int mbv_profile_read(mbv_handle* h, double* ms, int32_t* count) {
  if (!h || !ms || !count) return MBV_ERR_INVALID;
  for (auto& r : h->prof) {
    CUDA_TRY(h, cudaEventSynchronize(r.e1));
    float t = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&t, r.e0, r.e1));
    if (r.kind >= 0 && r.kind < 3) { ms[r.kind] += t; count[r.kind] += 1; }
    h->ev_pool.push_back(r.e0);
    h->ev_pool.push_back(r.e1);
  }
  h->prof.clear();
  return MBV_OK;
}

extern "C" int mbv_profile_read_launches(mbv_handle* h, float* ms, char* desc, int32_t desc_stride, int32_t cap, int32_t* n) {
  if (!h || !ms || !n || cap < 0) return MBV_ERR_INVALID;
  int k = 0;
  for (auto& r : h->prof) {
    CUDA_TRY(h, cudaEventSynchronize(r.e1));
    float t = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&t, r.e0, r.e1));
    if (k < cap) {
      ms[k] = t;
      if (desc && desc_stride > 0) snprintf(desc + (size_t)k * desc_stride, desc_stride, "%s", r.desc);
      ++k;
    }
    h->ev_pool.push_back(r.e0);
    h->ev_pool.push_back(r.e1);
  }
  h->prof.clear();
  *n = k;
  return MBV_OK;
}

extern "C" double mbv_decode_flops(mbv_handle* h, int32_t B, int32_t T) {
  if (!h || !h->weights_loaded) return 0.0;
  const mbv_config& c = h->cfg;
  double macs = h->conv_pre.macs_per_row * T;
  double L = T;
  for (int i = 0; i < h->n_stage; ++i) {
    macs += h->ups[i].macs_per_row * L;
    L *= c.upsample_rates[i];
    for (int j = 0; j < c.n_kernels; ++j)
      for (int p = 0; p < c.n_dilations; ++p) {
        macs += h->rb_c1[i][j][p].macs_per_row * L;
        if (c.resblock_type == 1) macs += h->rb_c2[i][j][p].macs_per_row * L;
      }
  }
  macs += h->conv_post.macs_per_row * (L + 1);
  return 2.0 * macs * B;
}

extern "C" double mbv_flow_flops(mbv_handle* h, int32_t B, int32_t T) {
  if (!h || !h->weights_loaded) return 0.0;
  double macs = 0;
  for (int f = 0; f < 4; ++f) {
    macs += h->fl_pre[f].macs_per_row + h->fl_post[f].macs_per_row;
    for (int l = 0; l < h->cfg.flow_layers; ++l) macs += h->fl_in[f][l].macs_per_row + h->fl_rs[f][l].macs_per_row;
  }
  return 2.0 * macs * T * B;
}
