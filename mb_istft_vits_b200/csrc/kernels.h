// kernels.h -- host-callable launchers of the CUDA kernels (internal to libmbistft.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace mbv {

// ---- conv_simt.cu
cudaError_t launch_conv_simt(int prec, const ConvArgs& a, cudaStream_t st);

// ---- conv_tc.cu
struct TcPlan {
  CUtensorMap tmA;      // activations: 3-D (Cp_in, L_in, B), box (KB, box_rows, 1), 128B swizzle
  CUtensorMap tmB;      // weights: 2-D (Cp_in, phases*taps*N_total), box (KB, 128), 128B swizzle
  CUtensorMap tmR;      // residual input (EPI_RES xin): 3-D (C, rows, B), box (32, 32, 1): the staged residual chunks of the epilogue warps
  CUtensorMap tmS;      // running ResBlock sum xs of the summing residual epilogues, same box as tmR
  CUtensorMap tmBh;     // weights with a 64-row box: the half tile a CTA of a cluster pair fetches and multicasts
  int pair_phase;       // cluster == 2: pairs are polyphase branches instead of channel tiles
  int cluster, rows, groups;  // cluster mode (pairs of CTAs share every weight tile): see conv_tc.cu
  int prefetch_res;     // 1: tmR is valid
  int n_time;           // time columns per tile (UMMA N)
  int slab_rows, box_rows, n_boxes;
  int slab_stage_bytes, w_stage_bytes, n_slab_stages, n_w_stages;
  int t_tiles, c_tiles, total_tiles;
  int xchg_off;         // byte offset of the gate exchange buffer
  int res_off;          // > 0: byte offset of the per-warp staged-residual rings (tmR then has a 32 x 32 box)
  int w_resident;       // weights of the single channel tile stay resident in the ring
  int smem_bytes;
  int grid;
  int no_pair_split;    // MBV_FLAG_NO_PAIR_SPLIT: A/B only
  // pointwise (1x1) convs on pw_tc_kernel (time on the accumulator lane, pw_tc.cu): tmA = flattened [C, B*L] activation map,
  // tmB = weight map with an N-row box
  int pw, pw_N, pw_kblocks, pw_R, pw_tiles, pw_w_bytes, pw_a_stage_bytes, pw_a_stages, pw_a_off, pw_bias_off, pw_bar_off;
};
// ---- pw_tc.cu
bool pw_eligible(int prec, const ConvArgs& a, int flags);
const char* pw_make_plan(int prec, const ConvArgs& a, int num_sms, TcPlan* plan);
bool gt_eligible(int prec, const ConvArgs& a, int flags, int num_sms);   // WN gate conv on gate_tm_kernel (plan->pw = 2)
const char* gt_make_plan(int prec, const ConvArgs& a, int num_sms, TcPlan* plan);
bool ct_eligible(int prec, const ConvArgs& a, int flags, int num_sms);   // k >= 5 convs of a 128-channel ResBlock stage on conv_tm_kernel (plan->pw = 3)
const char* ct_make_plan(int prec, const ConvArgs& a, int num_sms, TcPlan* plan);
cudaError_t launch_pw(int prec, const ConvArgs& a, const TcPlan& plan, cudaStream_t st, int pdl);
cudaError_t pw_set_attributes();
// Fills plan (tensor maps, staging, grid) for args; returns a message on failure, nullptr on success.
const char* tc_make_plan(int prec, const ConvArgs& a, int flags, int num_sms, TcPlan* plan);
cudaError_t launch_conv_tc(int prec, const ConvArgs& a, const TcPlan& plan, cudaStream_t st, int pdl = 1);
cudaError_t tc_set_attributes();
// fused ResBlock1 conv pair (c1 -> lrelu -> c2 -> residual add), 128-channel stages: see conv_tc.cu
struct TcPairPlan {
  CUtensorMap tmA, tmB, tmB2;   // activations, c1 weights, c2 weights
  int box_rows, n_boxes, slab_stage_bytes, n_slab_stages, n_w_stages;
  int t_tiles, total_tiles, h_off, w_off, bar_off, smem_bytes, grid;
  // pair_tm_kernel (time on the accumulator lane, CTA pairs, resident weights; pw_tc.cu): tm = 1
  int tm, tm_out_rows, tm_slab_kb_bytes, tm_slab_stage_bytes, tm_h_kb_bytes, tm_slab_off, tm_bias_off;
};
bool ptm_eligible(int prec, const ConvArgs& a, int flags, int num_sms);
const char* ptm_make_plan(int prec, const ConvArgs& a, int num_sms, TcPairPlan* plan);
cudaError_t launch_ptm(int prec, const ConvArgs& a, const TcPairPlan& plan, cudaStream_t st, int pdl);
const char* tc_make_pair_plan(int prec, const ConvArgs& a, int num_sms, TcPairPlan* plan);
cudaError_t launch_conv_pair(int prec, const ConvArgs& a, const TcPairPlan& plan, cudaStream_t st, int pdl = 1);
cudaError_t tc_pair_set_attributes();
// cuTensorMapEncodeTiled through the runtime's driver entry point (nullptr if unavailable); shared with tail.cu
void* tc_tensormap_encoder();

// ---- tail.cu
struct TailArgs {
  const float* logits;  // [B][F][n_ch] channels-last, F = L + 1 frames
  float* wav;           // [B][1][subbands*4*L]  (iSTFT: [B][1][4L])
  float* o_mb;          // MB: [B][S][4L]; MS: [B][S][16L] zero-stuffed; or null
  float* spec;          // [B][S][9][F] or null
  float* phase;
  int B, L, n_ch, variant;
  float coef[4][64];    // generic polyphase synthesis FIR G[c][r*16+(d+7)] = 4*h[c][4d+31-r] (trainable MS filter)
  float mod[8][4];      // PQMF fast path: cosine modulation 2*cos(theta_c(m)), m = k mod 8
  float g2[4][16];      // PQMF fast path: 4 * prototype[4d+31-r] * (-1)^floor(k/8), per output residue r
  int fast_pqmf;        // 1: variant MB (cosine-modulated bank): modulate once per sub-band sample, 16 MACs per output
  long long* dbg;       // MBV_TAIL_TIMELINE=1: CTA 0 / thread 0 clock stamps [tile < 8][8] (debug only)
};
cudaError_t launch_tail(const TailArgs& a, int precise, int num_sms, cudaStream_t st);
// conv_post + tail in one kernel (tail.cu): act = the 16-bit operand tensor [B][L+1][C] the last ResBlock wrote (reflect-padded,
// lrelu'd), w = conv_post's packed weights [7][128][C], bias = 72 HOST floats; C = 64 or 128; MB / MS variants only
cudaError_t launch_tail_fused(const TailArgs& t, const void* act, const void* w, const float* bias, int C, int f16, int num_sms,
                              cudaStream_t st);

// ---- misc.cu
// fp32 NCT [B][C][T] -> channels-last [B][T][Cp] operand (and optional fp32 copy), optional mask [B][T]
cudaError_t launch_pack_input(int prec, const float* src, const float* mask, void* dst_op, float* dst_f32,
                              int B, int C, int T, int Cp, cudaStream_t st);
// fp32 channels-last [B][T][Cp] -> NCT [B][C][T]
cudaError_t launch_unpack_output(const float* src, float* dst, int B, int C, int T, int Cp, cudaStream_t st);
// out[b][n] = bias[n] + sum_c w[n][c] * g[b][c]  (+ base[n] if given); w is [N][G] fp32
cudaError_t launch_cond_gemv(const float* g, const float* w, const float* bias, const float* base, float* out,
                             int B, int G, int N, int out_ld, cudaStream_t st);

// z = (m + noise * exp(logs)) * mask with stats = [m | logs] [B][2C][T], everything NCT (models.py:243-245)
cudaError_t launch_posterior_sample(const float* stats, const float* noise, const float* mask, float* z, int B, int C, int T,
                                    cudaStream_t st);

// per-utterance peak normalise (x0.9 if peak > 0.01), clip, x32767, truncate to int16 (tts_vits.py:204-216)
cudaError_t launch_pcm16(const float* wav, const int* n_valid, int B, int stride, int auto_normalize, unsigned int* peak_bits,
                         short* pcm, cudaStream_t st);

// alignment expansion + prior sampling (models.py:717-729): see misc.cu
cudaError_t launch_expand_prior(const float* m_p, const float* logs_p, const float* w_ceil, const float* x_mask,
                                const float* noise, float noise_scale, int B, int C, int Tx, int Ty, float* z_p,
                                float* y_mask, float* m_exp, float* logs_exp, float* attn, long long* y_lengths,
                                cudaStream_t st);

// ---- text.cu (text encoder helpers; ld = channel pitch of the operand copies)
cudaError_t launch_text_embed(const long long* tokens, const float* emb, const float* mask, float* x, void* xop, int n_rows, int C,
                              int ld_op, int n_vocab, int prec, cudaStream_t st);
cudaError_t launch_text_ln(const float* x, const float* y, int y_ld, const float* gamma, const float* beta, const float* mask,
                           float* x_out, void* xop, int n_rows, int C, int ld_op, int mask_out, int prec, cudaStream_t st);
cudaError_t launch_text_attention(const float* qkv, const float* mask, const float* rel_k, const float* rel_v, void* out, int B,
                                  int T, int C, int ld_op, int n_heads, int W, int prec, cudaStream_t st);

}  // namespace mbv
