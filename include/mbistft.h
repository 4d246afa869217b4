/*
 * mbistft.h -- C ABI of libmbistft.so: B200 (sm_100a) flow-reverse + iSTFT waveform decoders.
 *
 * This is the drop-in boundary for the reference's waveform hot path.  The reference has no plugin /
 * operator registry; the seam is two nn.Module attributes of SynthesizerTrn (models.py:634-647):
 *
 *     z = self.flow(z_p, y_mask, g=g, reverse=True)                   models.py:730
 *     o, o_mb, spec, phase = self.dec((z*y_mask)[:,:,:max_len], g=g)  models.py:734
 *
 * The Python shims in mb_istft_vits_b200/modules.py (NativeFlow / NativeDecoder) are assigned onto
 * those attributes and call the entry points below through ctypes.  Conventions:
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every call returns 0 on success, a negative mbv_status otherwise; mbv_last_error() gives text.
 *   - the caller owns every activation / output / workspace buffer (device memory); the library owns
 *     only its packed weight copy.  No allocation and no host synchronisation on the call path; all
 *     work is enqueued on the caller's stream (CUDA-graph capturable).
 *   - tensors at the boundary are fp32, contiguous, NCT (time contiguous), like the reference.
 *   - there is no CPU fallback: every compute entry fails with MBV_ERR_CUDA without a B200.
 */
#ifndef MBISTFT_H_
#define MBISTFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBV_ABI_VERSION 1

typedef enum {
  MBV_OK = 0,
  MBV_ERR_INVALID = -1,      /* bad argument / inconsistent sizes */
  MBV_ERR_UNSUPPORTED = -2,  /* geometry outside what the kernels implement (never a silent fallback) */
  MBV_ERR_WEIGHTS = -3,      /* missing / mis-shaped tensor in mbv_load_weights */
  MBV_ERR_WORKSPACE = -4,    /* workspace too small or misaligned */
  MBV_ERR_CUDA = -5          /* CUDA runtime / driver error (message has the detail) */
} mbv_status;

/* decoder family: models.py:248 iSTFT_Generator, :309 Multiband_iSTFT_Generator,
 * :387 Multistream_iSTFT_Generator (selected by the three booleans at models.py:634-644) */
typedef enum { MBV_VARIANT_ISTFT = 0, MBV_VARIANT_MB = 1, MBV_VARIANT_MS = 2 } mbv_variant;

/* arithmetic of the dense contractions (accumulators, head, iSTFT, PQMF are always fp32; the residual streams are fp32 on the
 * FP32 / TF32 paths and on BF16 without MBV_FLAG_RESIDUAL_FP16, saturating fp16 on BF16 with it, and folded into the fp16
 * operand tensors on FP16 -- DESIGN.md section 3) */
typedef enum {
  MBV_PREC_FP32 = 0, /* CUDA-core fp32 FMA: exact-order-independent reference path, slow */
  MBV_PREC_TF32 = 1, /* tcgen05 kind::tf32, operands rounded to tf32 (RNE), fp32 accumulate */
  MBV_PREC_BF16 = 2, /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate */
  MBV_PREC_FP16 = 3  /* tcgen05 kind::f16 (fp16 operands, saturating stores), fp32 accumulate.  Same tensor-core rate
                      * as bf16 with three more mantissa bits, range +-65504.  Because every 16-bit activation is then
                      * an fp16 tensor, the ResBlock / WaveNet residual streams are not stored separately: the residual
                      * add reads the fp16 operand tensor lrelu(x) that fed the block and inverts the leaky-relu
                      * ("single stream"), which removes a quarter of the ResBlock HBM traffic. */
} mbv_precision;

#define MBV_MAX_UPS 4
#define MBV_MAX_KERNELS 4
#define MBV_MAX_DILATIONS 3

/* Mirrors the `model` section of the reference configs (configs/ljs_mb_istft_vits.json:39-62) */
typedef struct {
  int32_t variant;                  /* mbv_variant */
  int32_t precision;                /* mbv_precision */
  int32_t inter_channels;           /* 192: latent channels (flow channels, decoder input) */
  int32_t hidden_channels;          /* flow WN width (192; 96 in the mini configs) */
  int32_t upsample_initial_channel; /* 512 / 256 */
  int32_t n_ups;
  int32_t upsample_rates[MBV_MAX_UPS];
  int32_t upsample_kernel_sizes[MBV_MAX_UPS];
  int32_t resblock_type;            /* 1 = ResBlock1 (modules.py:187), 2 = ResBlock2 (modules.py:237) */
  int32_t n_kernels;
  int32_t resblock_kernel_sizes[MBV_MAX_KERNELS];
  int32_t n_dilations;              /* 3 for ResBlock1, 2 for ResBlock2 */
  int32_t resblock_dilations[MBV_MAX_KERNELS][MBV_MAX_DILATIONS];
  int32_t n_fft;                    /* gen_istft_n_fft = 16 */
  int32_t hop;                      /* gen_istft_hop_size = 4 */
  int32_t subbands;                 /* 4 (MB/MS), 1 (iSTFT) */
  int32_t gin_channels;             /* 0 = no speaker conditioning */
  int32_t flow_kernel;              /* 5   } hard-coded at models.py:647 */
  int32_t flow_dilation_rate;       /* 1   } */
  int32_t flow_layers;              /* 4   } */
  int32_t flow_n;                   /* 4   } */
  int32_t device;                   /* CUDA device ordinal */
  int32_t flags;                    /* MBV_FLAG_* */
} mbv_config;

#define MBV_MAX_TEXT_LAYERS 12     /* text encoder depth (configs: n_layers 6; 3 in the mini configs) */
#define MBV_MAX_ENCQ_LAYERS 16     /* PosteriorEncoder WN depth (models.py:646) */

/* debug / tuning flags */
#define MBV_FLAG_FORCE_SIMT 4       /* run the CUDA-core conv on the tensor-core operand layout (cross-check) */
#define MBV_FLAG_RESIDUAL_FP16 8    /* bf16 path: keep the ResBlock residual stream and the WN hidden stream in (saturating) fp16 */
#define MBV_FLAG_CLUSTER_PAIRS 32    /* experimental: run the multi-tap convs as clusters of two CTAs that work on two time tiles of the
                                     * same weight group; each CTA fetches half of every weight tile and TMA-multicasts it to both.
                                     * Parity-tested; < 1 % faster per step on B200, hence opt-in. */
#define MBV_FLAG_NO_CTA_PAIRS 256   /* A/B and cross-check tests: do NOT run the multi-tap convs with an even number of channel tiles
                                     * (conv_pre, first upsampler, 256-channel ResBlocks) as cta_group::2 MMAs over CTA pairs
                                     * (one MMA spans two channel tiles, M = 256; each CTA stages half of the activation rows --
                                     * DESIGN.md section 4.1c).  Results are bit-identical either way. */
#define MBV_FLAG_NO_PW 512          /* A/B and cross-check tests: run the 1x1 convs of the WN stacks (coupling-layer pre, WN residual convs)
                                     * on the generic conv kernel (output channel on the accumulator lane) instead of pw_tc_kernel
                                     * (time on the lane, 256-bit epilogue accesses) */
#define MBV_FLAG_NO_PAIR_SPLIT 1024 /* A/B only: do not cut the leftover pair tiles of a CTA-pair launch's last round into column pieces */
#define MBV_FLAG_NO_PAIR_TM 2048   /* A/B and cross-check tests: run the k = 3 ResBlock1 conv pairs of a 128-channel stage as two launches instead of
                                     * pair_tm_kernel (one kernel, intermediate activation kept in shared memory) */
#define MBV_FLAG_NO_CONV_TM 4096    /* A/B and cross-check tests: run the k = 7 / k = 11 convs of a 128-channel ResBlock stage on the generic conv
                                     * kernel (single-CTA MMAs, channel on the lane) instead of conv_tm_kernel (time on the lane, cta_group::2
                                     * MMAs over CTA pairs that share every weight tile) */
#define MBV_FLAG_SPLIT_TAIL 128     /* keep conv_post as its own conv launch writing fp32 logits for the stand-alone tail kernel instead
                                     * of the fused conv_post + tail kernel (the default on the 16-bit paths; A/B and cross-check tests) */
#define MBV_FLAG_BRANCHES 64         /* experimental: run the parallel ResBlocks of a stage on the library's two extra streams (own
                                     * residual / operand buffers per branch) so that one branch's launches fill the drain and
                                     * partial last round of another's.  Results identical; measured neutral (9.32-9.35 vs
                                     * 9.33-9.56 ms per step, profiles/round2_branches_ab.txt), hence opt-in. */
#define MBV_FLAG_FUSED_PAIR 16      /* experimental: run each ResBlock1 conv pair of a 128-channel stage as ONE kernel that keeps
                                     * the intermediate activation in shared memory (conv_pair_kernel).  Parity-tested; currently
                                     * not faster than the two-launch path, hence opt-in. */

/* An EFFECTIVE weight tensor (weight-norm already folded: w = g*v/||v||), fp32, contiguous,
 * in HOST memory, named as in the reference state-dict minus weight_g/weight_v, e.g.
 * "dec.conv_pre.weight" [512,192,7], "dec.ups.0.weight" [C_in,C_out,16], "flow.flows.0.pre.bias". */
typedef struct {
  const char* name;
  const float* data;
  int32_t rank;
  int64_t shape[4];
} mbv_tensor;

typedef struct mbv_handle mbv_handle;

int mbv_abi_version(void);

/* Validate the geometry and create a handle bound to cfg->device. */
int mbv_create(const mbv_config* cfg, mbv_handle** out);
void mbv_destroy(mbv_handle* h);

/* Pack (transpose to tap-major/K-major, pad, fold the four Flips into pre/post, round to the
 * operand type) and upload all weights.  Must be called once before any compute call; every tensor
 * the configuration needs must be present. */
int mbv_load_weights(mbv_handle* h, const mbv_tensor* tensors, int32_t n);

/* Scratch the caller must provide for a batch of B utterances padded to T latent frames. */
int mbv_workspace_bytes(mbv_handle* h, int32_t B, int32_t T, size_t* bytes);

/* ResidualCouplingBlock.forward(x, x_mask, g, reverse=True)  (models.py:207-214).
 * z_p, z_out: [B, inter, T];  y_mask: [B, 1, T];  g: [B, gin, 1] or NULL.  z_out may alias z_p. */
int mbv_flow_reverse(mbv_handle* h, const float* z_p, const float* y_mask, const float* g,
                     float* z_out, int32_t B, int32_t T, void* ws, size_t ws_bytes, void* stream);

/* NEXT-row widening (the flow half of voice conversion, models.py:790-798):
 * ResidualCouplingBlock.forward(x, x_mask, g, reverse=False) (models.py:207-210) -- RCL0, Flip, RCL1, Flip, ... with
 * the mean-only coupling x1 <- m + x1 * mask (modules.py:345-347); the log-determinant is discarded as in the reference.
 * Same tensors and workspace as mbv_flow_reverse; mbv_flow_reverse(mbv_flow_forward(x)) == x * mask up to rounding. */
int mbv_flow_forward(mbv_handle* h, const float* x, const float* y_mask, const float* g, float* z_out,
                     int32_t B, int32_t T, void* ws, size_t ws_bytes, void* stream);

/* NEXT-row widening (completes voice conversion, models.py:790-798): PosteriorEncoder.forward
 * (models.py:236-246): x = pre(spec) * mask; x = WN(x, mask, g) (16 layers, models.py:646); stats = proj(x) * mask;
 * m, logs = split(stats); z = (m + noise * exp(logs)) * mask.  Available when mbv_load_weights received the enc_q.* tensors
 * (enc_q.pre / enc_q.enc.in_layers.N / enc_q.enc.res_skip_layers.N / enc_q.enc.cond_layer / enc_q.proj), else
 * MBV_ERR_WEIGHTS.  spec: [B, spec_channels (513), T]; y_mask: [B,1,T] (the caller builds it from the lengths, as
 * commons.sequence_mask does); g: [B,gin,1] or NULL; noise: [B,inter,T] (the reference draws torch.randn_like(m));
 * z: [B,inter,T]; stats: [B, 2*inter, T] = m | logs.  Its own workspace size: mbv_posterior_workspace_bytes. */
int mbv_posterior_workspace_bytes(mbv_handle* h, int32_t B, int32_t T, size_t* bytes);
int mbv_posterior_encode(mbv_handle* h, const float* spec, const float* y_mask, const float* g, const float* noise,
                         float* z, float* stats, int32_t B, int32_t T, void* ws, size_t ws_bytes, void* stream);

/* NEXT-row widening: TextEncoder.forward (models.py:172-181) -- embedding * sqrt(H), n_layers x
 * [relative-position windowed self-attention (attentions.py:101-254, window 4, shared embeddings), LayerNorm, k=3 conv
 * FFN with ReLU (attentions.py:257-303), LayerNorm], proj.  Available when mbv_load_weights received the enc_p.* tensors
 * (geometry is read off their shapes), else MBV_ERR_WEIGHTS.  tokens: [B, Tx] int64 (device); x_mask: [B,1,Tx] (the
 * caller builds it from the lengths like commons.sequence_mask); x_out: [B, hidden, Tx]; stats: [B, 2*inter, Tx] = m | logs.
 * The projections and FFN convs run on the conv kernels at the handle's precision; attention and LayerNorm in fp32. */
int mbv_text_workspace_bytes(mbv_handle* h, int32_t B, int32_t Tx, size_t* bytes);
int mbv_text_encode(mbv_handle* h, const int64_t* tokens, const float* x_mask, float* x_out, float* stats, int32_t B,
                    int32_t Tx, void* ws, size_t ws_bytes, void* stream);

/* dec.forward(z, g) (models.py:278-297 / 344-377 / 430-467).  z: [B, inter, T]; wav: [B,1,S*T]
 * with S = samples per latent frame (256).  Optional outputs (NULL to skip):
 *   o_mb : MB [B,4,64T];  MS [B,4,256T] (the zero-stuffed tensor the reference returns); iSTFT: must be NULL
 *   spec, phase : [B,4,9,16T+1] (MB/MS) or [B,9,64T+1] (iSTFT)
 * z_mask (optional, [B,1,T]) is multiplied into z on load, fusing the `z * y_mask` of models.py:734. */
int mbv_decode(mbv_handle* h, const float* z, const float* z_mask, const float* g, float* wav,
               float* o_mb, float* spec, float* phase, int32_t B, int32_t T, void* ws,
               size_t ws_bytes, void* stream);

/* The tail of SynthesizerTrn.infer in one call (models.py:730-734): flow reverse, mask, decode.
 * z_out (optional) receives the flow output like the reference's returned z. */
int mbv_flow_decode(mbv_handle* h, const float* z_p, const float* y_mask, const float* g,
                    float* z_out, float* wav, float* o_mb, float* spec, float* phase, int32_t B,
                    int32_t T, void* ws, size_t ws_bytes, void* stream);

/* NEXT-row widening: exact streaming decode.  The decoders are convolutional with a finite receptive
 * field (mbv_receptive_field latent frames per side: 25 for the MB / MS geometry, 13 single-band), so a stream of latent
 * chunks can be decoded with bounded latency and a result that equals the one-shot mbv_decode bit for bit -- unlike the
 * overlap-add chunking of the reference notebooks (infer.ipynb cells 4-6), which ignores the receptive field.
 *   mbv_stream_open   device state for B parallel utterances and chunks of <= max_chunk_frames latent frames: the latent
 *                     history (left halo + frames whose right halo has not arrived yet) and the decoded window
 *   mbv_stream_push   appends z_chunk [B, inter, n_frames] (device, already multiplied by the mask like dec's input) and
 *                     writes the samples of every frame whose right context is now complete -- all remaining frames when
 *                     `last` -- to wav_out [B, 1, 256 * n_frames_out] (device; capacity in frames given).  *first_frame /
 *                     *n_frames_out say which frames were emitted (n_frames_out may be 0).  Latency = halo frames.
 *                     ws: mbv_stream_workspace_bytes.  Asynchronous on `stream`; the state is not re-entrant.
 * PCM chunking (tts_vits.py:204-226) sits on top: mbv_pcm16 on the emitted samples, 20 ms slices on the host. */
typedef struct mbv_stream mbv_stream;
int mbv_receptive_field(mbv_handle* h);
int mbv_stream_open(mbv_handle* h, int32_t B, int32_t max_chunk_frames, mbv_stream** out);
int mbv_stream_workspace_bytes(mbv_stream* s, size_t* bytes);
int mbv_stream_halo(mbv_stream* s);
int mbv_stream_push(mbv_stream* s, const float* z_chunk, int32_t n_frames, int32_t last, const float* g, float* wav_out,
                    int32_t wav_capacity_frames, int64_t* first_frame, int32_t* n_frames_out, void* ws, size_t ws_bytes,
                    void* stream);
void mbv_stream_close(mbv_stream* s);

/* Introspection for benchmarks: number of kernel launches the last compute call enqueued, and the
 * algorithmic FLOPs (2*MACs of the dense contractions) of a decode / flow call at (B,T). */
int mbv_last_launch_count(mbv_handle* h);
double mbv_decode_flops(mbv_handle* h, int32_t B, int32_t T);
double mbv_flow_flops(mbv_handle* h, int32_t B, int32_t T);

/* Per-kernel device timing.  With profiling on, every launch of a compute call is bracketed by CUDA events on
 * the caller's stream.  mbv_profile_read synchronises those events, adds the elapsed milliseconds and launch
 * counts per kernel kind (0 = conv implicit-GEMM, 1 = fused tail, 2 = layout/GEMV helpers) into ms[3] / count[3],
 * and clears the record. */
int mbv_set_profiling(mbv_handle* h, int32_t on);
int mbv_profile_read(mbv_handle* h, double* ms, int32_t* count);
/* Same record, launch by launch (in launch order): elapsed ms and a short description ("conv m<epilogue mode>
 * Ci<in> N<rows> k<taps> d<dil> ph<phases> L<rows out> nt<tile width>", "tail", "") of up to cap launches; *n = how
 * many were written.  Clears the record. */
int mbv_profile_read_launches(mbv_handle* h, float* ms, char* desc, int32_t desc_stride, int32_t cap, int32_t* n);

/* Stand-alone entry for the fused tail (head + iSTFT + sub-band synthesis) on caller-provided
 * logits [B, F, n_logit_channels] (channels-last, F = frames): used by the parity tests and the
 * HBM-roofline benchmark of that kernel. */
int mbv_tail(mbv_handle* h, const float* logits, float* wav, float* o_mb, float* spec, float* phase,
             int32_t B, int32_t T, void* stream);

/* The fused conv_post + tail kernel alone (benchmarks / tests): act = the 16-bit operand tensor [B, 16T+1, C] that
 * conv_post consumes (device; bf16 or fp16 per the handle's precision) -> same outputs as mbv_tail.  16-bit precisions and
 * 4-band decoders only (else MBV_ERR_UNSUPPORTED). */
int mbv_tail_fused(mbv_handle* h, const void* act, float* wav, float* o_mb, float* spec, float* phase, int32_t B, int32_t T,
                   void* stream);

/* NEXT-row widening: the waveform post-processing of the reference's TTS service
 * (tts_vits.py:204-216): per utterance, peak-normalise to 0.9 if auto_normalize and the peak exceeds 0.01, clip to
 * [-1, 1], multiply by 32767 and truncate to int16.  wav: [B][stride] fp32 (device), n_samples: [B] valid sample
 * counts (device int32) or NULL = stride, scratch: >= 4*B bytes (device), pcm: [B][stride] int16 (device; samples
 * past n_samples[b] are written as 0).  Bit-exact with the reference's numpy arithmetic. */
int mbv_pcm16(mbv_handle* h, const float* wav, const int32_t* n_samples, int32_t B, int32_t stride, int32_t auto_normalize,
              void* scratch, int16_t* pcm, void* stream);

/* NEXT-row widening: the alignment expansion and prior sampling of SynthesizerTrn.infer,
 * models.py:717-729 (commons.generate_path, commons.py:128-143), as one gather kernel instead of a [B,1,Ty,Tx] attention
 * matrix and two batched matmuls:
 *   y_lengths = clamp_min(sum(w_ceil), 1);  y_mask[b,0,ty] = ty < y_lengths[b]
 *   tx(ty) = the token with cumsum(w_ceil)[tx-1] <= ty < cumsum(w_ceil)[tx]  (none for padded frames / masked tokens)
 *   z_p[b,c,ty] = m_p[b,c,tx] + noise[b,c,ty] * exp(logs_p[b,c,tx]) * noise_scale       (m = logs = 0 when there is no tx)
 * m_p, logs_p: [B,C,Tx];  w_ceil: [B,1,Tx] = ceil(exp(logw) * x_mask * length_scale);  x_mask: [B,1,Tx] or NULL;
 * noise: [B,C,Ty] standard normal drawn by the caller (the reference's torch.randn_like(m_p));  Ty = max(y_lengths),
 * computed by the caller exactly as the reference does (it needs the value on the host to size its tensors).
 * Outputs: z_p [B,C,Ty], y_mask [B,1,Ty]; optional (NULL to skip) m_exp / logs_exp [B,C,Ty] (the expanded statistics
 * infer() returns), attn [B,1,Ty,Tx], y_lengths [B] int64.  All device memory, fp32, contiguous. */
int mbv_expand_prior(mbv_handle* h, const float* m_p, const float* logs_p, const float* w_ceil, const float* x_mask,
                     const float* noise, float noise_scale, int32_t B, int32_t C, int32_t Tx, int32_t Ty, float* z_p,
                     float* y_mask, float* m_exp, float* logs_exp, float* attn, int64_t* y_lengths, void* stream);

const char* mbv_last_error(mbv_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MBISTFT_H_ */
