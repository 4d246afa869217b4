// text.cu -- the small kernels of the text encoder (models.py:140-181, attentions.py:13-47): token embedding,
// relative-position windowed self-attention, and residual + LayerNorm.  The dense projections (q/k/v, o, the two FFN
// convs, proj) run on the conv kernels of conv_tc.cu / conv_simt.cu; everything here is per-token work on at most a
// few hundred tokens per utterance: fp32 CUDA-core arithmetic, one launch each, no tensor cores.
//
// All activations are channels-last [B][T][C] like the rest of the library; "operand" copies are written in the
// element type the conv kernels consume (prec: 0 fp32, 1 tf32-rounded fp32, 2 bf16, 3 fp16).
#include "common.cuh"
#include "kernels.h"

namespace mbv {

__device__ __forceinline__ void store_operand(void* base, size_t idx, float v, int prec) {
  if (prec == 3) reinterpret_cast<__half*>(base)[idx] = to_half_sat(v);
  else if (prec == 2) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else if (prec == 1) reinterpret_cast<float*>(base)[idx] = round_tf32(v);
  else reinterpret_cast<float*>(base)[idx] = v;
}

// x[b][t][:] = emb[token[b][t]][:] * sqrt(C) * mask[b][t]        (models.py:173-177, attentions.py:37)
__global__ void __launch_bounds__(256) text_embed_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb,
                                                         const float* __restrict__ mask, float* __restrict__ x, void* __restrict__ xop,
                                                         int n_rows, int C, int ld_op, int n_vocab, float scale, int prec) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  long long tok = tokens[row];
  if (tok < 0) tok = 0;
  if (tok >= n_vocab) tok = n_vocab - 1;
  const float m = mask[row] * scale;
  for (int c = threadIdx.x & 31; c < C; c += 32) {
    const float v = emb[(size_t)tok * C + c] * m;
    x[(size_t)row * C + c] = v;
    store_operand(xop, (size_t)row * ld_op + c, v, prec);
  }
}

cudaError_t launch_text_embed(const long long* tokens, const float* emb, const float* mask, float* x, void* xop, int n_rows, int C,
                              int ld_op, int n_vocab, int prec, cudaStream_t st) {
  text_embed_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(tokens, emb, mask, x, xop, n_rows, C, ld_op, n_vocab, sqrtf((float)C), prec);
  return cudaGetLastError();
}

// x_out = LayerNorm(x + y) over the channels (modules.py:29-33: mean / biased variance, eps inside the root); the operand
// copy is multiplied by the mask (the FFN and proj consume x * x_mask, attentions.py:281,285 / models.py:178) and, for the
// last layer, so is x_out (attentions.py:46).  One warp per token.
__global__ void __launch_bounds__(256) text_ln_kernel(const float* __restrict__ x, const float* __restrict__ y, int y_ld,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ mask, float* __restrict__ x_out,
                                                      void* __restrict__ xop, int n_rows, int C, int ld_op, int mask_out, int prec) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float v[16];  // C <= 512
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < C) ? x[(size_t)row * C + c] + y[(size_t)row * y_ld + c] : 0.f;
    s += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + 1e-5f);
  const float m = mask[row];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      const float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
      x_out[(size_t)row * C + c] = mask_out ? o * m : o;
      store_operand(xop, (size_t)row * ld_op + c, o * m, prec);
    }
  }
}

cudaError_t launch_text_ln(const float* x, const float* y, int y_ld, const float* gamma, const float* beta, const float* mask,
                           float* x_out, void* xop, int n_rows, int C, int ld_op, int mask_out, int prec, cudaStream_t st) {
  if (C > 512) return cudaErrorInvalidValue;
  text_ln_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(x, y, y_ld, gamma, beta, mask, x_out, xop, n_rows, C, ld_op, mask_out, prec);
  return cudaGetLastError();
}

// MultiHeadAttention.attention with window_size 4 and shared relative embeddings (attentions.py:142-178):
//   s[i][j] = q_i . k_j / sqrt(dk)  +  [|j - i| <= W] q_i . Ek[j - i + W] / sqrt(dk);   s = -1e4 where mask_i * mask_j == 0
//   p = softmax_j(s);   out_i = sum_j p[i][j] v_j  +  sum_{|r| <= W} p[i][i + r] Ev[r + W]
// One CTA = (16 queries, one head, one utterance); qkv is [B][T][3C] fp32 (q | k | v), out the operand tensor [B][T][C].
// Shared memory: the scaled query tile and the 16 x T score matrix.
constexpr int ATT_Q = 16;
constexpr int ATT_THREADS = 128;

__global__ void __launch_bounds__(ATT_THREADS) text_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                    const float* __restrict__ rel_k, const float* __restrict__ rel_v,
                                                                    void* __restrict__ out, int T, int C, int ld_op, int dk, int W, int prec) {
  extern __shared__ float sm[];
  float* sq = sm;                    // [ATT_Q][dk]
  float* ss = sm + ATT_Q * dk;       // [ATT_Q][T]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_Q;
  const int tid = threadIdx.x;
  const size_t ld = (size_t)3 * C;
  const float* base = qkv + (size_t)b * T * ld;
  const float scale = rsqrtf((float)dk);
  const int nq = min(ATT_Q, T - q0);
  for (int i = tid; i < ATT_Q * dk; i += ATT_THREADS) {
    const int qi = i / dk, d = i % dk;
    sq[i] = (qi < nq) ? base[(size_t)(q0 + qi) * ld + h * dk + d] * scale : 0.f;
  }
  __syncthreads();
  // ---- scores: thread = key j
  for (int j = tid; j < T; j += ATT_THREADS) {
    const float* kr = base + (size_t)j * ld + C + h * dk;
    float acc[ATT_Q];
#pragma unroll
    for (int qi = 0; qi < ATT_Q; ++qi) acc[qi] = 0.f;
    for (int d = 0; d < dk; d += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(kr + d);
#pragma unroll
      for (int qi = 0; qi < ATT_Q; ++qi) {
        const float4 qv = *reinterpret_cast<const float4*>(sq + qi * dk + d);
        acc[qi] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, fmaf(qv.z, kv.z, fmaf(qv.w, kv.w, acc[qi]))));
      }
    }
    const float mj = mask[(size_t)b * T + j];
#pragma unroll
    for (int qi = 0; qi < ATT_Q; ++qi) {
      if (qi < nq) {
        const int r = j - (q0 + qi);
        float s = acc[qi];
        if (r >= -W && r <= W) {
          const float* ek = rel_k + (size_t)(r + W) * dk;
          float e = 0.f;
          for (int d = 0; d < dk; ++d) e = fmaf(sq[qi * dk + d], ek[d], e);
          s += e;
        }
        if (mj * mask[(size_t)b * T + q0 + qi] == 0.f) s = -1e4f;
        ss[qi * T + j] = s;
      }
    }
  }
  __syncthreads();
  // ---- softmax per query row: warp w takes rows w, w + 4, ...
  const int warp = tid >> 5, lane = tid & 31;
  for (int qi = warp; qi < nq; qi += ATT_THREADS / 32) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, ss[qi * T + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = expf(ss[qi * T + j] - mx);
      ss[qi * T + j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int j = lane; j < T; j += 32) ss[qi * T + j] *= inv;
  }
  __syncthreads();
  // ---- out: thread = channel d of the head (dk <= 128)
  if (tid < dk) {
    float acc[ATT_Q];
#pragma unroll
    for (int qi = 0; qi < ATT_Q; ++qi) acc[qi] = 0.f;
    const float* vr = base + 2 * C + h * dk + tid;
    for (int j = 0; j < T; ++j) {
      const float vv = vr[(size_t)j * ld];
#pragma unroll
      for (int qi = 0; qi < ATT_Q; ++qi) acc[qi] = fmaf(ss[qi * T + j], vv, acc[qi]);
    }
#pragma unroll
    for (int qi = 0; qi < ATT_Q; ++qi) {
      if (qi < nq) {
        float o = acc[qi];
        for (int r = -W; r <= W; ++r) {
          const int j = q0 + qi + r;
          if (j >= 0 && j < T) o = fmaf(ss[qi * T + j], rel_v[(size_t)(r + W) * dk + tid], o);
        }
        store_operand(out, ((size_t)b * T + q0 + qi) * ld_op + h * dk + tid, o, prec);
      }
    }
  }
}

cudaError_t launch_text_attention(const float* qkv, const float* mask, const float* rel_k, const float* rel_v, void* out, int B,
                                  int T, int C, int ld_op, int n_heads, int W, int prec, cudaStream_t st) {
  const int dk = C / n_heads;
  if (dk > ATT_THREADS || (dk & 3) != 0) return cudaErrorInvalidValue;
  const size_t smem = sizeof(float) * ((size_t)ATT_Q * dk + (size_t)ATT_Q * T);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(text_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dim3 grid((T + ATT_Q - 1) / ATT_Q, n_heads, B);
  text_attention_kernel<<<grid, ATT_THREADS, smem, st>>>(qkv, mask, rel_k, rel_v, out, T, C, ld_op, dk, W, prec);
  return cudaGetLastError();
}

}  // namespace mbv
