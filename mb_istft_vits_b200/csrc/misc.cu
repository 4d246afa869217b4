// misc.cu -- layout conversion at the ABI boundary and the per-utterance conditioning GEMVs.
#include "common.cuh"
#include "kernels.h"

namespace mbv {

// fp32 NCT [B][C][T] (time contiguous, as the reference passes tensors) -> channels-last [B][T][Cp]:
// 32x32 smem transpose so both the read (along T) and the write (along C) are coalesced.  Pad channels
// C..Cp are written as zeros.  Optional mask [B][T] fuses the `z * y_mask` of models.py:734.
template <typename Op>
__global__ void __launch_bounds__(256) pack_input_kernel(const float* __restrict__ src, const float* __restrict__ mask,
                                                         typename Op::T* __restrict__ dst_op, float* __restrict__ dst_f32,
                                                         int C, int T, int Cp) {
  // 64 channels x 32 time steps per block: reads are 128-byte rows along T, writes are 64 consecutive channels of a time step
  // (two per thread: 128 bytes of 16-bit operands / 256 bytes of fp32 per warp instruction)
  __shared__ __align__(8) float tile[32][66];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int t_in = t0 + tx;
  const float m = (mask && t_in < T) ? mask[(size_t)b * T + t_in] : 1.f;
#pragma unroll
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i;
    float v = 0.f;
    if (c < C && t_in < T) {
      v = src[((size_t)b * C + c) * T + t_in];
      if (mask) v *= m;
    }
    tile[tx][i] = v;
  }
  __syncthreads();
  const int c = c0 + 2 * tx;   // Cp is a multiple of 64
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i;
    if (t < T && c < Cp) {
      const float2 v = *reinterpret_cast<const float2*>(&tile[i][2 * tx]);
      const size_t o = ((size_t)b * T + t) * Cp + c;
      if (dst_op) {
        if constexpr (Op::kPrec == 3) *reinterpret_cast<__half2*>(dst_op + o) = __halves2half2(to_half_sat(v.x), to_half_sat(v.y));
        else if constexpr (Op::kPrec == 2) *reinterpret_cast<__nv_bfloat162*>(dst_op + o) = __floats2bfloat162_rn(v.x, v.y);
        else *reinterpret_cast<float2*>(dst_op + o) = make_float2(op_round<Op>(v.x), op_round<Op>(v.y));
      }
      if (dst_f32) *reinterpret_cast<float2*>(dst_f32 + o) = v;
    }
  }
}

cudaError_t launch_pack_input(int prec, const float* src, const float* mask, void* dst_op, float* dst_f32, int B, int C,
                              int T, int Cp, cudaStream_t st) {
  dim3 grid((T + 31) / 32, (Cp + 63) / 64, B);
  if (prec == 3)
    pack_input_kernel<OpF16><<<grid, 256, 0, st>>>(src, mask, (__half*)dst_op, dst_f32, C, T, Cp);
  else if (prec == 2)
    pack_input_kernel<OpBF16><<<grid, 256, 0, st>>>(src, mask, (__nv_bfloat16*)dst_op, dst_f32, C, T, Cp);
  else if (prec == 1)
    pack_input_kernel<OpTF32><<<grid, 256, 0, st>>>(src, mask, (float*)dst_op, dst_f32, C, T, Cp);
  else
    pack_input_kernel<OpF32><<<grid, 256, 0, st>>>(src, mask, (float*)dst_op, dst_f32, C, T, Cp);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) unpack_output_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            int C, int T, int Cp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    tile[i][tx] = (t < T && c < C) ? src[((size_t)b * T + t) * Cp + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (c < C && t < T) dst[((size_t)b * C + c) * T + t] = tile[tx][i];
  }
}

cudaError_t launch_unpack_output(const float* src, float* dst, int B, int C, int T, int Cp, cudaStream_t st) {
  dim3 grid((T + 31) / 32, (C + 31) / 32, B);
  unpack_output_kernel<<<grid, 256, 0, st>>>(src, dst, C, T, Cp);
  return cudaGetLastError();
}

// out[b][n] = bias[n] (+ base[n]) + sum_c w[n][c] * g[b][c]: ResBlock.cond (modules.py:209-215) and
// WN.cond_layer (modules.py:126-128,152-153) act on g [B,gin,1], i.e. one GEMV per utterance.
__global__ void __launch_bounds__(128) cond_gemv_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ base,
                                                        float* __restrict__ out, int G, int N, int out_ld) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f;
  for (int c = lane; c < G; c += 32) s = fmaf(w[(size_t)n * G + c], g[(size_t)b * G + c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[(size_t)b * out_ld + n] = s + (bias ? bias[n] : 0.f) + (base ? base[n] : 0.f);
}

cudaError_t launch_cond_gemv(const float* g, const float* w, const float* bias, const float* base, float* out, int B,
                             int G, int N, int out_ld, cudaStream_t st) {
  dim3 grid((N + 3) / 4, B);
  cond_gemv_kernel<<<grid, 128, 0, st>>>(g, w, bias, base, out, G, N, out_ld);
  return cudaGetLastError();
}

// PosteriorEncoder sampling (models.py:243-245): m, logs = split(stats); z = (m + noise * exp(logs)) * mask, all NCT.
__global__ void __launch_bounds__(256) posterior_sample_kernel(const float* __restrict__ stats, const float* __restrict__ noise,
                                                               const float* __restrict__ mask, float* __restrict__ z, int C, int T) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  const float m = stats[((size_t)b * 2 * C + c) * T + t];
  const float logs = stats[((size_t)b * 2 * C + C + c) * T + t];
  z[((size_t)b * C + c) * T + t] = (m + noise[((size_t)b * C + c) * T + t] * expf(logs)) * mask[(size_t)b * T + t];
}

cudaError_t launch_posterior_sample(const float* stats, const float* noise, const float* mask, float* z, int B, int C, int T,
                                    cudaStream_t st) {
  dim3 grid((T + 255) / 256, C, B);
  posterior_sample_kernel<<<grid, 256, 0, st>>>(stats, noise, mask, z, C, T);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Waveform post-processing of the reference's TTS service (tts_vits.py:204-216): per-utterance peak normalisation to
// 0.9 (only if the peak exceeds 0.01), clip to [-1, 1], scale by 32767 and truncate to int16.  Two passes: a per-
// utterance max |x| (warp shuffle -> one atomicMax on the float bit pattern per CTA), then the conversion.
// Every float operation is a single correctly-rounded IEEE op in the reference's order, so the PCM is bit-exact.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pcm_peak_kernel(const float* __restrict__ wav, const int* __restrict__ n_valid,
                                                       int stride, unsigned int* __restrict__ peak_bits) {
  const int b = blockIdx.y;
  const int n = n_valid ? n_valid[b] : stride;
  const float* x = wav + (size_t)b * stride;
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, wm[i]);
    atomicMax(peak_bits + b, __float_as_uint(m));  // non-negative floats order like their bit patterns
  }
}

__global__ void __launch_bounds__(256) pcm_convert_kernel(const float* __restrict__ wav, const int* __restrict__ n_valid,
                                                          int stride, const unsigned int* __restrict__ peak_bits,
                                                          int auto_normalize, short* __restrict__ pcm) {
  const int b = blockIdx.y;
  const int n = n_valid ? n_valid[b] : stride;
  const float peak = __uint_as_float(peak_bits[b]);
  const bool norm = auto_normalize && peak > 0.01f;
  const float* x = wav + (size_t)b * stride;
  short* o = pcm + (size_t)b * stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < stride; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < n) {
      v = x[i];
      if (norm) v = __fmul_rn(__fdiv_rn(v, peak), 0.9f);
      v = fminf(fmaxf(v, -1.f), 1.f);
      v = __fmul_rn(v, 32767.f);
    }
    o[i] = (short)(int)v;  // truncation toward zero, like numpy's astype(int16)
  }
}

cudaError_t launch_pcm16(const float* wav, const int* n_valid, int B, int stride, int auto_normalize, unsigned int* peak_bits,
                         short* pcm, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(peak_bits, 0, sizeof(unsigned int) * B, st);
  if (e != cudaSuccess) return e;
  int gx = (stride + 256 * 8 - 1) / (256 * 8);
  if (gx > 1024) gx = 1024;
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  pcm_peak_kernel<<<grid, 256, 0, st>>>(wav, n_valid, stride, peak_bits);
  pcm_convert_kernel<<<grid, 256, 0, st>>>(wav, n_valid, stride, peak_bits, auto_normalize, pcm);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// NEXT-row widening: alignment expansion + prior sampling of SynthesizerTrn.infer
// (models.py:717-729, commons.generate_path commons.py:128-143).  The reference materialises attn [B,1,Ty,Tx] and runs
// two batched matmuls to gather rows; attn has at most one 1 per output frame, so this is a gather:
//   tx(ty) = first token whose cumulative duration exceeds ty;   m, logs = m_p[:, tx], logs_p[:, tx]  (0 if none / padded)
//   z_p = m + noise * exp(logs) * noise_scale;   y_mask = ty < max(sum(w_ceil), 1)
// One CTA = one utterance x 128 output frames; every CTA redoes the (tiny, exact: integer-valued fp32) prefix sum.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) expand_prior_kernel(const float* __restrict__ m_p, const float* __restrict__ logs_p,
                                                           const float* __restrict__ w_ceil, const float* __restrict__ x_mask,
                                                           const float* __restrict__ noise, float noise_scale, int C, int Tx,
                                                           int Ty, float* __restrict__ z_p, float* __restrict__ y_mask,
                                                           float* __restrict__ m_exp, float* __restrict__ logs_exp,
                                                           float* __restrict__ attn, long long* __restrict__ y_lengths) {
  extern __shared__ float s_cum[];     // [Tx] inclusive prefix sums, then s_tx[128]
  int* s_tx = reinterpret_cast<int*>(s_cum + Tx);
  const int b = blockIdx.y, ty0 = blockIdx.x * 128, tid = threadIdx.x;
  if (tid == 0) {
    float acc = 0.f;
    const float* w = w_ceil + (size_t)b * Tx;
    for (int i = 0; i < Tx; ++i) { acc += w[i]; s_cum[i] = acc; }
  }
  __syncthreads();
  const float total = s_cum[Tx - 1];
  const int y_len = total < 1.f ? 1 : (int)total;   // clamp_min(sum, 1).long()
  if (y_lengths && blockIdx.x == 0 && tid == 0) y_lengths[b] = y_len;
  const int ty = ty0 + tid;
  int tx = Tx;
  if (ty < Ty) {
    // smallest tx with cum[tx] > ty  (path = [ty < cum[tx]] - [ty < cum[tx-1]], commons.py:139-141)
    int lo = 0, hi = Tx;
    const float fty = (float)ty;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_cum[mid] > fty) hi = mid; else lo = mid + 1;
    }
    tx = lo;
  }
  bool ok = (ty < Ty) && (tx < Tx) && (ty < y_len);
  if (ok && x_mask) ok = x_mask[(size_t)b * Tx + tx] != 0.f;
  s_tx[tid] = ok ? tx : -1;
  if (ty < Ty) {
    y_mask[(size_t)b * Ty + ty] = ty < y_len ? 1.f : 0.f;
    for (int c = 0; c < C; ++c) {
      const size_t src = ((size_t)b * C + c) * Tx + tx, dst = ((size_t)b * C + c) * Ty + ty;
      const float m = ok ? m_p[src] : 0.f, l = ok ? logs_p[src] : 0.f;
      // m + ((noise * exp(logs)) * noise_scale), every product / sum rounded separately like the reference's tensor ops
      z_p[dst] = __fadd_rn(m, __fmul_rn(__fmul_rn(noise[dst], expf(l)), noise_scale));
      if (m_exp) m_exp[dst] = m;
      if (logs_exp) logs_exp[dst] = l;
    }
  }
  if (attn) {
    __syncthreads();
    const int rows = min(128, Ty - ty0);
    float* a = attn + ((size_t)b * Ty + ty0) * Tx;
    for (int e = tid; e < rows * Tx; e += 128) {
      const int rr = e / Tx, cc = e - rr * Tx;
      a[e] = (s_tx[rr] == cc) ? 1.f : 0.f;
    }
  }
}

cudaError_t launch_expand_prior(const float* m_p, const float* logs_p, const float* w_ceil, const float* x_mask,
                                const float* noise, float noise_scale, int B, int C, int Tx, int Ty, float* z_p,
                                float* y_mask, float* m_exp, float* logs_exp, float* attn, long long* y_lengths,
                                cudaStream_t st) {
  dim3 grid((Ty + 127) / 128, B);
  const size_t smem = (size_t)Tx * 4 + 128 * 4;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  expand_prior_kernel<<<grid, 128, smem, st>>>(m_p, logs_p, w_ceil, x_mask, noise, noise_scale, C, Tx, Ty, z_p, y_mask, m_exp,
                                               logs_exp, attn, y_lengths);
  return cudaGetLastError();
}

}  // namespace mbv
