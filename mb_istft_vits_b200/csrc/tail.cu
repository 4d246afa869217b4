// tail.cu -- fused decoder tail: exp / pi*sin head -> 16-point inverse real DFT (registers) -> periodic Hann
// window -> overlap-add (hop 4) with the exact edge envelope -> {nothing | 4-band synthesis FIR with the x4
// zero-stuffing folded into a polyphase form} -> fp32 waveform.  HBM-bound by design: per latent frame it reads
// 16 frames x 72 logits x 4 B = 4608 B and writes 256 samples x 4 B = 1024 B.
//
// Reference semantics: models.py:366-375 (head, reshape), stft.py:197-202 (torch.istft,
// n_fft 16, hop 4, center=True, periodic Hann), pqmf.py:105-116 (zero-stuff x4 with gain 4, pad 31, 63-tap
// cross-correlation), models.py:463-465 (MS: same with the trainable multistream_conv_post).
//
// One CTA = one tile of one utterance: 256 consecutive STFT frames per band (one per thread) give 253 hop
// blocks of sub-band signal, of which 249 are owned outputs (the FIR needs +-2 hop blocks of halo):
// 97 % useful work.  Shared memory: logits tile 72 KB + frame scratch 20 KB + sub-band tile 16 KB -> 2 CTAs/SM,
// so one CTA's loads overlap the other's math.
#include "common.cuh"
#include "kernels.h"

namespace mbv {

constexpr int TAIL_THREADS = 256;
constexpr int TAIL_NF = 256;        // frames per band per tile
constexpr int FR_PITCH = 20;        // floats per frame row in smem (16 + pad: conflict-free float4)
constexpr int YMB_PITCH = 1024;     // floats per band in the sub-band tile (253 * 4 = 1012 used)

// cos(2*pi*m/16)
__device__ constexpr float kCos16[16] = {
    1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
    0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
    -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
    0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};

// periodic Hann / 16 (the irfft normalisation folded in): w[n] = 0.5 - 0.5 cos(2 pi n / 16)
__device__ constexpr float kWin16[16] = {
    0.0f / 16, 0.03806023374435663f / 16, 0.14644660940672624f / 16, 0.30865828381745514f / 16,
    0.5f / 16, 0.69134171618254486f / 16, 0.85355339059327376f / 16, 0.96193976625564337f / 16,
    1.0f / 16, 0.96193976625564337f / 16, 0.85355339059327376f / 16, 0.69134171618254486f / 16,
    0.5f / 16, 0.30865828381745514f / 16, 0.14644660940672624f / 16, 0.03806023374435663f / 16};
// squared window (envelope terms)
__device__ __forceinline__ float win_sq(int n) {
  const float w = kWin16[n] * 16.f;
  return w * w;
}

template <bool PRECISE>
__device__ __forceinline__ void head(float xm, float xp, float& mag, float& ph, float& re, float& im) {
  // spec = exp(x), phase = pi * sin(x)  (models.py:368-369); re/im = spec * (cos, sin)(phase)
  float s, c;
  if (PRECISE) {
    mag = expf(xm);
    ph = 3.14159265358979323846f * sinf(xp);
    sincosf(ph, &s, &c);
  } else {
    mag = __expf(xm);
    // one Cody-Waite step keeps the fast sine inside its accurate range for any logit
    const float k = rintf(xp * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, xp);
    r = fmaf(-k, -1.7484555e-7f, r);
    ph = 3.14159265358979323846f * __sinf(r);
    __sincosf(ph, &s, &c);
  }
  re = mag * c;
  im = mag * s;
}

// VARIANT: 0 = single-band iSTFT (no synthesis filter), 1 = MB / MS (4 bands + 63-tap synthesis FIR)
template <int VARIANT, bool PRECISE>
__global__ void __launch_bounds__(TAIL_THREADS, 2) tail_kernel(const __grid_constant__ TailArgs a, int tiles_per_utt) {
  constexpr int S = VARIANT == 0 ? 1 : 4;
  constexpr int NQ = VARIANT == 0 ? TAIL_NF - 3 : TAIL_NF - 7;  // owned hop blocks per tile
  constexpr int YOFF = VARIANT == 0 ? 0 : 2;                    // y-block index of the first owned block
  constexpr int NCH = S * 18;
  extern __shared__ __align__(16) float sm[];
  float* s_log = sm;                         // [TAIL_NF][NCH]
  float* s_fr = s_log + TAIL_NF * NCH;       // [TAIL_NF][FR_PITCH]
  float* s_y = s_fr + TAIL_NF * FR_PITCH;    // [S][YMB_PITCH]

  const int t = threadIdx.x;
  const int b = blockIdx.x / tiles_per_utt;
  const int tile = blockIdx.x % tiles_per_utt;
  const int L = a.L;            // hop blocks per band (= frames - 1)
  const int F = L + 1;
  const int Q0 = tile * NQ;     // first owned hop block
  const int QY0 = Q0 - YOFF;    // first y block held in smem
  const int F0 = QY0 - 1;       // first frame held in smem
  const bool last_tile = (tile == tiles_per_utt - 1);

  // ---- phase 0: coalesced copy of the logits rows [F0, F0+256) /\ [0, F) into smem
  {
    const int f_lo = F0 < 0 ? 0 : F0;
    const int f_hi = (F0 + TAIL_NF < F) ? F0 + TAIL_NF : F;
    const size_t g0 = ((size_t)b * F + f_lo) * NCH;
    const int n = (f_hi - f_lo) * NCH;
    float* dst = s_log + (f_lo - F0) * NCH;
    const float* src = a.logits + g0;
    if ((g0 & 3) == 0 && (((f_lo - F0) * NCH) & 3) == 0) {
      const int n4 = n >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = t; i < n4; i += TAIL_THREADS) d4[i] = __ldg(s4 + i);
      for (int i = (n4 << 2) + t; i < n; i += TAIL_THREADS) dst[i] = __ldg(src + i);
    } else {
      for (int i = t; i < n; i += TAIL_THREADS) dst[i] = __ldg(src + i);
    }
  }
  __syncthreads();

  const int f = F0 + t;  // this thread's frame
  const bool f_valid = (f >= 0 && f < F);
  // frames whose spec/phase this tile writes: the owned blocks' frames, plus the final frame L for the last tile
  const bool f_owned = f_valid && (f >= Q0) && (f < Q0 + NQ || (last_tile && f == L));

#pragma unroll 1
  for (int s = 0; s < S; ++s) {
    // ---- phase 1: head + inverse DFT + window for frame f, band s
    float fr[16];
    if (f_valid) {
      const float2* lp = reinterpret_cast<const float2*>(s_log + t * NCH + s * 18);
      float x[18];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float2 v = lp[i];
        x[2 * i] = v.x;
        x[2 * i + 1] = v.y;
      }
      float re[9], im[9];
      const bool emit = (a.spec != nullptr) && f_owned;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float mag, ph;
        head<PRECISE>(x[k], x[9 + k], mag, ph, re[k], im[k]);
        if (emit) {
          // spec/phase are [B][S][9][F]; consecutive threads = consecutive frames -> coalesced
          const size_t o = (((size_t)b * S + s) * 9 + k) * F + f;
          a.spec[o] = mag;
          a.phase[o] = ph;
        }
      }
      // x[n] = Re0 + (-1)^n Re8 + 2 sum_{k=1..7} (Re_k cos(2 pi k n/16) - Im_k sin(2 pi k n/16)); imag of bins 0, 8 ignored
      fr[0] = 0.f;  // w[0] == 0
#pragma unroll
      for (int n = 1; n <= 8; ++n) {
        float c = 0.f, sn = 0.f;
#pragma unroll
        for (int k = 1; k <= 7; ++k) {
          c = fmaf(re[k], kCos16[(k * n) & 15], c);
          sn = fmaf(im[k], kCos16[(k * n + 12) & 15], sn);
        }
        const float base = re[0] + ((n & 1) ? -re[8] : re[8]);
        fr[n] = (base + 2.f * (c - sn)) * kWin16[n];
        if (n < 8) fr[16 - n] = (base + 2.f * (c + sn)) * kWin16[16 - n];
      }
    } else {
#pragma unroll
      for (int n = 0; n < 16; ++n) fr[n] = 0.f;
    }
    {
      float4* dst = reinterpret_cast<float4*>(s_fr + t * FR_PITCH);
      dst[0] = make_float4(fr[0], fr[1], fr[2], fr[3]);
      dst[1] = make_float4(fr[4], fr[5], fr[6], fr[7]);
      dst[2] = make_float4(fr[8], fr[9], fr[10], fr[11]);
      dst[3] = make_float4(fr[12], fr[13], fr[14], fr[15]);
    }
    __syncthreads();
    // ---- phase 2: overlap-add.  y block q (sub-band samples 4q..4q+3) = frame q-1 part 3 + q part 2 + q+1 part 1 + q+2 part 0
    if (t < TAIL_NF - 3) {
      const int q = QY0 + t;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q >= 0 && q < L) {
        const float4 p3 = *reinterpret_cast<const float4*>(s_fr + (t + 0) * FR_PITCH + 12);
        const float4 p2 = *reinterpret_cast<const float4*>(s_fr + (t + 1) * FR_PITCH + 8);
        const float4 p1 = *reinterpret_cast<const float4*>(s_fr + (t + 2) * FR_PITCH + 4);
        const float4 p0 = *reinterpret_cast<const float4*>(s_fr + (t + 3) * FR_PITCH + 0);
        y.x = p3.x + p2.x + p1.x + p0.x;
        y.y = p3.y + p2.y + p1.y + p0.y;
        y.z = p3.z + p2.z + p1.z + p0.z;
        y.w = p3.w + p2.w + p1.w + p0.w;
        // window-square envelope: 1.5 in steady state; frame -1 (q == 0) and frame F (q == L-1) do not exist
        float e0 = 1.5f, e1 = 1.5f, e2 = 1.5f, e3 = 1.5f;
        if (q == 0) { e0 -= win_sq(12); e1 -= win_sq(13); e2 -= win_sq(14); e3 -= win_sq(15); }
        if (q == L - 1) { e0 -= win_sq(0); e1 -= win_sq(1); e2 -= win_sq(2); e3 -= win_sq(3); }
        y.x /= e0; y.y /= e1; y.z /= e2; y.w /= e3;
        const bool owned = (t >= YOFF) && (t < YOFF + NQ);
        if (VARIANT == 0) {
          if (owned) *reinterpret_cast<float4*>(a.wav + (size_t)b * 4 * L + 4 * (size_t)q) = y;
        } else if (a.o_mb != nullptr && owned) {
          if (a.variant == 1) {  // MB: y_mb_hat [B][S][4L]
            *reinterpret_cast<float4*>(a.o_mb + ((size_t)b * S + s) * 4 * L + 4 * (size_t)q) = y;
          } else {  // MS: the zero-stuffed tensor [B][S][16L], gain 4 (models.py:463)
            float4* o = reinterpret_cast<float4*>(a.o_mb + ((size_t)b * S + s) * 16 * L + 16 * (size_t)q);
            o[0] = make_float4(4.f * y.x, 0.f, 0.f, 0.f);
            o[1] = make_float4(4.f * y.y, 0.f, 0.f, 0.f);
            o[2] = make_float4(4.f * y.z, 0.f, 0.f, 0.f);
            o[3] = make_float4(4.f * y.w, 0.f, 0.f, 0.f);
          }
        }
      }
      if (VARIANT != 0) *reinterpret_cast<float4*>(s_y + s * YMB_PITCH + 4 * t) = y;
    }
    __syncthreads();
  }

  if (VARIANT == 0) return;

  // ---- phase 3: polyphase synthesis FIR.  Thread t owns hop block Q0+t: sub-band positions j = 4(Q0+t)+e,
  // outputs n = 4j + r.  out[4j+r] = sum_c sum_{d=-7..8} G[c][r][d] * y[c][j+d],  G = 4*h[c][4d+31-r] (0 if outside).
  if (t < NQ && Q0 + t < L) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[20];  // y[c][4t .. 4t+19] (local), position j+d -> index 8 + e + d
      const float4* yp = reinterpret_cast<const float4*>(s_y + c * YMB_PITCH + 4 * t);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float4 u = yp[i];
        v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int d = 0; d < 16; ++d)  // d - 7 in [-7, 8]
            acc[4 * e + r] = fmaf(a.coef[c][r * 16 + d], v[8 + e + d - 7], acc[4 * e + r]);
    }
    float4* o = reinterpret_cast<float4*>(a.wav + (size_t)b * 16 * L + 16 * (size_t)(Q0 + t));
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    o[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
    o[3] = make_float4(acc[12], acc[13], acc[14], acc[15]);
  }
}

template <int VARIANT, bool PRECISE>
static cudaError_t launch_tail_t(const TailArgs& a, cudaStream_t st) {
  constexpr int S = VARIANT == 0 ? 1 : 4;
  constexpr int NQ = VARIANT == 0 ? TAIL_NF - 3 : TAIL_NF - 7;
  const int tiles = (a.L + NQ - 1) / NQ;
  const size_t smem = sizeof(float) * ((size_t)TAIL_NF * S * 18 + (size_t)TAIL_NF * FR_PITCH + (size_t)S * YMB_PITCH);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tail_kernel<VARIANT, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  tail_kernel<VARIANT, PRECISE><<<a.B * tiles, TAIL_THREADS, smem, st>>>(a, tiles);
  return cudaGetLastError();
}

cudaError_t launch_tail(const TailArgs& a, int precise, cudaStream_t st) {
  if (a.variant == 0) return precise ? launch_tail_t<0, true>(a, st) : launch_tail_t<0, false>(a, st);
  return precise ? launch_tail_t<1, true>(a, st) : launch_tail_t<1, false>(a, st);
}

}  // namespace mbv
