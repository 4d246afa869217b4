// common.cuh -- shared device-side definitions: operand types, the conv argument block and the fused
// epilogues used by both the tcgen05 implicit-GEMM conv kernel and the CUDA-core fp32 conv kernel.
//
// Layout convention inside the library (not at the ABI): every activation is channels-last
// [B][rows][Cp] with Cp = channels padded to a multiple of 64; "operand" copies hold the element type
// the tensor cores consume (bf16 / tf32-rounded fp32 / fp32), residual streams are always fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbv {

// ------------------------------------------------------------------------------------------------
// operand element types
// ------------------------------------------------------------------------------------------------
struct OpF32 { using T = float; static constexpr int kPrec = 0; };
struct OpTF32 { using T = float; static constexpr int kPrec = 1; };
struct OpBF16 { using T = __nv_bfloat16; static constexpr int kPrec = 2; };
struct OpF16 { using T = __half; static constexpr int kPrec = 3; };   // kind::f16 with fp16 operands (saturating stores)

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <typename Op> __device__ __forceinline__ float op_round(float x);
template <> __device__ __forceinline__ float op_round<OpF32>(float x) { return x; }
template <> __device__ __forceinline__ float op_round<OpTF32>(float x) { return round_tf32(x); }
template <> __device__ __forceinline__ float op_round<OpBF16>(float x) { return x; }
template <> __device__ __forceinline__ float op_round<OpF16>(float x) { return x; }

template <typename Op> __device__ __forceinline__ float op_load(const typename Op::T* p) { return *p; }
template <> __device__ __forceinline__ float op_load<OpBF16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float op_load<OpF16>(const __half* p) { return __half2float(*p); }

// fp16 stores saturate: one F2FP.SATFINITE, +-inf / overflow clamp to +-65504
__device__ __forceinline__ __half to_half_sat(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return __ushort_as_half(h);
}

// store W (multiple of 8) consecutive operand elements, 16-byte vectorised
template <typename Op, int W>
__device__ __forceinline__ void op_store_vec(typename Op::T* dst, const float* v) {
  if constexpr (Op::kPrec == 3) {
#pragma unroll
    for (int i = 0; i < W; i += 8) {
      uint4 u;
      __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h2[j] = __halves2half2(to_half_sat(v[i + 2 * j]), to_half_sat(v[i + 2 * j + 1]));
      *reinterpret_cast<uint4*>(dst + i) = u;
    }
  } else if constexpr (Op::kPrec == 2) {
#pragma unroll
    for (int i = 0; i < W; i += 8) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v[i + 0], v[i + 1]);
      __nv_bfloat162 b = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
      __nv_bfloat162 c = __floats2bfloat162_rn(v[i + 4], v[i + 5]);
      __nv_bfloat162 d = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      u.z = *reinterpret_cast<uint32_t*>(&c);
      u.w = *reinterpret_cast<uint32_t*>(&d);
      *reinterpret_cast<uint4*>(dst + i) = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; i += 4) {
      float4 f = make_float4(op_round<Op>(v[i]), op_round<Op>(v[i + 1]), op_round<Op>(v[i + 2]),
                             op_round<Op>(v[i + 3]));
      *reinterpret_cast<float4*>(dst + i) = f;
    }
  }
}

template <int W> __device__ __forceinline__ void f32_load_vec(float* v, const float* src) {
#pragma unroll
  for (int i = 0; i < W; i += 4) {
    float4 f = *reinterpret_cast<const float4*>(src + i);
    v[i] = f.x; v[i + 1] = f.y; v[i + 2] = f.z; v[i + 3] = f.w;
  }
}
template <int W> __device__ __forceinline__ void f32_store_vec(float* dst, const float* v) {
#pragma unroll
  for (int i = 0; i < W; i += 4)
    *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

// fp16 residual stream: saturating round-to-nearest store (to_half_sat above), widening load
template <int W> __device__ __forceinline__ void res_load_vec(float* v, const void* base, size_t off, int half) {
  if (half) {
    const __half* p = reinterpret_cast<const __half*>(base) + off;
#pragma unroll
    for (int i = 0; i < W; i += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(p + i);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h2[j]); v[i + 2 * j] = f.x; v[i + 2 * j + 1] = f.y; }
    }
  } else {
    f32_load_vec<W>(v, reinterpret_cast<const float*>(base) + off);
  }
}
template <int W> __device__ __forceinline__ void res_store_vec(void* base, size_t off, const float* v, int half) {
  if (half) {
    __half* p = reinterpret_cast<__half*>(base) + off;
#pragma unroll
    for (int i = 0; i < W; i += 8) {
      uint4 u;
      __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h2[j] = __halves2half2(to_half_sat(v[i + 2 * j]), to_half_sat(v[i + 2 * j + 1]));
      *reinterpret_cast<uint4*>(p + i) = u;
    }
  } else {
    f32_store_vec<W>(reinterpret_cast<float*>(base) + off, v);
  }
}

// ------------------------------------------------------------------------------------------------
// epilogues
// ------------------------------------------------------------------------------------------------
enum EpiMode : int {
  EPI_ACT = 0,   // y = acc + bias [ * mask ]; optional fp32 copy; operand copies act[j] = lrelu(y + add_j)
  EPI_RES = 1,   // x = xin + acc + bias; optional xout / operand copy / resblock-sum handling
  EPI_F32 = 2,   // out = acc + bias (fp32, pitch ld, only n < n_valid)
  EPI_GATE = 3,  // act = tanh(acc + bias + g) * sigmoid(acc2 + bias' + g')      (commons.py:100-107)
  EPI_RS = 4,    // WN res/skip update (modules.py:169-175)
  EPI_POST = 5   // coupling update z[:, off+n] = (z - (acc+bias)*mask)*mask      (modules.py:338-352)
};

struct EpiParams {
  int mode;
  int n_valid;      // logical output channels >= n_valid are never stored (packed rows are padded to 128)
  int ld;           // channel pitch (elements) of every output / residual buffer of this conv
  int rows_out;     // rows per utterance of the mapped outputs (act[], xout when mapped)
  int rows_res;     // rows per utterance of xin / xout / xs / mask
  int row_mul;      // mapped row = row * row_mul + row_add + phase
  int row_add;
  int dup_src;      // if mapped row == dup_src also store to mapped row dup_dst (ReflectionPad1d((1,0)))
  int dup_dst;
  const float* bias; int bias_bs;     // bias[b * bias_bs + n]
  const float* add2; int add2_bs;     // EPI_GATE: per-utterance conditioning (cond_layer(g)) or null
  const float* mask;                  // [B][rows_res] or null
  // residual-stream buffers: fp32, or fp16 when res_half (decoder ResBlock stream with bf16 operands: the stream is
  // re-rounded once per residual add, 2^-11 relative, invisible next to the 2^-9 operand rounding -- DESIGN.md 3)
  const void* xin;
  void* xout;                         // residual / plain output
  void* xs;                           // resblock running sum
  int res_half;                       // 0 fp32, 1 plain fp16, 2 single stream: xin is an fp16 operand tensor holding
                                      // lrelu(x) (inverted with inv_slope on load), no separate residual output
  float inv_slope;
  void* act[3];
  const float* act_add[3]; int act_add_bs;  // per-utterance addend before lrelu (ResBlock cond(g))
  int n_act;
  float slope;      // leaky-relu slope of the operand copies (1 = identity)
  float scale;      // 1 / num_kernels for the final resblock-sum
  int sum_mode;     // EPI_RES: 0 none, 1 xs = x, 2 xs += x, 3 final: v = (xs + x) * scale, 4 final of a single resblock: v = x * scale
  int n_split;      // EPI_RS: width of the residual half (0 = last layer)
  int ch_off;       // EPI_POST: first channel updated; EPI_GATE: first channel of this layer's slot in the acts buffer
  int first;        // EPI_RS: first WN layer (skip accumulator is set, not added)
  float post_sign;  // EPI_POST: +1 reverse (x1 - m), -1 forward direction (x1 + m)
};

// One row, W consecutive output columns starting at n0.  acc/acc2 are modified in place.
template <typename Op, int W>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, int b, int row, int phase, int n0,
                                               float* acc, float* acc2) {
  using T = typename Op::T;
  float tmp[W];
  {
    // packed weight rows are padded to a multiple of 128: skip chunks outside the destination's channels
    const int c0 = (p.mode == EPI_RS && p.n_split > 0 && n0 >= p.n_split) ? n0 - p.n_split : n0;
    if (c0 >= p.n_valid) return;
  }
  // bias (indexed by packed weight row; the gate packs 128-row tiles of [64 tanh | 64 sigmoid] rows)
  const int brow = (p.mode == EPI_GATE) ? 128 * (n0 >> 6) + (n0 & 63) : n0;
  {
    const float* bp = p.bias + (size_t)b * p.bias_bs + brow;
    f32_load_vec<W>(tmp, bp);
#pragma unroll
    for (int i = 0; i < W; ++i) acc[i] += tmp[i];
  }
  const int mrow = row * p.row_mul + p.row_add + phase;
  const size_t res_off = ((size_t)b * p.rows_res + row) * p.ld + n0;
  const size_t map_off = ((size_t)b * p.rows_out + mrow) * p.ld + n0;
  const float m = p.mask ? p.mask[(size_t)b * p.rows_res + row] : 1.f;

  switch (p.mode) {
    case EPI_ACT: {
      if (p.mask) {
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] *= m;
      }
      if (p.xout) res_store_vec<W>(p.xout, map_off, acc, p.res_half);
      for (int j = 0; j < p.n_act; ++j) {
        if (p.act_add[j]) {
          f32_load_vec<W>(tmp, p.act_add[j] + (size_t)b * p.act_add_bs + n0);
#pragma unroll
          for (int i = 0; i < W; ++i) tmp[i] = lrelu(acc[i] + tmp[i], p.slope);
        } else {
#pragma unroll
          for (int i = 0; i < W; ++i) tmp[i] = lrelu(acc[i], p.slope);
        }
        op_store_vec<Op, W>(reinterpret_cast<T*>(p.act[j]) + map_off, tmp);
      }
    } break;
    case EPI_RES: {
      res_load_vec<W>(tmp, p.xin, res_off, p.res_half);
#pragma unroll
      for (int i = 0; i < W; ++i) acc[i] += tmp[i];
      if (p.xout) res_store_vec<W>(p.xout, res_off, acc, p.res_half);
      if (p.sum_mode == 1) {
        res_store_vec<W>(p.xs, res_off, acc, p.res_half);
      } else if (p.sum_mode == 2) {
        res_load_vec<W>(tmp, p.xs, res_off, p.res_half);
#pragma unroll
        for (int i = 0; i < W; ++i) tmp[i] += acc[i];
        res_store_vec<W>(p.xs, res_off, tmp, p.res_half);
      } else if (p.sum_mode == 3) {
        res_load_vec<W>(tmp, p.xs, res_off, p.res_half);
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] = (tmp[i] + acc[i]) * p.scale;
      } else if (p.sum_mode == 4) {
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] *= p.scale;
      }
      if (p.n_act) {
#pragma unroll
        for (int i = 0; i < W; ++i) tmp[i] = lrelu(acc[i], p.slope);
        T* dst = reinterpret_cast<T*>(p.act[0]);
        op_store_vec<Op, W>(dst + map_off, tmp);
        if (mrow == p.dup_src)
          op_store_vec<Op, W>(dst + ((size_t)b * p.rows_out + p.dup_dst) * p.ld + n0, tmp);
      }
    } break;
    case EPI_F32: {
      float* dst = reinterpret_cast<float*>(p.xout) + ((size_t)b * p.rows_out + mrow) * p.ld + n0;
      if (n0 + W <= p.n_valid && (p.ld & 3) == 0) {
        f32_store_vec<W>(dst, acc);
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i)
          if (n0 + i < p.n_valid) dst[i] = acc[i];
      }
    } break;
    case EPI_GATE: {
      f32_load_vec<W>(tmp, p.bias + (size_t)b * p.bias_bs + brow + 64);
#pragma unroll
      for (int i = 0; i < W; ++i) acc2[i] += tmp[i];
      if (p.add2) {
        const float* gp = p.add2 + (size_t)b * p.add2_bs + brow;
        f32_load_vec<W>(tmp, gp);
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] += tmp[i];
        f32_load_vec<W>(tmp, gp + 64);
#pragma unroll
        for (int i = 0; i < W; ++i) acc2[i] += tmp[i];
      }
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const float t = tanhf(acc[i]);
        const float s = 1.f / (1.f + __expf(-acc2[i]));
        tmp[i] = t * s;
      }
      op_store_vec<Op, W>(reinterpret_cast<T*>(p.act[0]) + map_off + p.ch_off, tmp);
    } break;
    case EPI_RS: {
      if (p.n_split > 0 && n0 < p.n_split) {
        // residual half: x = (x + rs) * mask -> fp32 stream + operand copy for the next in_layer
        f32_load_vec<W>(tmp, reinterpret_cast<const float*>(p.xin) + res_off);
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] = (acc[i] + tmp[i]) * m;
        f32_store_vec<W>(reinterpret_cast<float*>(p.xout) + res_off, acc);
        op_store_vec<Op, W>(reinterpret_cast<T*>(p.act[0]) + map_off, acc);
      } else {
        // skip half: output += rs; the last layer also applies the mask and emits the operand copy
        const size_t so = ((size_t)b * p.rows_res + row) * p.ld + (n0 - p.n_split);
        if (!p.first) {
          f32_load_vec<W>(tmp, reinterpret_cast<const float*>(p.xs) + so);
#pragma unroll
          for (int i = 0; i < W; ++i) acc[i] += tmp[i];
        }
        if (p.n_split > 0) {
          f32_store_vec<W>(reinterpret_cast<float*>(p.xs) + so, acc);
        } else {
#pragma unroll
          for (int i = 0; i < W; ++i) acc[i] *= m;
          op_store_vec<Op, W>(reinterpret_cast<T*>(p.act[0]) + so, acc);
        }
      }
    } break;
    case EPI_POST: {
      const size_t zo = ((size_t)b * p.rows_res + row) * p.ld + p.ch_off + n0;
      f32_load_vec<W>(tmp, reinterpret_cast<const float*>(p.xin) + zo);
#pragma unroll
      for (int i = 0; i < W; ++i) acc[i] = (tmp[i] - p.post_sign * (acc[i] * m)) * m;
      f32_store_vec<W>(reinterpret_cast<float*>(p.xout) + zo, acc);
      op_store_vec<Op, W>(reinterpret_cast<T*>(p.act[0]) + zo, acc);
    } break;
    default: break;
  }
}

// ------------------------------------------------------------------------------------------------
// conv argument block (shared by both conv kernels)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPhases = 8;

struct ConvArgs {
  const void* x;      // operand activations [B][L_in][x_ld]; the conv reads channels [0, Cp_in) of every row
  int x_ld;           // channel pitch of x in elements (>= Cp_in)
  const void* w;      // packed weights [phase][tap][N_total][Cp_in] (operand type, K-major)
  int B;
  int L_in;           // rows per utterance of x
  int L_out;          // rows computed per utterance and phase
  int Cp_in;          // padded input channels (multiple of 64)
  int N_total;        // packed weight rows per tap (multiple of 128)
  int taps;
  int dil;            // row step between taps
  int n_phases;       // >1 for the polyphase transposed convolutions
  int shift0[kMaxPhases];  // row offset of tap 0 per phase
  int gate;           // EPI_GATE: weight rows packed as 128-row tiles [64 tanh | 64 sigmoid] of 64 consecutive channels
  // fused ResBlock conv pair (conv_pair_kernel): w / taps / dil / shift0 describe c1, w2 the second conv (same taps,
  // dilation 1); bias_h / slope_h are c1's bias and the leaky-relu between the convs; epi is c2's RES epilogue
  const void* w2;
  const float* bias_h;
  float slope_h;
  EpiParams epi;
};

}  // namespace mbv
