// conv_tc.cu -- implicit-GEMM Conv1d / polyphase ConvTranspose1d on the 5th-gen tensor cores (sm_100a).
//
//   D[C_out 128, time N] (+)= sum_tap sum_kblock  W_tap[C_out 128, 64ch] * X_tap[time N, 64ch]^T        (N <= 256)
//
// * activations are channels-last, so BOTH operands are K-major: A = one tap of the packed weight (128 output
//   channels x 64 input channels), B = an N-row time tile of the activation, rows shifted by the tap offset.
//   Putting the output CHANNEL on the accumulator lane (M) means an epilogue warp's 32 threads own 32 consecutive
//   channels of one time step: every global load / store of the fused epilogue is a fully coalesced 64-128 B
//   access on the channels-last tensors, and per-channel constants (bias, conditioning) are registers.
//   (v1 had time on M: each thread owned a row and every warp access touched 32 different lines -- the ncu
//   summary in profiles/r01_ncu_full_v1_baseline.txt shows 32 sectors/request and an LSU-bound epilogue.)
// * TMA (cp.async.bulk.tensor, 128B swizzle) stages operands; per k-block ONE activation slab with the halo of all
//   taps (N + (taps-1)*dil rows) is loaded and every tap's MMA reads it through a row-offset shared-memory
//   descriptor, so activations cross L2->SMEM once, not `taps` times.  Rows outside the utterance (= the conv's
//   zero padding) are zero-filled by TMA through a 3-D (C, time, utterance) tensor map.
// * tcgen05.mma (cta_group::1, M=128, N<=256, kind::f16 bf16 or kind::tf32) accumulates in TMEM; two accumulator
//   stages let the epilogue of tile i overlap the MMAs of tile i+1.
// * warp roles: warps 0-7 = epilogue (tcgen05.ld -> registers -> fused bias / residual / leaky-relu / gate / mask ->
//   coalesced global access), warp 8 = TMA producer, warp 9 = TMEM allocator + MMA issuer.  The two single-thread
//   roles get the HIGHEST warp ids: the issue arbiter favours high warp ids, and the instruction-heavy epilogue
//   warps must never delay an MMA or TMA issue.
// * persistent: grid = min(#tiles, #SMs), static round-robin tile order (channel tile fastest so neighbouring CTAs
//   share the activation slab in L2).
//
// Reference semantics implemented: F.conv1d 'same' (commons.py:14-15, modules.py:191-206,220-224),
// F.conv_transpose1d as S polyphase branches (models.py:320-323), WN gate (commons.py:100-107).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../include/mbistft.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace mbv {

// Epilogue warps per CTA.  The fused epilogues are latency-bound (per-tile timelines: IPC ~0.3 per scheduler with two
// warps each), so every mode but the gate (whose warp pairs exchange through 32 KB of shared memory) runs FOUR warps
// per TMEM lane quarter; with 18 warps the register budget is 112 per thread, which is why the residual input is no
// longer double-buffered in registers (an L2 prefetch one tile ahead plus the extra warps hide its latency instead).
template <int MODE> struct EpiWarps { static constexpr int value = (MODE == EPI_GATE || MODE == EPI_ACT) ? 8 : 12; };
template <int MODE> struct TcThreads { static constexpr int value = 64 + 32 * EpiWarps<MODE>::value; };  // epilogue + producer + MMA
constexpr int TC_ACC_STRIDE = 256;   // TMEM columns per accumulator stage

template <typename Op> __device__ __forceinline__ void op_store1(typename Op::T* p, float v) { *p = op_round<Op>(v); }
// (Measured and dropped: st.global.cg for these 16-bit epilogue stores -- L1 bypass, which helped the 256-bit row-per-thread
//  accesses of the time-on-lane kernels a lot -- changes nothing here: 8.59 vs 8.59 ms per step, interleaved builds.)
template <> __device__ __forceinline__ void op_store1<OpBF16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void op_store1<OpF16>(__half* p, float v) { *p = to_half_sat(v); }

// residual-stream element.  RH = 0: fp32;  1: saturating fp16;  2: "single stream" (fp16 operands only): the residual
// input IS the fp16 operand tensor lrelu(x) that fed the block's first conv -- leaky-relu is invertible, x = v >= 0 ? v :
// v / slope -- and the only output is the next operand tensor, so a ResBlock conv pair moves one 2-byte tensor in and one
// out instead of two and two.
template <int RH> struct ResT { using T = __half; };
template <> struct ResT<0> { using T = float; };
template <int RH> __device__ __forceinline__ float res_ld(const typename ResT<RH>::T* p) {
  if constexpr (RH != 0) return __half2float(*p); else return *p;
}
template <int RH> __device__ __forceinline__ void res_st(typename ResT<RH>::T* p, float v) {
  if constexpr (RH != 0) *p = to_half_sat(v); else *p = v;
}

// MUFU.TANH: one instruction, max abs error 2^-11 -- below the 2^-9 rounding the bf16 operand copy applies anyway.
__device__ __forceinline__ float mufu_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <typename Op> __device__ __forceinline__ float gate_tanh(float x);
template <typename Op> __device__ __forceinline__ float gate_sigmoid(float x);

__device__ __forceinline__ float fast_tanh(float x) {
  // tanh(x) = 1 - 2 / (exp(2x) + 1); exact limits at +-inf, abs error ~1e-7 relative to the fp32 reference
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}

template <> __device__ __forceinline__ float gate_tanh<OpBF16>(float x) { return mufu_tanh(x); }
template <> __device__ __forceinline__ float gate_sigmoid<OpBF16>(float x) { return fmaf(mufu_tanh(0.5f * x), 0.5f, 0.5f); }
template <> __device__ __forceinline__ float gate_tanh<OpF16>(float x) { return mufu_tanh(x); }
template <> __device__ __forceinline__ float gate_sigmoid<OpF16>(float x) { return fmaf(mufu_tanh(0.5f * x), 0.5f, 0.5f); }
template <> __device__ __forceinline__ float gate_tanh<OpTF32>(float x) { return fast_tanh(x); }
template <> __device__ __forceinline__ float gate_sigmoid<OpTF32>(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

struct TcRt {  // runtime scalars the kernel needs beyond ConvArgs
  int n_time;        // time columns per tile (UMMA N), multiple of 16, <= 256 (gate: <= 128)
  int slab_rows;     // n_time + (taps-1)*dil
  int box_rows;      // rows per TMA box of the slab
  int n_boxes;       // 1 or 2
  int slab_stage_bytes, w_stage_bytes, n_slab_stages, n_w_stages;
  int t_tiles, c_tiles, total_tiles;
  int xchg_off;      // byte offset of the gate exchange buffer in dynamic smem
  int w_resident;    // all weight tiles of the (single) channel tile fit the ring: load them once per CTA
  int prefetch_res;  // issue an L2 prefetch of the tile's residual input (tmR) when its operand loads start
  int cluster;       // 1: launched as clusters of 2 CTAs that work on two time tiles of the SAME (phase, channel tile) in
                     // lock step; every weight tile is fetched once per pair (each CTA loads half and multicasts it)
  int rows, groups;  // cluster mode: B * t_tiles time tiles, n_phases * c_tiles weight groups
  // Tail-wave split: the persistent grid's last, partial round (total_tiles % grid tiles) would keep most SMs idle for a
  // whole tile time; those tiles are cut into split_k column pieces of split_n columns, one per CTA.  Virtual tile
  // indices >= split_from address the pieces.  (Results are bit-identical: a column's accumulation does not depend on
  // the tile width.)
  int split_from, split_k, split_n, virt_tiles;
  int pair_phase;    // CTA pairs (CL == 2): the pair is two polyphase branches with equal input shift instead of two channel tiles
  int res_off;       // > 0: byte offset of the per-warp residual staging rings (two 2 KB stages per epilogue warp): the 2-byte
                     // residual input of chunk i+1 is fetched by TMA while chunk i is processed (EPI_RES / EPI_RS, RH != 0)
  int pack2;         // EPI_ACT, 2-byte streams: lane pairs trade halves and store 4-byte channel pairs (MBV_NO_PACK2=1 clears it: A/B only)
  int pdl;           // launch with programmatic stream serialization
  int rotate;        // two channel tiles, even grid: swap which one a CTA takes every round.  With a static round-robin
                     // an even CTA would otherwise ALWAYS get channel tile 0; when the second tile is half padding
                     // (192 = 128 + 64 rows: flow pre / res convs) its epilogue is half the work and half the CTAs idle.
  long long* dbg;    // MBV_TIMELINE=1: per-CTA clock stamps [cta][role 0..2][tile][2] (debug only)
};

// ------------------------------------------------------------------------------------------------
// Fused epilogues.  One thread = one output channel n (weight row) x 32 consecutive time steps held in registers.
// MODE and the channel pitch are compile-time so that each kernel carries only its own epilogue and every row
// offset (i * pitch) folds into the load/store immediate: no per-element address arithmetic, no branches.
//   STEP > 0: compile-time row pitch in elements;  STEP == 0: runtime pitch `rstep`.
//   FULL: all 32 time steps are inside the utterance (the common case), else the first `nt` are.
// ------------------------------------------------------------------------------------------------
#define MBV_EL(i) if (FULL || (i) < nt)

template <typename Op, int STEP, bool FULL, int RH>
__device__ __forceinline__ void epi_act(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                        size_t rstep, const float* acc) {
  using T = typename Op::T;
  const size_t step = STEP > 0 ? (size_t)STEP : rstep;
  const float bias = p.bias[(size_t)b * p.bias_bs + n];
  const size_t off0 = ((size_t)b * p.rows_out + (size_t)t_first * p.row_mul + p.row_add + phase) * p.ld + n;
  float y[32];
  if (p.mask) {
    const float* mp = p.mask + (size_t)b * p.rows_res + t_first;
#pragma unroll
    for (int i = 0; i < 32; ++i) { y[i] = acc[i] + bias; MBV_EL(i) y[i] *= mp[i]; }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = acc[i] + bias;
  }
  if (p.xout) {
    typename ResT<RH>::T* xo = reinterpret_cast<typename ResT<RH>::T*>(p.xout) + off0;
#pragma unroll
    for (int i = 0; i < 32; ++i) MBV_EL(i) res_st<RH>(xo + i * step, y[i]);
  }
  const float slope = p.slope;
  for (int j = 0; j < p.n_act; ++j) {
    const float* addp = j == 0 ? p.act_add[0] : (j == 1 ? p.act_add[1] : p.act_add[2]);
    void* actp = j == 0 ? p.act[0] : (j == 1 ? p.act[1] : p.act[2]);
    const float add = addp ? addp[(size_t)b * p.act_add_bs + n] : 0.f;
    T* dst = reinterpret_cast<T*>(actp) + off0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float v = y[i] + add;
      MBV_EL(i) op_store1<Op>(dst + i * step, fmaxf(v, v * slope));  // leaky-relu, slope in (0,1]
    }
  }
}

// ---- packed 2-byte stores (round 2).  With the channel on the lane a thread stores its channel for 32 time steps, one
// 2-byte element per instruction: 1024 warp-level stores per 128 x 256 tile and output stream.  Per-tile stamps with the
// epilogue's stores switched off (profiles/round2_timeline_k3_stage0.txt) show what they cost the MMAs, which fetch their
// operands through the same L1 / shared-memory data path: a stage-0 k = 3 c1 tile takes 9.87 K cycles with the scalar
// stores and 7.8 K without any (k = 11: 23.7 K against 22.5 K = exactly 128 cycles per MMA); the TMEM loads cost nothing.
// Two adjacent channels (lanes 2j, 2j+1) therefore trade halves of their time steps: the even lane keeps steps 0..15 of
// BOTH channels, the odd lane 16..31, and each stores 16 four-byte words (channel pair): half the store instructions,
// the same bytes and the same values (the conversions are the scalar path's, element by element): 9.87 K -> 9.44 K cycles
// per k = 3 tile.  EPI_ACT only: the same exchange in the residual-add epilogue (12 warps, 128 registers) spilled ~80 bytes.
template <int KIND> __device__ __forceinline__ uint32_t pack2_16(float lo, float hi) {
  uint32_t d;
  if constexpr (KIND == 0) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// row0_even: address of (first row of the chunk, channel n & ~1) in a tensor of 2-byte elements; step: row pitch in elements
template <int KIND, int STEP, bool FULL>
__device__ __forceinline__ void store_pairs16(void* row0_even, size_t rstep, int nt, int odd, const float* v) {
  const size_t step = STEP > 0 ? (size_t)STEP : rstep;
  uint32_t own[8], oth[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t w_lo = pack2_16<KIND>(v[2 * k], v[2 * k + 1]);            // time steps 2k, 2k+1
    const uint32_t w_hi = pack2_16<KIND>(v[16 + 2 * k], v[16 + 2 * k + 1]);  // time steps 16+2k, 16+2k+1
    own[k] = odd ? w_hi : w_lo;
    oth[k] = __shfl_xor_sync(0xffffffffu, odd ? w_lo : w_hi, 1);             // the partner channel's words for MY time steps
  }
  const int r0 = odd ? 16 : 0;
  char* base = reinterpret_cast<char*>(row0_even) + (size_t)r0 * step * 2;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t ev = odd ? oth[k] : own[k], od = odd ? own[k] : oth[k];   // even channel -> low half, odd channel -> high half
    if (FULL || r0 + 2 * k < nt) *reinterpret_cast<uint32_t*>(base + (size_t)(2 * k) * step * 2) = __byte_perm(ev, od, 0x5410);
    if (FULL || r0 + 2 * k + 1 < nt) *reinterpret_cast<uint32_t*>(base + (size_t)(2 * k + 1) * step * 2) = __byte_perm(ev, od, 0x7632);
  }
}

// epi_act with packed stores: the whole warp is inside the destination's channels (warp-uniform, checked by the caller)
template <typename Op, int STEP, bool FULL, int RH>
__device__ __forceinline__ void epi_act_packed(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                               size_t rstep, const float* acc) {
  static_assert(sizeof(typename Op::T) == 2, "packed stores: 2-byte operand types");
  constexpr int KIND = (Op::kPrec == 2) ? 0 : 1;
  const int odd = n & 1;
  const float bias = p.bias[(size_t)b * p.bias_bs + n];
  const size_t off_even = ((size_t)b * p.rows_out + (size_t)t_first * p.row_mul + p.row_add + phase) * p.ld + (n - odd);
  float y[32];
  if (p.mask) {
    const float* mp = p.mask + (size_t)b * p.rows_res + t_first;
#pragma unroll
    for (int i = 0; i < 32; ++i) { y[i] = acc[i] + bias; MBV_EL(i) y[i] *= mp[i]; }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = acc[i] + bias;
  }
  if (p.xout) {
    if constexpr (RH != 0) {
      store_pairs16<1, STEP, FULL>(reinterpret_cast<__half*>(p.xout) + off_even, rstep, nt, odd, y);
    } else {
      const size_t step = STEP > 0 ? (size_t)STEP : rstep;
      float* xo = reinterpret_cast<float*>(p.xout) + off_even + odd;
#pragma unroll
      for (int i = 0; i < 32; ++i) MBV_EL(i) xo[i * step] = y[i];
    }
  }
  const float slope = p.slope;
  for (int j = 0; j < p.n_act; ++j) {
    const float* addp = j == 0 ? p.act_add[0] : (j == 1 ? p.act_add[1] : p.act_add[2]);
    void* actp = j == 0 ? p.act[0] : (j == 1 ? p.act[1] : p.act[2]);
    const float add = addp ? addp[(size_t)b * p.act_add_bs + n] : 0.f;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float t = y[i] + add; v[i] = fmaxf(t, t * slope); }
    store_pairs16<KIND, STEP, FULL>(reinterpret_cast<typename Op::T*>(actp) + off_even, rstep, nt, odd, v);
  }
}

template <typename Op, int LD, int RH>
__device__ __forceinline__ void tc_epilogue_act_packed(const EpiParams& p, int b, int n, int phase, int t_first, int nt, const float* acc) {
  const size_t rstep = (size_t)p.row_mul * p.ld;
  if (LD > 0 && p.row_mul == 1) {
    if (nt == 32) epi_act_packed<Op, LD, true, RH>(p, b, n, phase, t_first, nt, rstep, acc);
    else epi_act_packed<Op, LD, false, RH>(p, b, n, phase, t_first, nt, rstep, acc);
  } else {
    if (nt == 32) epi_act_packed<Op, 0, true, RH>(p, b, n, phase, t_first, nt, rstep, acc);
    else epi_act_packed<Op, 0, false, RH>(p, b, n, phase, t_first, nt, rstep, acc);
  }
}

// FOLDED: the caller (staged residual path of conv_tc_kernel) has already inverted the leaky-relu of a single-stream
// residual and added the running ResBlock sum xs into xpre.
template <typename Op, int STEP, bool FULL, int RH, bool FOLDED>
__device__ __forceinline__ void epi_res(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                        size_t rstep, const float* acc, const float* xpre) {
  using T = typename Op::T;
  const size_t step = STEP > 0 ? (size_t)STEP : rstep;
  const float bias = p.bias[(size_t)b * p.bias_bs + n];
  const size_t base = ((size_t)b * p.rows_res + t_first) * p.ld + n;
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = xpre[i];  // xin, prefetched while the MMAs of this tile were running
  if constexpr (RH == 2 && !FOLDED) {
    const float inv = p.inv_slope;
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = fminf(x[i], x[i] * inv);  // inverse leaky-relu (inv_slope >= 1): v < 0 -> v / slope
  }
  const int sm = p.sum_mode;
  if (!FOLDED && (sm == 2 || sm == 3)) {
    float sv[32];
    const typename ResT<RH>::T* sp = reinterpret_cast<const typename ResT<RH>::T*>(p.xs) + base;
#pragma unroll
    for (int i = 0; i < 32; ++i) { sv[i] = 0.f; MBV_EL(i) sv[i] = res_ld<RH>(sp + i * step); }
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (x[i] + sv[i]) + acc[i] + bias;  // same order as the staged path: bit-identical
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = x[i] + acc[i] + bias;
  }
  if (p.xout) {
    typename ResT<RH>::T* xo = reinterpret_cast<typename ResT<RH>::T*>(p.xout) + base;
#pragma unroll
    for (int i = 0; i < 32; ++i) MBV_EL(i) res_st<RH>(xo + i * step, x[i]);
  }
  if (sm == 1 || sm == 2) {
    typename ResT<RH>::T* so = reinterpret_cast<typename ResT<RH>::T*>(p.xs) + base;
#pragma unroll
    for (int i = 0; i < 32; ++i) MBV_EL(i) res_st<RH>(so + i * step, x[i]);
  }
  if (p.n_act) {
    const float slope = p.slope, scale = (sm >= 3) ? p.scale : 1.f;
    T* dst = reinterpret_cast<T*>(p.act[0]) + ((size_t)b * p.rows_out + t_first + p.row_add) * p.ld + n;
    if (sm >= 3) {  // mean over the parallel ResBlocks (models.py:361)
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] *= scale;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) MBV_EL(i) op_store1<Op>(dst + i * step, fmaxf(x[i], x[i] * slope));
    // ReflectionPad1d((1,0)) of the conv_post input: mapped row dup_src is also stored at row dup_dst
    const int di = p.dup_src - p.row_add - t_first;
    if (p.dup_src >= 0 && di >= 0 && di < nt) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) if (i == di) v = x[i];
      op_store1<Op>(reinterpret_cast<T*>(p.act[0]) + ((size_t)b * p.rows_out + p.dup_dst) * p.ld + n, fmaxf(v, v * slope));
    }
  }
}

template <typename Op, int STEP, bool FULL>
__device__ __forceinline__ void epi_f32(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                        size_t rstep, const float* acc) {
  const size_t step = STEP > 0 ? (size_t)STEP : rstep;
  const float bias = p.bias[(size_t)b * p.bias_bs + n];
  float* o = reinterpret_cast<float*>(p.xout) + ((size_t)b * p.rows_out + t_first) * p.ld + n;
#pragma unroll
  for (int i = 0; i < 32; ++i) MBV_EL(i) o[i * step] = acc[i] + bias;
}

template <typename Op, int STEP, bool FULL, int RH>
__device__ __forceinline__ void epi_rs(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                       size_t rstep, const float* acc, const float* xpre) {
  using T = typename Op::T;
  const size_t step = STEP > 0 ? (size_t)STEP : rstep;
  const float bias = p.bias[(size_t)b * p.bias_bs + n];
  const bool res_half = (p.n_split > 0 && n < p.n_split);
  const int c = res_half ? n : n - p.n_split;
  const size_t base = ((size_t)b * p.rows_res + t_first) * p.ld + c;
  const float* mp = p.mask + (size_t)b * p.rows_res + t_first;
  float x[32];
  if (res_half) {  // x = (x + rs) * mask -> fp32 stream + operand copy for the next in_layer
#pragma unroll
    for (int i = 0; i < 32; ++i) { x[i] = 0.f; MBV_EL(i) x[i] = (xpre[i] + acc[i] + bias) * mp[i]; }
    T* dst = reinterpret_cast<T*>(p.act[0]) + base;
    if constexpr (RH == 2) {  // the operand copy is the stream
#pragma unroll
      for (int i = 0; i < 32; ++i) MBV_EL(i) op_store1<Op>(dst + i * step, x[i]);
    } else if constexpr (RH == 1) {  // fp16 stream next to the bf16 operand copy
      __half* xo = reinterpret_cast<__half*>(p.xout) + base;
#pragma unroll
      for (int i = 0; i < 32; ++i) MBV_EL(i) { xo[i * step] = to_half_sat(x[i]); op_store1<Op>(dst + i * step, x[i]); }
    } else {
      float* xo = reinterpret_cast<float*>(p.xout) + base;
#pragma unroll
      for (int i = 0; i < 32; ++i) MBV_EL(i) { xo[i * step] = x[i]; op_store1<Op>(dst + i * step, x[i]); }
    }
  } else {         // skip half: output += rs; the last layer applies the mask and emits the operand copy
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = acc[i] + bias + xpre[i];  // xpre = running skip sum (zeros for the first layer)
    if (p.n_split > 0) {
      float* so = reinterpret_cast<float*>(p.xs) + base;
#pragma unroll
      for (int i = 0; i < 32; ++i) MBV_EL(i) so[i * step] = x[i];
    } else {
      T* dst = reinterpret_cast<T*>(p.act[0]) + base;
#pragma unroll
      for (int i = 0; i < 32; ++i) MBV_EL(i) op_store1<Op>(dst + i * step, x[i] * mp[i]);
    }
  }
}

template <typename Op, int STEP, bool FULL>
__device__ __forceinline__ void epi_post(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                         size_t rstep, const float* acc, const float* xpre) {
  using T = typename Op::T;
  const size_t step = STEP > 0 ? (size_t)STEP : rstep;
  const float bias = p.bias[(size_t)b * p.bias_bs + n];
  const size_t base = ((size_t)b * p.rows_res + t_first) * p.ld + p.ch_off + n;
  const float* mp = p.mask + (size_t)b * p.rows_res + t_first;
  float* zo = reinterpret_cast<float*>(p.xout) + base;
  T* dst = reinterpret_cast<T*>(p.act[0]) + base;
  float z[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    z[i] = 0.f;
    MBV_EL(i) { const float m = mp[i]; z[i] = (xpre[i] - p.post_sign * ((acc[i] + bias) * m)) * m; }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) MBV_EL(i) { zo[i * step] = z[i]; op_store1<Op>(dst + i * step, z[i]); }
}

template <typename Op, int MODE, int STEP, bool FULL, int RH, bool FOLDED>
__device__ __forceinline__ void epi_dispatch(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                             size_t rstep, const float* acc, const float* xpre) {
  if constexpr (MODE == EPI_ACT) epi_act<Op, STEP, FULL, RH>(p, b, n, phase, t_first, nt, rstep, acc);
  else if constexpr (MODE == EPI_RES) epi_res<Op, STEP, FULL, RH, FOLDED>(p, b, n, phase, t_first, nt, rstep, acc, xpre);
  else if constexpr (MODE == EPI_F32) epi_f32<Op, STEP, FULL>(p, b, n, phase, t_first, nt, rstep, acc);
  else if constexpr (MODE == EPI_RS) epi_rs<Op, STEP, FULL, RH>(p, b, n, phase, t_first, nt, rstep, acc, xpre);
  else epi_post<Op, STEP, FULL>(p, b, n, phase, t_first, nt, rstep, acc, xpre);
}

// Residual-like input of a chunk (xin for EPI_RES / the residual half of EPI_RS, the running skip sum for the skip
// half, z for EPI_POST).  It does not depend on the accumulator, so the epilogue warps issue these loads one chunk
// AHEAD -- for the first chunk of a tile that is before the tile's MMAs have finished -- hiding the DRAM latency.
template <int MODE, int LD, int RH>
__device__ __forceinline__ void epi_prefetch(const EpiParams& p, int b, int n, int t_first, int nt, float* xpre) {
  const size_t step = LD > 0 ? (size_t)LD : (size_t)p.ld;
  if constexpr ((MODE == EPI_RES && RH != 0) || (MODE == EPI_RS && RH != 0)) {
    // (EPI_RS single stream: every row of the conv is a residual row, xin = the fp16 operand tensor h)
    const __half* hs = reinterpret_cast<const __half*>(p.xin) + ((size_t)b * p.rows_res + t_first) * p.ld + n;
    if (nt == 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) xpre[i] = __half2float(hs[i * step]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) xpre[i] = (i < nt) ? __half2float(hs[i * step]) : 0.f;
    }
    return;
  }
  const float* src = nullptr;
  if constexpr (MODE == EPI_RES) {
    src = reinterpret_cast<const float*>(p.xin) + ((size_t)b * p.rows_res + t_first) * p.ld + n;
  } else if constexpr (MODE == EPI_RS) {
    const bool res_half = (p.n_split > 0 && n < p.n_split);
    const int c = res_half ? n : n - p.n_split;
    if (res_half) src = reinterpret_cast<const float*>(p.xin) + ((size_t)b * p.rows_res + t_first) * p.ld + c;
    else if (!p.first) src = reinterpret_cast<const float*>(p.xs) + ((size_t)b * p.rows_res + t_first) * p.ld + c;
  } else if constexpr (MODE == EPI_POST) {
    src = reinterpret_cast<const float*>(p.xin) + ((size_t)b * p.rows_res + t_first) * p.ld + p.ch_off + n;
  }
  if (src != nullptr && nt == 32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) xpre[i] = src[i * step];
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) xpre[i] = (src != nullptr && i < nt) ? src[i * step] : 0.f;
  }
}

// LD: compile-time channel pitch of the destination buffers (0 = runtime).  The immediate-offset fast path also
// needs row_mul == 1 (everything but the polyphase upsamplers).
template <typename Op, int MODE, int LD, int RH, bool FOLDED = false>
__device__ __forceinline__ void tc_epilogue32(const EpiParams& p, int b, int n, int phase, int t_first, int nt,
                                              const float* acc, const float* xpre) {
  const size_t rstep = (size_t)p.row_mul * p.ld;
  if (LD > 0 && p.row_mul == 1) {
    if (nt == 32) epi_dispatch<Op, MODE, LD, true, RH, FOLDED>(p, b, n, phase, t_first, nt, rstep, acc, xpre);
    else epi_dispatch<Op, MODE, LD, false, RH, FOLDED>(p, b, n, phase, t_first, nt, rstep, acc, xpre);
  } else {
    if (nt == 32) epi_dispatch<Op, MODE, 0, true, RH, FOLDED>(p, b, n, phase, t_first, nt, rstep, acc, xpre);
    else epi_dispatch<Op, MODE, 0, false, RH, FOLDED>(p, b, n, phase, t_first, nt, rstep, acc, xpre);
  }
}

template <typename Op, int MODE, int LD, int RH, int CL>
__global__ void __launch_bounds__(TcThreads<MODE>::value, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmWh,
               const __grid_constant__ CUtensorMap tmS, const ConvArgs a, const TcRt rt) {
  using T = typename Op::T;
  constexpr int KB = TC_ROW_BYTES / (int)sizeof(T);  // channels per k-block
  constexpr int TC_EPI_WARPS = EpiWarps<MODE>::value;
  constexpr int TC_WARP_TMA = TC_EPI_WARPS, TC_WARP_MMA = TC_EPI_WARPS + 1;
  constexpr int CSTEP = 32 * (TC_EPI_WARPS / 4);  // column stride between the chunks of one epilogue warp
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smX = smem;                                                   // activation slabs
  uint8_t* smW = smX + (size_t)rt.n_slab_stages * rt.slab_stage_bytes;   // weight tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smW + (size_t)rt.n_w_stages * rt.w_stage_bytes);
  const int iXF = 0, iXE = iXF + rt.n_slab_stages, iWF = iXE + rt.n_slab_stages, iWE = iWF + rt.n_w_stages;
  constexpr bool kStagedRes = (MODE == EPI_RES || MODE == EPI_RS) && RH != 0;  // 2-byte residual input: may be staged by TMA
  const int iCF = iWE + rt.n_w_stages, iCE = iCF + 2, iRF = iCE + 2, nBars = iRF + (kStagedRes ? 2 * TC_EPI_WARPS : 0);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + nBars);
  float* xchg = reinterpret_cast<float*>(smem + rt.xchg_off);  // EPI_GATE: sigmoid -> tanh warp exchange, 4 pairs x 2 x 4 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < rt.n_slab_stages; ++i) { mbar_init(BAR(iXF + i), 1); mbar_init(BAR(iXE + i), 1); }
    // cluster mode: a weight stage may be refilled (by BOTH CTAs' multicasts) only when both CTAs have consumed it
    for (int i = 0; i < rt.n_w_stages; ++i) { mbar_init(BAR(iWF + i), 1); mbar_init(BAR(iWE + i), (CL == 1) ? 2 : 1); }
    // CTA pairs (CL == 2): the even CTA issues the MMAs of both, so ITS accumulator-free barrier collects the epilogue warps of both
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(iCF + i), 1); mbar_init(BAR(iCE + i), (CL == 2) ? 2 * TC_EPI_WARPS : TC_EPI_WARPS); }
    if constexpr (kStagedRes) {
      for (int i = 0; i < 2 * TC_EPI_WARPS; ++i) mbar_init(BAR(iRF + i), 1);
      tma_prefetch_desc(&tmR);
      tma_prefetch_desc(&tmS);
    }
    fence_barrier_init();
  }
  if (warp == TC_WARP_MMA) {
    if constexpr (CL == 2) tmem_alloc2(smem_u32(tmem_ptr_smem), 512u);  // the same 512 columns in both CTAs of the pair
    else tmem_alloc(smem_u32(tmem_ptr_smem), 512u);
  }
  tc_fence_before();
  __syncthreads();
  if ((CL != 0)) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the
  // tail of the previous kernel in the stream; nothing below may touch global memory before that kernel has
  // completed and flushed.  Our own dependents may be scheduled as soon as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int kblocks = a.Cp_in / KB;
  const uint32_t slab_bytes = (uint32_t)rt.n_boxes * rt.box_rows * TC_ROW_BYTES;
  const uint32_t w_tile_bytes = TC_M * TC_ROW_BYTES;
  const uint32_t crank = blockIdx.x & 1u;  // cluster rank (clusters are pairs of consecutive blocks)
  // tile index -> (channel tile, phase, time tile, utterance, does this CTA have a time tile).  Cluster mode walks pairs:
  // tiles 2w and 2w+1 (always on the two CTAs of one cluster, the grid is even) share the weight group w % groups and take
  // the time tiles 2j and 2j+1, j = w / groups; the odd one out at the end only relays weight tiles.
  struct Work { int ct, phase, tt, b, t_off, n_cols; bool row_ok; };
  auto decode_work = [&](int tile) {
    Work wk;
    wk.t_off = 0; wk.n_cols = rt.n_time;
    if constexpr (CL == 2) {
      // CTA pairs: the unit of the split is the pair (both CTAs work on the same columns of one time tile); split_from counts pairs
      int w = tile >> 1;
      if (w >= rt.split_from && rt.split_k > 1) {
        const int s = w - rt.split_from;
        w = rt.split_from + s / rt.split_k;
        wk.t_off = (s % rt.split_k) * rt.split_n;
        wk.n_cols = min(rt.split_n, rt.n_time - wk.t_off);
        tile = 2 * w + (tile & 1);
      }
    } else if (tile >= rt.split_from && rt.split_k > 1) {  // a piece of a tile of the last round
      const int s = tile - rt.split_from;
      tile = rt.split_from + s / rt.split_k;
      wk.t_off = (s % rt.split_k) * rt.split_n;
      wk.n_cols = min(rt.split_n, rt.n_time - wk.t_off);
    }
    if constexpr (CL == 2) {
      // CTA pairs: tiles 2w and 2w+1 (the two CTAs of a cluster) are the two channel tiles 2p and 2p+1 of ONE time tile
      // (or, for an upsampler with one channel tile, two polyphase branches that read the same input rows)
      int rest = tile >> 1;
      if (rt.pair_phase) {
        const int ppairs = a.n_phases >> 1;
        wk.ct = rest % rt.c_tiles; rest /= rt.c_tiles;
        wk.phase = 2 * (rest % ppairs) + (tile & 1); rest /= ppairs;
      } else {
        const int cpairs = rt.c_tiles >> 1;
        wk.ct = 2 * (rest % cpairs) + (tile & 1); rest /= cpairs;
        wk.phase = rest % a.n_phases; rest /= a.n_phases;
      }
      wk.tt = rest % rt.t_tiles; wk.b = rest / rt.t_tiles;
      wk.row_ok = true;
    } else if constexpr (CL == 1) {
      const int w = tile >> 1, grp = w % rt.groups, row = 2 * (w / rt.groups) + (tile & 1);
      wk.ct = grp % rt.c_tiles; wk.phase = grp / rt.c_tiles;
      wk.row_ok = row < rt.rows;
      wk.tt = row % rt.t_tiles; wk.b = row / rt.t_tiles;
    } else {
      int rest = tile;
      wk.ct = rest % rt.c_tiles; rest /= rt.c_tiles;
      if (rt.rotate) wk.ct = (wk.ct + tile / (int)gridDim.x) & 1;
      wk.phase = rest % a.n_phases; rest /= a.n_phases;
      wk.tt = rest % rt.t_tiles; wk.b = rest / rt.t_tiles;
      wk.row_ok = true;
    }
    return wk;
  };

  if (warp == TC_WARP_TMA) {
    // ===================== TMA producer =====================
    // The whole warp runs the (warp-uniform) loop so addresses and coordinates live in uniform registers; one
    // elected lane issues the copies.
    int sx = 0, sw = 0;
    uint32_t px = 0, pw = 0;
    for (int tile = blockIdx.x; tile < rt.virt_tiles; tile += gridDim.x) {
      const Work wk = decode_work(tile);
      const int ct = wk.ct, phase = wk.phase, b = wk.b;
      const int t0 = wk.tt * rt.n_time + wk.t_off;
      const int wrow0 = phase * a.taps * a.N_total + ct * TC_M;
      // CTA pairs: this CTA stages its HALF of the tile's time rows (the B operand of a cta_group::2 MMA is split by rows)
      const int xrow0 = t0 + a.shift0[phase] + ((CL == 2) ? (int)crank * (wk.n_cols >> 1) : 0);
      if (rt.dbg && lane == 0) rt.dbg[((size_t)blockIdx.x * 7 + 0) * 64 + 2 * ((tile / gridDim.x) & 31)] = clock64();
      // (A whole-tile TMA L2 prefetch of the residual box issued here was measured to be too early: a tile-time later
      //  a third of it had been evicted again and DRAM reads grew 40 %.  The epilogue warps prefetch two chunks ahead.)
      for (int kb = 0; kb < kblocks; ++kb) {
        if (wk.row_ok) {
          mbar_wait(BAR(iXE + sx), px ^ 1);
          if (elect_one()) {
            const uint32_t dst = smem_u32(smX + (size_t)sx * rt.slab_stage_bytes);
            if constexpr (CL == 2) {
              // both CTAs' copies complete on the EVEN CTA's barrier, which expects the bytes of both
              if (crank == 0) mbar_expect_tx(BAR(iXF + sx), 2 * slab_bytes);
              const uint32_t xf = mapa_shared(BAR(iXF + sx), 0);
              for (int i = 0; i < rt.n_boxes; ++i)
                tma_load_3d_2sm(dst + (uint32_t)(i * rt.box_rows) * TC_ROW_BYTES, &tmX, xf, kb * KB, xrow0 + i * rt.box_rows, b);
            } else {
              mbar_expect_tx(BAR(iXF + sx), slab_bytes);
              for (int i = 0; i < rt.n_boxes; ++i)
                tma_load_3d(dst + (uint32_t)(i * rt.box_rows) * TC_ROW_BYTES, &tmX, BAR(iXF + sx), kb * KB,
                            xrow0 + i * rt.box_rows, b);
            }
          }
          __syncwarp();
          if (++sx == rt.n_slab_stages) { sx = 0; px ^= 1; }
        }
        for (int tap = 0; tap < a.taps; ++tap) {
          if (rt.w_resident && tile != (int)blockIdx.x) continue;  // weights already resident from the first tile
          mbar_wait(BAR(iWE + sw), pw ^ 1);
          if (elect_one()) {
            const uint32_t wdst = smem_u32(smW + (size_t)sw * rt.w_stage_bytes);
            if constexpr (CL == 2) {  // this CTA's own 128 weight rows (its channel tile = its half of the M = 256 A operand)
              if (crank == 0) mbar_expect_tx(BAR(iWF + sw), 2 * w_tile_bytes);
              tma_load_2d_2sm(wdst, &tmW, mapa_shared(BAR(iWF + sw), 0), kb * KB, wrow0 + tap * a.N_total);
            } else if constexpr (CL == 1) {  // this CTA's 64 rows of the tile, delivered to both CTAs of the pair
              mbar_expect_tx(BAR(iWF + sw), w_tile_bytes);
              tma_load_2d_mc(wdst + crank * (w_tile_bytes / 2), &tmWh, BAR(iWF + sw), kb * KB,
                             wrow0 + tap * a.N_total + (int)crank * (TC_M / 2), (uint16_t)3);
            } else {
              mbar_expect_tx(BAR(iWF + sw), w_tile_bytes);
              tma_load_2d(wdst, &tmW, BAR(iWF + sw), kb * KB, wrow0 + tap * a.N_total);
            }
          }
          __syncwarp();
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop, one elected lane issues.  Per (tap, k-block): 4 MMAs whose descriptors differ only by
    // +32 bytes in the low word -- the issue path must stay far below the 128 tensor-core cycles one MMA takes.
    // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16 or tf32, both K-major, N, M=128
    constexpr uint32_t fmt = (Op::kPrec == 3) ? 0u : ((Op::kPrec == 2) ? 1u : 2u);  // F16 / BF16 / TF32
    constexpr int KIND = (Op::kPrec >= 2) ? 2 : 1;
    const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(((CL == 2) ? 2 * TC_M : TC_M) >> 4) << 24);
    const uint32_t tap_step = (uint32_t)(a.dil * TC_ROW_BYTES) >> 4;  // descriptor-lo increment per tap
    int sx = 0, sw = 0, sc = 0;
    uint32_t px = 0, pw = 0, pc = 0;
    for (int tile = blockIdx.x; tile < rt.virt_tiles; tile += gridDim.x) {
      if ((CL == 2) && crank != 0) break;  // CTA pairs: the even CTA issues the MMAs of both
      if ((CL == 1) && !decode_work(tile).row_ok) {
        // no time tile for this CTA (odd tile count): keep the pair's weight ring moving -- wait until each stage has
        // fully landed here, then release it on both CTAs
        for (int i = 0; i < kblocks * a.taps; ++i) {
          mbar_wait(BAR(iWF + sw), pw);
          if (elect_one()) { mbar_arrive(BAR(iWE + sw)); mbar_arrive_remote(BAR(iWE + sw), crank ^ 1u); }
          __syncwarp();
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
        continue;
      }
      const uint32_t idesc = idesc0 | ((uint32_t)(decode_work(tile).n_cols >> 3) << 17);  // UMMA N = this tile's columns
      long long tw0 = rt.dbg ? clock64() : 0, wait_c = 0, wait_x = 0, wait_w = 0;
      mbar_wait(BAR(iCE + sc), pc ^ 1);
      if (rt.dbg) wait_c = clock64() - tw0;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(sc * TC_ACC_STRIDE);
      if (rt.dbg && lane == 0) rt.dbg[((size_t)blockIdx.x * 7 + 1) * 64 + 2 * ((tile / gridDim.x) & 31)] = clock64();
      uint32_t accum = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        if (rt.dbg) tw0 = clock64();
        mbar_wait(BAR(iXF + sx), px);
        if (rt.dbg) wait_x += clock64() - tw0;
        tc_fence_after();
        uint32_t x_lo = desc_lo(smem_u32(smX + (size_t)sx * rt.slab_stage_bytes));
        for (int tap = 0; tap < a.taps; ++tap) {
          if (rt.w_resident) { sw = kb * a.taps + tap; pw = 0; }  // stage = (k-block, tap); phase 0 stays complete
          if (rt.dbg) tw0 = clock64();
          mbar_wait(BAR(iWF + sw), pw);
          if (rt.dbg) wait_w += clock64() - tw0;
          tc_fence_after();
          const uint32_t w_lo = desc_lo(smem_u32(smW + (size_t)sw * rt.w_stage_bytes));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if constexpr (CL == 2) tc_mma2<KIND>(tmem_d, desc64(w_lo + 2 * k), desc64(x_lo + 2 * k), idesc, (k == 0) ? accum : 1u);
              else tc_mma<KIND>(tmem_d, desc64(w_lo + 2 * k), desc64(x_lo + 2 * k), idesc, (k == 0) ? accum : 1u);
            }
            if constexpr (CL == 2) tc_commit2_mc(BAR(iWE + sw), (uint16_t)3);
            else if constexpr (CL == 1) tc_commit_mc(BAR(iWE + sw), (uint16_t)3);
            else if (!rt.w_resident) tc_commit(BAR(iWE + sw));
          }
          __syncwarp();
          accum = 1;
          x_lo += tap_step;  // next tap = `dil` rows further into the slab
          if (!rt.w_resident && ++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
        if (elect_one()) { if constexpr (CL == 2) tc_commit2_mc(BAR(iXE + sx), (uint16_t)3); else tc_commit(BAR(iXE + sx)); }
        __syncwarp();
        if (++sx == rt.n_slab_stages) { sx = 0; px ^= 1; }
      }
      if (elect_one()) { if constexpr (CL == 2) tc_commit2_mc(BAR(iCF + sc), (uint16_t)3); else tc_commit(BAR(iCF + sc)); }
      __syncwarp();
      if (rt.dbg && lane == 0) {
        rt.dbg[((size_t)blockIdx.x * 7 + 1) * 64 + 2 * ((tile / gridDim.x) & 31) + 1] = clock64();
        if (tile / (int)gridDim.x < 16) {  // issuer-side waits of this tile: accumulator free | slabs | weight tiles
          long long* w = rt.dbg + ((size_t)blockIdx.x * 7 + 6) * 64 + 3 * (tile / gridDim.x);
          w[0] = wait_c; w[1] = wait_x; w[2] = wait_w;
        }
      }
      if (++sc == 2) { sc = 0; pc ^= 1; }
    }
  } else if (warp < TC_EPI_WARPS) {
    // ===================== epilogue warps =====================
    // Work items of this warp: (tile, 32-column chunk c = grp*32 + CSTEP*j); tiles are decoded once per tile.
    constexpr bool kPrefetch = (MODE == EPI_RES || MODE == EPI_RS || MODE == EPI_POST);
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = warp >> 2;             // column group: which interleaved 32-column chunks this warp takes
    const int n_valid = a.epi.n_valid;
    const int c_first = half * 32;
    int sc = 0;
    uint32_t pc = 0;

    struct TileInfo { int b, n, nb, phase, t0, t_lim, n_cols; bool valid, row_ok, wv; };
    auto decode = [&](int tile) {
      TileInfo ti;
      const Work wk = decode_work(tile);
      const int ct = wk.ct;
      ti.phase = wk.phase;
      ti.b = wk.b;
      ti.row_ok = wk.row_ok;
      ti.t0 = wk.tt * rt.n_time + wk.t_off;
      ti.n_cols = wk.n_cols;
      ti.n = ct * TC_M + q * 32 + lane;  // weight row = output channel of this thread
      ti.nb = ct * TC_M + q * 32;        // first channel of this warp: its 32 channels are 64 contiguous bytes of a residual row
      ti.wv = ti.nb < a.epi.ld && ti.nb < n_valid;
      // which logical channel does this row write, and is it inside the destination buffer?
      if (MODE == EPI_RS && a.epi.n_split > 0) ti.valid = (ti.n < a.epi.n_split ? ti.n : ti.n - a.epi.n_split) < n_valid;
      else if (MODE == EPI_GATE) ti.valid = (64 * ct + ((q * 32 + lane) & 63)) < n_valid;  // logical channel of this row
      else ti.valid = ti.n < n_valid;
      ti.t_lim = min(a.L_out, ti.t0 + wk.n_cols);
      return ti;
    };
    // The residual loads are latency-bound (per-tile timelines, MBV_TIMELINE).  L2 prefetches cost no registers, so at
    // the start of a tile every lane asks L2 for its rows of the WHOLE NEXT tile (its warp's 32 channels = 64-128
    // contiguous bytes per row); one tile period later the demand loads find them in L2.
    auto l2_prefetch = [&](const TileInfo& ti, int c) {
      if constexpr (MODE == EPI_RES) {
        const int t = ti.t0 + c + lane;
        if (ti.valid && c < ti.n_cols && t < ti.t_lim) {
          const size_t off = ((size_t)ti.b * a.epi.rows_res + t) * a.epi.ld + (ti.n - lane);
          const char* p = reinterpret_cast<const char*>(a.epi.xin) + off * (RH != 0 ? 2 : 4);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
      }
    };
    const int nch = (rt.n_time - c_first + CSTEP - 1) / CSTEP;  // chunks per tile for this warp
    auto prefetch = [&](const TileInfo& ti, int c, float* dst) {
      const int t_first = ti.t0 + c;
      if (ti.valid && t_first < ti.t_lim)
        epi_prefetch<MODE, LD, RH>(a.epi, ti.b, ti.n, t_first, min(ti.t_lim - t_first, 32), dst);
    };

    // (Measured and dropped: keeping the fp16 residual packed two-per-register and loading it one chunk AHEAD -- same
    //  register count on paper -- spilled ~100 bytes and made every ResBlock conv 5-10 % slower, 8.9 -> 9.3 ms per step.)
    // Staged residual (kStagedRes): every epilogue warp owns one 2 KB shared-memory buffer + mbarrier per
    // staged stream (the residual input xin and, for the ResBlock-sum epilogues, the running sum xs).  A chunk's values are
    // copied to registers at its very start, and the ONE TMA box per stream (32 channels x 32 rows of the 2-byte tensor,
    // rows past the utterance zero-filled) for the warp's NEXT chunk -- of this tile or the first one of its next tile --
    // is requested right away, so it is in flight while this chunk is processed: the loads cost no registers, no LSU
    // queue slots behind the warp's own stores, and their DRAM latency is off the chunk's critical path (per-chunk
    // stamps of the register path: 1-2 K of a chunk's 3 K cycles went into issuing and awaiting 32 two-byte loads,
    // profiles/r02_res_epilogue_timeline.txt; the synchronous xs loads made the summing c2 convs 25 % slower still).
    constexpr bool staged = kStagedRes;  // (the planner refuses these epilogues without the staging buffers)
    const bool staged_xs = staged && MODE == EPI_RES && (a.epi.sum_mode == 2 || a.epi.sum_mode == 3);
    const uint32_t res_buf = smem_u32(smem + rt.res_off) + (uint32_t)warp * 2048u;
    const uint32_t xs_buf = res_buf + (uint32_t)TC_EPI_WARPS * 2048u;
    const uint32_t bar_r = BAR(iRF + 2 * warp), bar_s = BAR(iRF + 2 * warp + 1);
    auto has_load = [&](const TileInfo& t, int c) { return t.wv && c < t.n_cols && t.t0 + c < t.t_lim; };
    auto res_issue = [&](const TileInfo& t, int c) {
      if (elect_one()) {
        mbar_expect_tx(bar_r, 2048u);
        tma_load_3d(res_buf, &tmR, bar_r, t.nb, t.t0 + c, t.b);
        if (staged_xs) {
          mbar_expect_tx(bar_s, 2048u);
          tma_load_3d(xs_buf, &tmS, bar_s, t.nb, t.t0 + c, t.b);
        }
      }
      __syncwarp();
    };
    uint32_t rph = 0;       // phase of the warp's staging barriers
    bool rpending = false;  // the boxes of the next chunk this warp consumes have been requested

    float xcur[32];
    int gate_chunk = 0;
    int tile = blockIdx.x;
    TileInfo ti = decode(tile);
    if (kPrefetch && !staged && tile < rt.virt_tiles)
      for (int j = 0; j < nch; ++j) l2_prefetch(ti, c_first + CSTEP * j);

    while (tile < rt.virt_tiles) {
      const bool have_next = tile + (int)gridDim.x < rt.virt_tiles;
      TileInfo tn = ti;
      if (have_next) {
        tn = decode(tile + gridDim.x);
        if (tn.row_ok && !staged) for (int j = 0; j < nch; ++j) l2_prefetch(tn, c_first + CSTEP * j);
      }
      if (!ti.row_ok) { ti = tn; tile += gridDim.x; continue; }  // cluster mode: this CTA only relayed weights for this tile
      if constexpr (kStagedRes) {
        if (!rpending && has_load(ti, c_first)) { res_issue(ti, c_first); rpending = true; }  // cold start
      }
      mbar_wait(BAR(iCF + sc), pc);
      tc_fence_after();
      if (rt.dbg && warp == 0 && lane == 0) rt.dbg[((size_t)blockIdx.x * 7 + 2) * 64 + 2 * ((tile / gridDim.x) & 31)] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sc * TC_ACC_STRIDE);
      float gate_bias = 0.f;
      if constexpr (MODE == EPI_GATE) {  // bias + cond_layer(g) of this thread's weight row, once per tile
        gate_bias = a.epi.bias[(size_t)ti.b * a.epi.bias_bs + ti.n];
        if (a.epi.add2) gate_bias += a.epi.add2[(size_t)ti.b * a.epi.add2_bs + ti.n];
      }
      for (int c = c_first; c < ti.n_cols; c += CSTEP) {
        float acc[32];
        // debug stamps of warp 0's chunks (tiles 0..9): issue | accumulator in registers | residual in registers | done
        long long* cdbg = nullptr;
        if (rt.dbg && warp == 0 && lane == 0 && tile / (int)gridDim.x < 10)
          cdbg = rt.dbg + ((size_t)blockIdx.x * 7 + 3) * 64 + ((tile / gridDim.x) * 3 + (c - c_first) / CSTEP) * 4;
        if (cdbg) cdbg[0] = clock64();
        tmem_ld32(taddr + (uint32_t)c, acc);
        if constexpr (kStagedRes) {
          if (has_load(ti, c)) {
            if (!rpending) res_issue(ti, c);
            mbar_wait(bar_r, rph);
            const uint32_t src = res_buf + (uint32_t)lane * 2u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              unsigned short hv;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(src + (uint32_t)i * 64u));
              xcur[i] = __half2float(__ushort_as_half(hv));
            }
            if constexpr (MODE == EPI_RES && RH == 2) {  // single stream: xin holds lrelu(x); inverse leaky-relu (inv_slope >= 1)
              const float inv = a.epi.inv_slope;
#pragma unroll
              for (int i = 0; i < 32; ++i) xcur[i] = fminf(xcur[i], xcur[i] * inv);
            }
            if constexpr (MODE == EPI_RES) {
              if (staged_xs) {  // running ResBlock sum: folded into the residual here, epi_res then skips its own xs loads
                mbar_wait(bar_s, rph);
                const uint32_t ssrc = xs_buf + (uint32_t)lane * 2u;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  unsigned short hv;
                  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(ssrc + (uint32_t)i * 64u));
                  xcur[i] += __half2float(__ushort_as_half(hv));
                }
              }
            }
            rph ^= 1u;
            __syncwarp();  // every lane has copied its values out before the next boxes may land in the buffers
            // request the next chunk of this warp: the next one of this tile, else the first one of its next tile
            const int nc = c + CSTEP;
            rpending = false;
            if (nc < ti.n_cols) { if (has_load(ti, nc)) { res_issue(ti, nc); rpending = true; } }
            else if (have_next && tn.row_ok && has_load(tn, c_first)) { res_issue(tn, c_first); rpending = true; }
          }
        }
        if constexpr (kPrefetch && !kStagedRes) prefetch(ti, c, xcur);  // residual of THIS chunk, in flight with the TMEM load
        tmem_ld_wait();
        if (cdbg) {
          cdbg[1] = clock64();
          if constexpr (kPrefetch) {
            float sdep = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sdep += xcur[i];
            if (sdep == 1.2345e-30f) cdbg[1] = 0;  // forces the wait for every residual load
            cdbg[2] = clock64();
          }
        }
        const int t_first = ti.t0 + c;
        if constexpr (MODE == EPI_GATE) {
          // One accumulator tile = [64 tanh rows | 64 sigmoid rows] of the same 64 channels: lanes i and i+64 belong
          // together but live in different warps (q and q+2).  The sigmoid warp hands its 32x32 block to its tanh
          // partner through shared memory (two buffers, one named barrier per chunk), which gates and stores.
          const EpiParams& p = a.epi;
          const bool is_sig = q >= 2;
          const int pair = (q & 1) + 2 * half;
          const uint32_t xb = smem_u32(xchg) + (uint32_t)((pair * 2 + (gate_chunk & 1)) * 4096) + (uint32_t)lane * 4u;
          gate_chunk++;
          if (is_sig) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float sg = gate_sigmoid<Op>(acc[i] + gate_bias);
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(xb + (uint32_t)i * 128u), "f"(sg) : "memory");
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = gate_tanh<Op>(acc[i] + gate_bias);  // overlaps the partner's sigmoids
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
          if (!is_sig && ti.valid && t_first < ti.t_lim) {
            using T = typename Op::T;
            const int nt = min(ti.t_lim - t_first, 32);
            const int ch = 64 * (ti.n >> 7) + (ti.n & 63);
            T* dst = reinterpret_cast<T*>(p.act[0]) + ((size_t)ti.b * p.rows_out + t_first) * p.ld + p.ch_off + ch;
            const size_t step = LD > 0 ? (size_t)LD : (size_t)p.ld;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float sg;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sg) : "r"(xb + (uint32_t)i * 128u) : "memory");
              if (i < nt) op_store1<Op>(dst + i * step, acc[i] * sg);
            }
          }
        } else if (MODE == EPI_ACT && sizeof(T) == 2 && rt.pack2 && t_first < ti.t_lim && ti.nb + 32 <= n_valid) {
          // (warp-uniform condition: all 32 channels of this warp are stored -> lane pairs may trade halves)
          if constexpr (MODE == EPI_ACT && sizeof(T) == 2)
            tc_epilogue_act_packed<Op, LD, RH>(a.epi, ti.b, ti.n, ti.phase, t_first, min(ti.t_lim - t_first, 32), acc);
        } else if (ti.valid && t_first < ti.t_lim) {
          tc_epilogue32<Op, MODE, LD, RH, kStagedRes>(a.epi, ti.b, ti.n, ti.phase, t_first, min(ti.t_lim - t_first, 32), acc, xcur);
        }
        if (cdbg) cdbg[3] = clock64();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if ((CL == 2) && crank != 0) mbar_arrive_remote(BAR(iCE + sc), 0u);  // the accumulator-free barrier lives in the even CTA
        else mbar_arrive(BAR(iCE + sc));
      }
      if (rt.dbg && warp == 0 && lane == 0) rt.dbg[((size_t)blockIdx.x * 7 + 2) * 64 + 2 * ((tile / gridDim.x) & 31) + 1] = clock64();
      if (++sc == 2) { sc = 0; pc ^= 1; }
      ti = tn;
      tile += gridDim.x;
    }
  }
  tc_fence_before();
  __syncthreads();
  if ((CL != 0)) cluster_sync_all();  // the peer may still be multicasting into this CTA's ring / arriving on its barriers
  if (warp == TC_WARP_MMA) {
    tc_fence_after();
    if constexpr (CL == 2) tmem_dealloc2(tmem_base, 512u);
    else tmem_dealloc(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused ResBlock1 conv pair (modules.py:217-224):  x' = x + c2(lrelu(c1(a) + b1)) + b2,  a = lrelu(x) the operand tensor.
// At 128 channels a k=3 conv is HBM-bound (96 FLOP/B) and even the k=7 / k=11 pairs spend a third of their time moving
// the intermediate h = lrelu(c1(a)) out to HBM and back (profiles/r02_launch_times_bf16.txt: stage-1 pairs take
// 302 / 379 / 496 us against 106 / 246 / 388 us of tensor work).  This kernel keeps h on the SM:
//   * one CTA = all 128 channels x 160 output columns.  c1 accumulates D1[128, 176] for the columns [t0-8, t0+168)
//     (the halo c2 needs, K <= 17); the epilogue warps turn it into bf16 lrelu(.) rows -- zero outside the utterance,
//     which is c2's own zero padding -- and store them K-major, 128B-swizzled, into a shared-memory tile that is the
//     B operand of c2 (tap j = rows [8 - (K-1)/2 + j, +160)).  c2 accumulates D2[128, 160] into one of TWO further TMEM
//     buffers and the normal RES epilogue drains it: conv 1 and conv 2 of the next tile run while the residual add of
//     this tile is still streaming to HBM (160 columns is what 176 + 2 x 160 <= 512 TMEM columns allow).
//   * the tensor core never waits for HBM for the second conv, the intermediate never exists in memory, and the pair
//     reads a and the residual once and writes its two outputs once.
//   * warp roles: 8 warps residual add (epilogue 2), 4 warps h tile (epilogue 1), 1 TMA producer, 1 MMA issuer.
// Geometry handled: Cp_in = C_out = 128 (one channel tile, two k-blocks), stride 1, c2 dilation 1; everything else
// takes the two-launch path.
// ------------------------------------------------------------------------------------------------
constexpr int PAIR_N = 160;      // output columns per tile
constexpr int PAIR_HALO = 8;     // columns of D1 before / after the outputs
constexpr int PAIR_N1 = PAIR_N + 2 * PAIR_HALO;  // columns of D1 (176)
constexpr int PAIR_D2_COL0 = 192, PAIR_D2_STRIDE = 160;  // TMEM: D1 [0,176), D2 buffers [192,352) and [352,512)
constexpr int PAIR_H_BYTES = 2 * PAIR_N1 * TC_ROW_BYTES;

struct PairRt {
  int pdl;
  int box_rows, n_boxes, slab_stage_bytes, n_slab_stages, n_w_stages;
  int t_tiles, total_tiles;
  int h_off, w_off, bar_off;  // byte offsets in (aligned) dynamic smem
  long long* dbg;             // MBV_TIMELINE=9: CTA 0 clock stamps [tile][8] (debug only)
};

template <typename Op, int LD, int RH>
__global__ void __launch_bounds__(TcThreads<EPI_RES>::value, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const ConvArgs a, const PairRt rt) {
  using T = typename Op::T;
  constexpr int KB = 64;
  constexpr int NEPI = EpiWarps<EPI_RES>::value;  // 12 = 8 residual-add warps (epilogue 2) + 4 h-tile warps (epilogue 1)
  constexpr int NEPI2 = 8, NEPI1 = NEPI - NEPI2;
  constexpr int WARP_TMA = NEPI, WARP_MMA = NEPI + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smX = smem;
  uint8_t* smH = smem + rt.h_off;
  uint8_t* smW = smem + rt.w_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + rt.bar_off);
  const int iXF = 0, iXE = iXF + rt.n_slab_stages, iWF = iXE + rt.n_slab_stages, iWE = iWF + rt.n_w_stages;
  const int iD1F = iWE + rt.n_w_stages, iHR = iD1F + 1, iD2F = iHR + 1, iD2E = iD2F + 2, nBars = iD2E + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + nBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int i = 0; i < rt.n_slab_stages; ++i) { mbar_init(BAR(iXF + i), 1); mbar_init(BAR(iXE + i), 1); }
    for (int i = 0; i < rt.n_w_stages; ++i) { mbar_init(BAR(iWF + i), 1); mbar_init(BAR(iWE + i), 1); }
    mbar_init(BAR(iD1F), 1);
    mbar_init(BAR(iHR), NEPI1);
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(iD2F + i), 1); mbar_init(BAR(iD2E + i), NEPI2); }
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_ptr_smem), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int taps = a.taps;
  const uint32_t slab_bytes = (uint32_t)rt.n_boxes * rt.box_rows * TC_ROW_BYTES;
  const uint32_t w_tile_bytes = TC_M * TC_ROW_BYTES;

  if (warp == WARP_TMA) {
    // ===================== TMA producer: per tile  [slab kb, W1 taps of kb] x 2, then W2 (kb, tap) =====================
    int sx = 0, sw = 0;
    uint32_t px = 0, pw = 0;
    for (int tile = blockIdx.x; tile < rt.total_tiles; tile += gridDim.x) {
      const int tt = tile % rt.t_tiles, b = tile / rt.t_tiles;
      const int xrow0 = tt * PAIR_N - PAIR_HALO + a.shift0[0];
      for (int kb = 0; kb < 2; ++kb) {
        mbar_wait(BAR(iXE + sx), px ^ 1);
        if (elect_one()) {
          mbar_expect_tx(BAR(iXF + sx), slab_bytes);
          const uint32_t dst = smem_u32(smX + (size_t)sx * rt.slab_stage_bytes);
          for (int i = 0; i < rt.n_boxes; ++i)
            tma_load_3d(dst + (uint32_t)(i * rt.box_rows) * TC_ROW_BYTES, &tmX, BAR(iXF + sx), kb * KB,
                        xrow0 + i * rt.box_rows, b);
        }
        __syncwarp();
        if (++sx == rt.n_slab_stages) { sx = 0; px ^= 1; }
        for (int tap = 0; tap < taps; ++tap) {
          mbar_wait(BAR(iWE + sw), pw ^ 1);
          if (elect_one()) {
            mbar_expect_tx(BAR(iWF + sw), w_tile_bytes);
            tma_load_2d(smem_u32(smW + (size_t)sw * TC_M * TC_ROW_BYTES), &tmW1, BAR(iWF + sw), kb * KB, tap * TC_M);
          }
          __syncwarp();
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
      }
      for (int kb = 0; kb < 2; ++kb)
        for (int tap = 0; tap < taps; ++tap) {
          mbar_wait(BAR(iWE + sw), pw ^ 1);
          if (elect_one()) {
            mbar_expect_tx(BAR(iWF + sw), w_tile_bytes);
            tma_load_2d(smem_u32(smW + (size_t)sw * TC_M * TC_ROW_BYTES), &tmW2, BAR(iWF + sw), kb * KB, tap * TC_M);
          }
          __syncwarp();
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer =====================
    constexpr uint32_t fmt = (Op::kPrec == 3) ? 0u : 1u;
    const uint32_t idesc1 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(PAIR_N1 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(PAIR_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t tap_step1 = (uint32_t)(a.dil * TC_ROW_BYTES) >> 4;
    const uint32_t h_row0 = (uint32_t)(PAIR_HALO - (taps - 1) / 2);
    int sx = 0, sw = 0, d2 = 0;
    uint32_t px = 0, pw = 0, ph = 0, pd2 = 0;
    for (int tile = blockIdx.x; tile < rt.total_tiles; tile += gridDim.x) {
      // ---- conv 1 -> D1 (TMEM columns [0, 176)); D1 is free: the previous tile's epilogue 1 signalled HR before conv 2
      uint32_t accum = 0;
      long long* dbg = (rt.dbg && blockIdx.x == 0 && lane == 0 && tile / (int)gridDim.x < 16) ? rt.dbg + (tile / gridDim.x) * 8 : nullptr;
      if (dbg) dbg[0] = clock64();
      for (int kb = 0; kb < 2; ++kb) {
        mbar_wait(BAR(iXF + sx), px);
        tc_fence_after();
        uint32_t x_lo = desc_lo(smem_u32(smX + (size_t)sx * rt.slab_stage_bytes));
        for (int tap = 0; tap < taps; ++tap) {
          mbar_wait(BAR(iWF + sw), pw);
          tc_fence_after();
          const uint32_t w_lo = desc_lo(smem_u32(smW + (size_t)sw * TC_M * TC_ROW_BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma<2>(tmem_base, desc64(w_lo + 2 * k), desc64(x_lo + 2 * k), idesc1, (k == 0) ? accum : 1u);
            tc_commit(BAR(iWE + sw));
          }
          __syncwarp();
          accum = 1;
          x_lo += tap_step1;
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
        if (elect_one()) tc_commit(BAR(iXE + sx));
        __syncwarp();
        if (++sx == rt.n_slab_stages) { sx = 0; px ^= 1; }
      }
      if (elect_one()) tc_commit(BAR(iD1F));
      __syncwarp();
      if (dbg) dbg[1] = clock64();
      // ---- conv 2 -> D2 buffer d2: needs the h tile (epilogue 1 of this tile) and that buffer drained (epilogue 2 of tile i-2)
      mbar_wait(BAR(iHR), ph);
      mbar_wait(BAR(iD2E + d2), pd2 ^ 1);
      tc_fence_after();
      if (dbg) dbg[2] = clock64();
      accum = 0;
      for (int kb = 0; kb < 2; ++kb) {
        const uint32_t h_lo = desc_lo(smem_u32(smH + (size_t)kb * PAIR_N1 * TC_ROW_BYTES) + h_row0 * TC_ROW_BYTES);
        for (int tap = 0; tap < taps; ++tap) {
          mbar_wait(BAR(iWF + sw), pw);
          tc_fence_after();
          const uint32_t w_lo = desc_lo(smem_u32(smW + (size_t)sw * TC_M * TC_ROW_BYTES));
          const uint32_t b_lo = h_lo + (uint32_t)tap * (TC_ROW_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma<2>(tmem_base + PAIR_D2_COL0 + d2 * PAIR_D2_STRIDE, desc64(w_lo + 2 * k), desc64(b_lo + 2 * k), idesc2, (k == 0) ? accum : 1u);
            tc_commit(BAR(iWE + sw));
          }
          __syncwarp();
          accum = 1;
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
      }
      if (elect_one()) tc_commit(BAR(iD2F + d2));
      __syncwarp();
      if (dbg) dbg[3] = clock64();
      ph ^= 1;
      if (++d2 == 2) { d2 = 0; pd2 ^= 1; }
    }
  } else if (warp >= NEPI2 && warp < NEPI) {
    // ===================== epilogue-1 warps (4): D1 -> h tile =====================
    // h = lrelu(D1 + b1) in the operand type, zero outside the utterance, into the c2 operand tile.  These warps run
    // ahead of epilogue 2: tile i+1's h is written while tile i's residual add is still streaming to HBM.  The h tile is
    // free by then: D1 of tile i+1 being complete implies conv 2 of tile i (issued earlier) has finished reading it.
    const int q = warp & 3;
    const int n = q * 32 + lane;
    const float bias_h = a.bias_h[n], slope_h = a.slope_h;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t chl = (uint32_t)((q & 1) * 32 + lane);   // channel inside k-block q>>1
    const uint32_t h_thread = smem_u32(smH) + (uint32_t)(q >> 1) * (PAIR_N1 * TC_ROW_BYTES) + (chl & 7u) * 2u;
    const uint32_t chq = chl >> 3;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < rt.total_tiles; tile += gridDim.x) {
      const int t0 = (tile % rt.t_tiles) * PAIR_N;
      long long* dbg = (rt.dbg && blockIdx.x == 0 && q == 0 && lane == 0 && tile / (int)gridDim.x < 16) ? rt.dbg + (tile / gridDim.x) * 8 : nullptr;
      mbar_wait(BAR(iD1F), ph);
      tc_fence_after();
      if (dbg) dbg[4] = clock64();
      // two TMEM loads in flight: chunk c+32 is fetched while chunk c is converted and stored
      float accA[32], accB[32];
      tmem_ld32(taddr, accA);
      auto emit = [&](const float* acc, int c) {
        const int t_abs0 = t0 - PAIR_HALO + c;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = acc[i] + bias_h;
          v = fmaxf(v, v * slope_h);
          const int t_abs = t_abs0 + i;
          if (t_abs < 0 || t_abs >= a.L_out) v = 0.f;
          const uint32_t addr = h_thread + (uint32_t)(c + i) * TC_ROW_BYTES + ((chq ^ (uint32_t)(i & 7)) << 4);
          unsigned short bits;
          if constexpr (Op::kPrec == 3) bits = __half_as_ushort(to_half_sat(v));
          else bits = __bfloat16_as_ushort(__float2bfloat16_rn(v));
          if (c + i < PAIR_N1) asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(bits) : "memory");
        }
      };
#pragma unroll 1
      for (int c = 0; c < PAIR_N1; c += 64) {
        tmem_ld_wait();
        if (c + 32 < PAIR_N1) tmem_ld32(taddr + (uint32_t)(c + 32), accB);
        emit(accA, c);
        if (c + 32 < PAIR_N1) {
          tmem_ld_wait();
          if (c + 64 < PAIR_N1) tmem_ld32(taddr + (uint32_t)(c + 64), accA);
          emit(accB, c + 32);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(iHR));
      if (dbg) dbg[5] = clock64();
      ph ^= 1;
    }
  } else if (warp < NEPI2) {
    // ===================== epilogue-2 warps (8): the ResBlock residual add on D2 =====================
    const int q = warp & 3, grp = warp >> 2;
    const int n = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + PAIR_D2_COL0;
    uint32_t ph = 0;
    int d2 = 0;
    float xcur[32];
    for (int tile = blockIdx.x; tile < rt.total_tiles; tile += gridDim.x) {
      const int tt = tile % rt.t_tiles, b = tile / rt.t_tiles;
      const int t0 = tt * PAIR_N;
      const int t_lim = min(a.L_out, t0 + PAIR_N);
      long long* dbg = (rt.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && tile / (int)gridDim.x < 16) ? rt.dbg + (tile / gridDim.x) * 8 : nullptr;
      {  // ask L2 for the residual rows of this warp's chunks of the NEXT tile (demand loads then hit L2)
        const int tile_n = tile + (int)gridDim.x;
        if (tile_n < rt.total_tiles && a.epi.xin != nullptr) {
          const int bn = tile_n / rt.t_tiles, t0n = (tile_n % rt.t_tiles) * PAIR_N;
          for (int c = grp * 32; c < PAIR_N; c += 64) {
            const int t = t0n + c + lane;
            if (t < a.L_out) {
              const size_t off = ((size_t)bn * a.epi.rows_res + t) * a.epi.ld + (n - lane);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.epi.xin) + off * (RH != 0 ? 2 : 4)));
            }
          }
        }
      }
      mbar_wait(BAR(iD2F + d2), ph);
      tc_fence_after();
      if (dbg) dbg[6] = clock64();
      for (int c = grp * 32; c < PAIR_N; c += 64) {
        float acc[32];
        tmem_ld32(taddr + (uint32_t)(d2 * PAIR_D2_STRIDE + c), acc);
        const int t_first = t0 + c;
        const bool live = t_first < t_lim;
        if (live) epi_prefetch<EPI_RES, LD, RH>(a.epi, b, n, t_first, min(t_lim - t_first, 32), xcur);
        tmem_ld_wait();
        if (live) tc_epilogue32<Op, EPI_RES, LD, RH>(a.epi, b, n, 0, t_first, min(t_lim - t_first, 32), acc, xcur);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(iD2E + d2));
      if (dbg) dbg[7] = clock64();
      if (++d2 == 2) { d2 = 0; ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn();
void* tc_tensormap_encoder() { return reinterpret_cast<void*>(get_encode_fn()); }

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

const char* tc_make_plan(int prec, const ConvArgs& a, int flags, int num_sms, TcPlan* plan) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
  plan->pw = 0;
  plan->no_pair_split = (flags & MBV_FLAG_NO_PAIR_SPLIT) ? 1 : 0;
  if (pw_eligible(prec, a, flags)) return pw_make_plan(prec, a, num_sms, plan);  // 1x1 convs of the WN stacks: pw_tc.cu
  static const int gt_env = getenv("MBV_GATE_TM") ? atoi(getenv("MBV_GATE_TM")) : 1;  // A/B measurements only
  if (gt_env && gt_eligible(prec, a, flags, num_sms)) return gt_make_plan(prec, a, num_sms, plan);  // WN gate convs: pw_tc.cu
  if (ct_eligible(prec, a, flags, num_sms)) return ct_make_plan(prec, a, num_sms, plan);  // k >= 5 convs of a 128-channel ResBlock stage: pw_tc.cu
  const int esize = prec >= 2 ? 2 : 4;
  const int KB = TC_ROW_BYTES / esize;
  if (a.Cp_in % 64 != 0) return "tcgen05 conv: padded input channels must be a multiple of 64";
  const int n_logical = a.N_total;
  if (n_logical % TC_M != 0) return "tcgen05 conv: packed output channels must be a multiple of 128";
  // Time tile: as few tiles as a 256-column accumulator allows, then shrunk to what the length needs.
  // (Measured, profiles/r02_tile_policy_ab.txt: choosing narrower tiles to remove the wave-quantisation tail of the
  //  persistent grid -- e.g. 192 instead of 256 columns at stage 0, 16 rounds instead of 13 -- is 4-6 % SLOWER per
  //  step: every tile re-streams its weight taps L2 -> SMEM -> tensor core, so bytes per FLOP grow as the tile narrows
  //  and the MMA-bound layers lose more than the shorter last round gains.)
  const int tiles = (a.L_out + 255) / 256;
  int n_time = ((a.L_out + tiles - 1) / tiles + 15) / 16 * 16;
  if (n_time < 16) n_time = 16;
  plan->n_time = n_time;
  plan->t_tiles = (a.L_out + n_time - 1) / n_time;
  plan->c_tiles = n_logical / TC_M;
  plan->total_tiles = a.B * a.n_phases * plan->t_tiles * plan->c_tiles;
  const int halo = (a.taps - 1) * a.dil;
  plan->slab_rows = n_time + halo;
  plan->n_boxes = plan->slab_rows > 256 ? 2 : 1;
  plan->box_rows = ((plan->slab_rows + plan->n_boxes - 1) / plan->n_boxes + 7) / 8 * 8;
  if (plan->box_rows > 256) return "tcgen05 conv: activation slab exceeds two 256-row TMA boxes";
  plan->slab_stage_bytes = ((plan->n_boxes * plan->box_rows * TC_ROW_BYTES + 1023) / 1024) * 1024;
  plan->w_stage_bytes = TC_M * TC_ROW_BYTES;
  const int xchg = a.gate ? 32 * 1024 : 0;
  // Staged residual input (see the epilogue warps): two 2 KB TMA stages + two mbarriers per epilogue warp, for the
  // epilogues that read a 2-byte residual tensor (those kernels have no register-load path; measured against it at
  // this commit's parent, profiles/round2_res_stage_ab.txt: every residual-add launch 2-6 % faster, the step 1.7-2 %).
  const bool staged_mode = (a.epi.mode == EPI_RES || a.epi.mode == EPI_RS) && a.epi.res_half != 0 && prec >= 2;
  const bool want_staged = staged_mode;
  if (staged_mode && (a.epi.xin == nullptr || a.epi.row_mul != 1 || (a.epi.mode == EPI_RS && a.epi.n_split < a.N_total)))
    return "tcgen05 conv: a 2-byte residual epilogue needs xin, unit row mapping and residual rows only";
  const int kStagedWarps = EpiWarps<EPI_RES>::value;
  static_assert(EpiWarps<EPI_RES>::value == EpiWarps<EPI_RS>::value, "staged residual rings are sized for both modes");
  const bool staged_xs = want_staged && a.epi.mode == EPI_RES && (a.epi.sum_mode == 2 || a.epi.sum_mode == 3) && a.epi.xs != nullptr;
  if (want_staged && a.epi.mode == EPI_RES && (a.epi.sum_mode == 2 || a.epi.sum_mode == 3) && !staged_xs)
    return "tcgen05 conv: summing residual epilogue without a running-sum buffer";
  const int res_bytes = want_staged ? kStagedWarps * 2048 * (staged_xs ? 2 : 1) : 0;
  const int extra_bars = staged_mode ? 2 * kStagedWarps : 0;
  // Bytes in flight are what hide the ~1.3 us L2->SMEM TMA round trip (measured: with 5 x 16 KB weight stages the MMA
  // warp found its weights missing on 45 % of its waits): give the weight ring everything two slab stages leave over.
  const int budget = 222 * 1024 - xchg - res_bytes;
  // A slab lasts taps x 512 tensor-core cycles: few taps need more slabs in flight, many taps more weight stages.
  plan->n_slab_stages = a.taps == 1 ? 4 : (a.taps <= 4 ? 3 : 2);
  while (plan->n_slab_stages > 2 &&
         (budget - plan->n_slab_stages * plan->slab_stage_bytes) / plan->w_stage_bytes < (a.taps <= 4 ? 4 : 6))
    plan->n_slab_stages--;
  plan->n_w_stages = (budget - plan->n_slab_stages * plan->slab_stage_bytes) / plan->w_stage_bytes;
  if (plan->n_w_stages > 10) plan->n_w_stages = 10;
  if (plan->n_w_stages < 2) return "tcgen05 conv: not enough shared memory for two weight stages";
  const int nbars = 2 * plan->n_slab_stages + 2 * plan->n_w_stages + 4 + extra_bars;
  plan->xchg_off = (plan->n_slab_stages * plan->slab_stage_bytes + plan->n_w_stages * plan->w_stage_bytes + nbars * 8 + 16 +
                    127) / 128 * 128;
  plan->res_off = want_staged ? plan->xchg_off + xchg : 0;
  plan->smem_bytes = 1024 + plan->xchg_off + xchg + res_bytes;
  if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns
  plan->grid = plan->total_tiles < num_sms ? plan->total_tiles : num_sms;
  if (plan->grid < 1) plan->grid = 1;

  plan->cluster = 0;
  // CTA pairs (cta_group::2): the two CTAs of a cluster take the two channel tiles of ONE time tile.  One MMA spans both
  // (M = 256); each CTA feeds its own 128 weight rows and only HALF of the activation rows from its shared memory, so the
  // operand bytes read per MMA and CTA fall from 12 KB to 8 KB -- the single-CTA MMAs are bound by the shared-memory port
  // (operand reads + TMA writes ~ 107 B/clk of 128), see DESIGN.md.  Needs an even number of channel tiles.
  // Measured (profiles/round2_cta_pairs_ab.txt): stage-0 k=11 convs 245 -> 210 us (1295 -> 1520 TFLOP/s), k=7 156 -> 145,
  // first upsampler 207 -> 180, conv_pre 74 -> 66.  The k=3 residual convs were slightly slower in pairs when their residual
  // came by register loads (113 -> 125 us) and stayed single-CTA; with the TMA-staged residual they gain too (re-measured at
  // the end of round 2, profiles/round2_res_pair_taps_ab.txt: c2 / c1 time ratio 1.21 -> 1.16, the summing conv 1.12 -> 1.02),
  // so every multi-tap RES epilogue pairs up now (MBV_RES_PAIR_TAPS=5 restores the old split).
  static const int res_pair_taps = getenv("MBV_RES_PAIR_TAPS") ? atoi(getenv("MBV_RES_PAIR_TAPS")) : 2;  // A/B measurements only
  const bool pair_mode = (a.epi.mode == EPI_ACT && a.taps > 1) || (a.epi.mode == EPI_RES && a.taps >= res_pair_taps);
  // A pair is two channel tiles of one time tile or -- polyphase upsamplers with an odd number of channel tiles -- two
  // branches that read the same input rows (k16 / stride 4: branches (0,1) and (2,3) have the same first input row).
  bool phase_pairs = (plan->c_tiles % 2 != 0) && (a.n_phases % 2 == 0);
  for (int r = 0; phase_pairs && r + 1 < a.n_phases; r += 2) phase_pairs = a.shift0[r] == a.shift0[r + 1];
  plan->pair_phase = 0;
  if (!(flags & MBV_FLAG_NO_CTA_PAIRS) && prec >= 2 && pair_mode && (plan->c_tiles % 2 == 0 || phase_pairs) && num_sms >= 2 &&
      n_time % 16 == 0) {
    plan->cluster = 2;
    plan->pair_phase = (plan->c_tiles % 2 == 0) ? 0 : 1;
    plan->slab_rows = n_time / 2 + halo;
    plan->n_boxes = 1;
    plan->box_rows = (plan->slab_rows + 7) / 8 * 8;
    plan->slab_stage_bytes = ((plan->box_rows * TC_ROW_BYTES + 1023) / 1024) * 1024;
    plan->n_slab_stages = a.taps <= 4 ? 4 : 3;
    plan->n_w_stages = (budget - plan->n_slab_stages * plan->slab_stage_bytes) / plan->w_stage_bytes;
    if (plan->n_w_stages > 10) plan->n_w_stages = 10;
    const int nb2 = 2 * plan->n_slab_stages + 2 * plan->n_w_stages + 4 + extra_bars;
    plan->xchg_off = (plan->n_slab_stages * plan->slab_stage_bytes + plan->n_w_stages * plan->w_stage_bytes + nb2 * 8 + 16 + 127) / 128 * 128;
    plan->res_off = want_staged ? plan->xchg_off + xchg : 0;
    plan->smem_bytes = 1024 + plan->xchg_off + xchg + res_bytes;
    if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;
    plan->grid = (plan->total_tiles < num_sms ? plan->total_tiles : num_sms) & ~1;  // whole pairs (total_tiles is even)
  }
  plan->rows = a.B * plan->t_tiles;
  plan->groups = a.n_phases * plan->c_tiles;

  const CUtensorMapDataType dt = prec == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                 : (prec == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.L_in, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.x_ld * esize, (cuuint64_t)a.L_in * a.x_ld * esize};
    cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)plan->box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&plan->tmA, dt, 3, const_cast<void*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the activation map";
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.n_phases * a.taps * a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * esize};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)TC_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&plan->tmB, dt, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the weight map";
  }
  // small layers (stage-1 k=3 convs, flow post): the whole weight set of the channel tile stays in shared memory
  plan->w_resident = (plan->cluster == 0 && plan->c_tiles == 1 && a.n_phases == 1 && a.taps * (a.Cp_in / KB) <= plan->n_w_stages) ? 1 : 0;
  // Everything else streams its weight taps L2 -> SMEM once per tile and is feed-bound for few taps (k=3 / upsampler /
  // gate convs issue MMAs at ~70 % inside a tile): run clusters of two CTAs on two time tiles of the same weight group,
  // each CTA fetching half of every weight tile and multicasting it to both.
  // Measured (same box, interleaved, profiles/r02_cluster_ab.txt): the MMA-bound layers gain 3-8 %, the step as a whole
  // < 1 % -- so the feed is not what holds the k=3 / upsampler convs at ~70 % issue efficiency -- hence opt-in.
  const int use_cluster = (flags & MBV_FLAG_CLUSTER_PAIRS) ? 1 : 0;
  // (1x1 convs -- flow pre / res / post -- are epilogue- or launch-bound and measured 5-12 % slower in pairs: taps > 1 only)
  const bool mode_ok = a.epi.mode == EPI_ACT || a.epi.mode == EPI_RES || a.epi.mode == EPI_F32 || a.epi.mode == EPI_GATE;
  if (plan->cluster == 0 && use_cluster && prec >= 2 && mode_ok && !plan->w_resident && (num_sms & 1) == 0 && plan->rows >= 2 && a.taps > 1) {
    const long long pairs = (long long)plan->groups * ((plan->rows + 1) / 2);
    if (pairs >= num_sms / 2) {
      plan->cluster = 1;
      plan->total_tiles = (int)(2 * pairs);
      plan->grid = num_sms;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.n_phases * a.taps * a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * esize};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)(TC_M / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&plan->tmBh, dt, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the half-tile weight map";
  }
  plan->prefetch_res = 0;
  if ((a.epi.mode == EPI_RES || want_staged) && a.epi.xin != nullptr && a.epi.row_mul == 1) {
    const int rs = a.epi.res_half ? 2 : 4;
    cuuint64_t dims[3] = {(cuuint64_t)a.epi.ld, (cuuint64_t)a.epi.rows_res, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.epi.ld * rs, (cuuint64_t)a.epi.rows_res * a.epi.ld * rs};
    cuuint32_t box[3] = {(cuuint32_t)(a.epi.ld < TC_M ? a.epi.ld : TC_M), (cuuint32_t)n_time, 1};
    if (want_staged) { box[0] = 32; box[1] = 32; }  // one epilogue warp's chunk: 32 channels (64 B) x 32 rows
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&plan->tmR, a.epi.res_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<void*>(a.epi.xin), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) plan->prefetch_res = 1;  // a failed encode only loses the prefetch
  }
  if (!plan->prefetch_res) {
    if (want_staged) return "cuTensorMapEncodeTiled failed for the residual map";
    plan->tmR = plan->tmA;  // placeholder, never dereferenced
  }
  plan->tmS = plan->tmR;
  if (plan->res_off > 0 && staged_xs) {  // the running ResBlock sum: same geometry as the residual tensor
    cuuint64_t dims[3] = {(cuuint64_t)a.epi.ld, (cuuint64_t)a.epi.rows_res, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.epi.ld * 2, (cuuint64_t)a.epi.rows_res * a.epi.ld * 2};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&plan->tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a.epi.xs, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the running-sum map";
  }
  return nullptr;
}

// kernel table: (mode, compile-time pitch) instantiations; pitch 0 = runtime
// CL = 1: the cluster-pair variant (weight tiles multicast between two CTAs); compiled only where a plan can ask for it
// (16-bit operands, multi-tap convs: ACT / RES / F32 / GATE epilogues)
template <typename Op, int MODE, int LD, int RH = 0>
static cudaError_t launch_one(const ConvArgs& a, const TcPlan& p, const TcRt& rt, cudaStream_t st, bool set_attr);

template <typename Op, int MODE, int LD, int RH, int CL>
static cudaError_t launch_one_cl(const ConvArgs& a, const TcPlan& p, const TcRt& rt, cudaStream_t st, bool set_attr) {
  auto k = conv_tc_kernel<Op, MODE, LD, RH, CL>;
  if (set_attr) return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(TcThreads<MODE>::value);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[0].val.programmaticStreamSerializationAllowed = rt.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (rt.cluster) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, k, p.tmA, p.tmB, p.tmR, p.tmBh, p.tmS, a, rt);
}

template <typename Op, int MODE, int LD, int RH>
static cudaError_t launch_one(const ConvArgs& a, const TcPlan& p, const TcRt& rt, cudaStream_t st, bool set_attr) {
  constexpr bool kHasCluster = (Op::kPrec >= 2) && (MODE == EPI_ACT || MODE == EPI_RES || MODE == EPI_F32 || MODE == EPI_GATE);
  constexpr bool kHasPairs = (Op::kPrec >= 2) && (MODE == EPI_ACT || MODE == EPI_RES);   // cta_group::2 variant
  if constexpr (kHasPairs) {
    if (set_attr) {
      cudaError_t e = launch_one_cl<Op, MODE, LD, RH, 2>(a, p, rt, st, true);
      if (e != cudaSuccess) return e;
    } else if (rt.cluster == 2) {
      return launch_one_cl<Op, MODE, LD, RH, 2>(a, p, rt, st, false);
    }
  } else {
    if (!set_attr && rt.cluster == 2) return cudaErrorInvalidValue;
  }
  if constexpr (kHasCluster) {
    if (set_attr) {
      cudaError_t e = launch_one_cl<Op, MODE, LD, RH, 1>(a, p, rt, st, true);
      if (e != cudaSuccess) return e;
    } else if (rt.cluster) {
      return launch_one_cl<Op, MODE, LD, RH, 1>(a, p, rt, st, false);
    }
  } else {
    if (!set_attr && rt.cluster) return cudaErrorInvalidValue;
  }
  return launch_one_cl<Op, MODE, LD, RH, 0>(a, p, rt, st, set_attr);
}

template <typename Op>
static cudaError_t dispatch(const ConvArgs& a, const TcPlan& p, const TcRt& rt, cudaStream_t st, bool set_attr, int mode,
                            int ld, int res_half) {
  if constexpr (Op::kPrec == 3) {
    if (res_half == 2) {  // single stream: ResBlock residual adds (decoder) and WN residual adds (flow)
#define MBV_CASE_S(M, L) if (mode == M && ld == L) return launch_one<Op, M, L, 2>(a, p, rt, st, set_attr);
      MBV_CASE_S(EPI_RES, 128) MBV_CASE_S(EPI_RES, 256) MBV_CASE_S(EPI_RS, 192)
#undef MBV_CASE_S
      if (mode == EPI_RES) return launch_one<Op, EPI_RES, 0, 2>(a, p, rt, st, set_attr);
      if (mode == EPI_RS) return launch_one<Op, EPI_RS, 0, 2>(a, p, rt, st, set_attr);
      return cudaErrorInvalidValue;
    }
  }
  if (res_half == 1) {  // fp16 residual stream: decoder ACT (upsampler) and RES (ResBlock) epilogues only
    if constexpr (Op::kPrec == 2) {
#define MBV_CASE_H(M, L) if (mode == M && ld == L) return launch_one<Op, M, L, 1>(a, p, rt, st, set_attr);
      MBV_CASE_H(EPI_ACT, 128) MBV_CASE_H(EPI_ACT, 256) MBV_CASE_H(EPI_RES, 128) MBV_CASE_H(EPI_RES, 256)
      MBV_CASE_H(EPI_ACT, 192) MBV_CASE_H(EPI_RS, 192)
#undef MBV_CASE_H
      if (mode == EPI_ACT) return launch_one<Op, EPI_ACT, 0, 1>(a, p, rt, st, set_attr);
      if (mode == EPI_RES) return launch_one<Op, EPI_RES, 0, 1>(a, p, rt, st, set_attr);
      if (mode == EPI_RS) return launch_one<Op, EPI_RS, 0, 1>(a, p, rt, st, set_attr);
    }
    return cudaErrorInvalidValue;
  }
  if (res_half != 0) return cudaErrorInvalidValue;
#define MBV_CASE(M, L) if (mode == M && ld == L) return launch_one<Op, M, L>(a, p, rt, st, set_attr);
  MBV_CASE(EPI_ACT, 128) MBV_CASE(EPI_ACT, 256) MBV_CASE(EPI_ACT, 192)
  MBV_CASE(EPI_RES, 128) MBV_CASE(EPI_RES, 256)
  MBV_CASE(EPI_GATE, 768) MBV_CASE(EPI_RS, 192) MBV_CASE(EPI_POST, 192)
#undef MBV_CASE
  switch (mode) {
    case EPI_ACT: return launch_one<Op, EPI_ACT, 0>(a, p, rt, st, set_attr);
    case EPI_RES: return launch_one<Op, EPI_RES, 0>(a, p, rt, st, set_attr);
    case EPI_F32: return launch_one<Op, EPI_F32, 0>(a, p, rt, st, set_attr);
    case EPI_GATE: return launch_one<Op, EPI_GATE, 0>(a, p, rt, st, set_attr);
    case EPI_RS: return launch_one<Op, EPI_RS, 0>(a, p, rt, st, set_attr);
    default: return launch_one<Op, EPI_POST, 0>(a, p, rt, st, set_attr);
  }
}

cudaError_t tc_set_attributes() {
  ConvArgs a{};
  TcPlan p{};
  TcRt rt{};
  const int lds[] = {0, 128, 192, 256, 768};
  for (int mode = EPI_ACT; mode <= EPI_POST; ++mode)
    for (int ld : lds) {
      cudaError_t e = dispatch<OpBF16>(a, p, rt, nullptr, true, mode, ld, 0);
      if (e != cudaSuccess) return e;
      e = dispatch<OpTF32>(a, p, rt, nullptr, true, mode, ld, 0);
      if (e != cudaSuccess) return e;
      e = dispatch<OpF16>(a, p, rt, nullptr, true, mode, ld, 0);
      if (e != cudaSuccess) return e;
      if (mode == EPI_ACT || mode == EPI_RES || mode == EPI_RS) {
        e = dispatch<OpBF16>(a, p, rt, nullptr, true, mode, ld, 1);
        if (e != cudaSuccess) return e;
      }
      if (mode == EPI_RES || mode == EPI_RS) {
        e = dispatch<OpF16>(a, p, rt, nullptr, true, mode, ld, 2);
        if (e != cudaSuccess) return e;
      }
    }
  return pw_set_attributes();
}

// MBV_TIMELINE=<mode>: after every launch whose epilogue mode matches, print CTA 0's per-tile clock stamps (debug)
static void timeline_dump(const ConvArgs& a, const TcPlan& p, long long* dbg, cudaStream_t st) {
  cudaStreamSynchronize(st);
  std::vector<long long> hbuf(7 * 64);
  cudaMemcpy(hbuf.data(), dbg, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  const int nt = (p.total_tiles + p.grid - 1) / p.grid;
  long long t0 = hbuf[0];
  fprintf(stderr, "[timeline] mode %d taps %d Cp_in %d N_total %d n_time %d tiles/CTA %d slab_st %d w_st %d\n", a.epi.mode, a.taps,
          a.Cp_in, a.N_total, p.n_time, nt, p.n_slab_stages, p.n_w_stages);
  for (int i = 0; i < nt && i < 32; ++i)
    fprintf(stderr, "  tile %2d  prod_start %7lld | mma %7lld .. %7lld (waits: acc %5lld slab %5lld weights %5lld) | epi %7lld .. %7lld\n", i,
            hbuf[2 * i] - t0, hbuf[64 + 2 * i] - t0, hbuf[64 + 2 * i + 1] - t0, i < 16 ? hbuf[384 + 3 * i] : 0, i < 16 ? hbuf[384 + 3 * i + 1] : 0,
            i < 16 ? hbuf[384 + 3 * i + 2] : 0, hbuf[128 + 2 * i] - t0, hbuf[128 + 2 * i + 1] - t0);
  for (int i = 0; i < nt && i < 10; ++i)
    for (int j = 0; j < 3; ++j) {
      const long long* c = &hbuf[192 + (i * 3 + j) * 4];
      if (c[0] == 0) continue;
      fprintf(stderr, "    tile %2d chunk %d  start %7lld | acc +%5lld | residual +%5lld | done +%5lld\n", i, j, c[0] - t0,
              c[1] - c[0], c[2] ? c[2] - c[0] : 0, c[3] - c[0]);
    }
  cudaMemset(dbg, 0, 148 * 7 * 64 * sizeof(long long));
}

cudaError_t launch_conv_tc(int prec, const ConvArgs& a, const TcPlan& p, cudaStream_t st, int pdl) {
  if (p.pw) return launch_pw(prec, a, p, st, pdl);
  static long long* dbg = nullptr;
  static int dbg_mode = -2;
  if (dbg_mode == -2) {
    const char* e = getenv("MBV_TIMELINE");
    dbg_mode = e ? atoi(e) : -1;
    if (dbg_mode >= 0) { cudaMalloc(&dbg, 148 * 7 * 64 * sizeof(long long)); cudaMemset(dbg, 0, 148 * 7 * 64 * sizeof(long long)); }
  }
  TcRt rt;
  rt.pdl = pdl;
  static const int no_pack2 = getenv("MBV_NO_PACK2") ? atoi(getenv("MBV_NO_PACK2")) : 0;  // A/B measurements only
  rt.pack2 = no_pack2 ? 0 : 1;
  rt.dbg = (dbg_mode >= 0 && a.epi.mode == dbg_mode) ? dbg : nullptr;
  rt.n_time = p.n_time; rt.slab_rows = p.slab_rows; rt.box_rows = p.box_rows; rt.n_boxes = p.n_boxes;
  rt.slab_stage_bytes = p.slab_stage_bytes; rt.w_stage_bytes = p.w_stage_bytes;
  rt.n_slab_stages = p.n_slab_stages; rt.n_w_stages = p.n_w_stages;
  rt.t_tiles = p.t_tiles; rt.c_tiles = p.c_tiles; rt.total_tiles = p.total_tiles;
  rt.prefetch_res = p.prefetch_res;
  rt.xchg_off = p.xchg_off;
  rt.res_off = p.res_off;
  rt.w_resident = p.w_resident;
  rt.rotate = (p.c_tiles == 2 && (p.grid & 1) == 0 && p.total_tiles > p.grid) ? 1 : 0;
  rt.cluster = p.cluster; rt.rows = p.rows; rt.groups = p.groups;
  rt.pair_phase = p.pair_phase;
  if (p.cluster) rt.rotate = 0;
  rt.split_from = p.total_tiles; rt.split_k = 1; rt.split_n = p.n_time; rt.virt_tiles = p.total_tiles;
  static const int no_split = getenv("MBV_NO_SPLIT") ? atoi(getenv("MBV_NO_SPLIT")) : 0;  // A/B measurements only
  static const int no_pair_split = getenv("MBV_NO_PAIR_SPLIT") ? atoi(getenv("MBV_NO_PAIR_SPLIT")) : 0;  // A/B measurements only
  if (!no_split && !no_pair_split && !p.no_pair_split && p.cluster == 2 && (p.total_tiles & 1) == 0 && (p.grid & 1) == 0) {
    // CTA pairs: the same split in units of pairs (stage 0: 896 pair tiles on 74 pairs = 12.1 rounds of work in 13 rounds;
    // the 8 leftover pair tiles become 32 pieces of 64 columns: 12.25 rounds)
    const int pairs = p.total_tiles / 2, gp = p.grid / 2, remp = pairs % gp;
    if (pairs > gp && remp > 0 && 2 * remp <= gp) {
      int k = gp / remp;
      const int kmax = p.n_time / 64;
      if (k > kmax) k = kmax;
      if (k >= 2) {
        const int n_sub = ((p.n_time + k - 1) / k + 31) / 32 * 32;   // each CTA stages half of a piece's rows
        const int ksub = (p.n_time + n_sub - 1) / n_sub;
        if (ksub >= 2 && remp * ksub <= gp && p.n_time % 32 == 0) {
          rt.split_from = pairs - remp; rt.split_k = ksub; rt.split_n = n_sub;   // (pair units)
          rt.virt_tiles = 2 * (rt.split_from + remp * ksub);
        }
      }
    }
  }
  const int rem = p.total_tiles % p.grid;
  if (!no_split && !p.cluster && p.total_tiles > p.grid && rem > 0 && 2 * rem <= p.grid) {
    int k = p.grid / rem;                          // pieces per tile so that the last round still fits the grid
    const int kmax = p.n_time / 64;                // pieces of >= 64 columns
    if (k > kmax) k = kmax;
    if (k >= 2) {
      const int n_sub = ((p.n_time + k - 1) / k + 15) / 16 * 16;
      const int ksub = (p.n_time + n_sub - 1) / n_sub;
      if (ksub >= 2 && rem * ksub <= p.grid) {
        rt.split_from = p.total_tiles - rem; rt.split_k = ksub; rt.split_n = n_sub;
        rt.virt_tiles = rt.split_from + rem * ksub;
      }
    }
  }
  cudaError_t e;
  if (prec == 3) e = dispatch<OpF16>(a, p, rt, st, false, a.epi.mode, a.epi.ld, a.epi.res_half);
  else if (prec == 2) e = dispatch<OpBF16>(a, p, rt, st, false, a.epi.mode, a.epi.ld, a.epi.res_half);
  else if (a.epi.res_half) e = cudaErrorInvalidValue;
  else e = dispatch<OpTF32>(a, p, rt, st, false, a.epi.mode, a.epi.ld, 0);
  if (rt.dbg && e == cudaSuccess) timeline_dump(a, p, rt.dbg, st);
  return e;
}


// ------------------------------------------------------------------------------------------------
// fused ResBlock conv pair: plan + launch
// ------------------------------------------------------------------------------------------------
const char* tc_make_pair_plan(int prec, const ConvArgs& a, int num_sms, TcPairPlan* plan) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
  if (prec < 2) return "conv pair: 16-bit operand types only";
  if (a.Cp_in != 128 || a.N_total != 128 || a.n_phases != 1) return "conv pair: 128 channels, stride 1 only";
  if ((a.taps - 1) / 2 > PAIR_HALO || (a.taps & 1) == 0) return "conv pair: odd kernel size <= 17";
  const int slab_rows = PAIR_N1 + (a.taps - 1) * a.dil;
  plan->n_boxes = slab_rows > 256 ? 2 : 1;
  plan->box_rows = ((slab_rows + plan->n_boxes - 1) / plan->n_boxes + 7) / 8 * 8;
  if (plan->box_rows > 256) return "conv pair: activation slab exceeds two 256-row TMA boxes";
  plan->slab_stage_bytes = ((plan->n_boxes * plan->box_rows * TC_ROW_BYTES + 1023) / 1024) * 1024;
  plan->n_slab_stages = 2;
  plan->h_off = plan->n_slab_stages * plan->slab_stage_bytes;
  plan->w_off = plan->h_off + PAIR_H_BYTES;
  const int budget = 225 * 1024 - plan->w_off;
  plan->n_w_stages = budget / (TC_M * TC_ROW_BYTES);
  if (plan->n_w_stages > 10) plan->n_w_stages = 10;
  if (plan->n_w_stages < 3) return "conv pair: not enough shared memory for the weight ring";
  plan->bar_off = plan->w_off + plan->n_w_stages * TC_M * TC_ROW_BYTES;
  const int nbars = 2 * plan->n_slab_stages + 2 * plan->n_w_stages + 6;
  plan->smem_bytes = 1024 + plan->bar_off + nbars * 8 + 16;
  plan->t_tiles = (a.L_out + PAIR_N - 1) / PAIR_N;
  plan->total_tiles = a.B * plan->t_tiles;
  plan->grid = plan->total_tiles < num_sms ? plan->total_tiles : num_sms;
  const CUtensorMapDataType dt = prec == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.L_in, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.x_ld * 2, (cuuint64_t)a.L_in * a.x_ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)plan->box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&plan->tmA, dt, 3, const_cast<void*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the activation map";
  }
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.taps * a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)TC_M};
    cuuint32_t estr[2] = {1, 1};
    if (enc(which ? &plan->tmB2 : &plan->tmB, dt, 2, const_cast<void*>(which ? a.w2 : a.w), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for a weight map";
  }
  return nullptr;
}

template <typename Op, int LD, int RH>
static cudaError_t launch_pair_one(const ConvArgs& a, const TcPairPlan& p, const PairRt& rt, cudaStream_t st, bool set_attr) {
  auto k = conv_pair_kernel<Op, LD, RH>;
  if (set_attr) return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(TcThreads<EPI_RES>::value);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = rt.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, p.tmA, p.tmB, p.tmB2, a, rt);
}

static cudaError_t pair_dispatch(int prec, const ConvArgs& a, const TcPairPlan& p, const PairRt& rt, cudaStream_t st, bool set_attr) {
  const int ld = a.epi.ld, rh = a.epi.res_half;
  if (prec == 2 && rh == 1) return ld == 128 ? launch_pair_one<OpBF16, 128, 1>(a, p, rt, st, set_attr) : launch_pair_one<OpBF16, 0, 1>(a, p, rt, st, set_attr);
  if (prec == 2 && rh == 0) return launch_pair_one<OpBF16, 0, 0>(a, p, rt, st, set_attr);
  if (prec == 3 && rh == 2) return ld == 128 ? launch_pair_one<OpF16, 128, 2>(a, p, rt, st, set_attr) : launch_pair_one<OpF16, 0, 2>(a, p, rt, st, set_attr);
  return cudaErrorInvalidValue;
}

cudaError_t tc_pair_set_attributes() {
  ConvArgs a{};
  TcPairPlan p{};
  PairRt rt{};
  const int combos[5][3] = {{2, 128, 1}, {2, 0, 1}, {2, 0, 0}, {3, 128, 2}, {3, 0, 2}};
  for (auto& c : combos) {
    a.epi.ld = c[1];
    a.epi.res_half = c[2];
    cudaError_t e = pair_dispatch(c[0], a, p, rt, nullptr, true);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_conv_pair(int prec, const ConvArgs& a, const TcPairPlan& p, cudaStream_t st, int pdl) {
  PairRt rt;
  rt.pdl = pdl;
  rt.box_rows = p.box_rows; rt.n_boxes = p.n_boxes; rt.slab_stage_bytes = p.slab_stage_bytes;
  rt.n_slab_stages = p.n_slab_stages; rt.n_w_stages = p.n_w_stages; rt.t_tiles = p.t_tiles; rt.total_tiles = p.total_tiles;
  rt.h_off = p.h_off; rt.w_off = p.w_off; rt.bar_off = p.bar_off;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("MBV_TIMELINE");
    dbg_on = (e && atoi(e) == 9) ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, 16 * 8 * sizeof(long long)); cudaMemset(dbg, 0, 16 * 8 * sizeof(long long)); }
  }
  rt.dbg = dbg_on ? dbg : nullptr;
  cudaError_t e = pair_dispatch(prec, a, p, rt, st, false);
  if (dbg_on && e == cudaSuccess) {
    cudaStreamSynchronize(st);
    long long hb[16 * 8];
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[pair timeline] k%d d%d slab_rows/box %d x%d w_stages %d tiles/CTA %d\n", a.taps, a.dil, p.box_rows, p.n_boxes,
            p.n_w_stages, (p.total_tiles + p.grid - 1) / p.grid);
    const long long t0 = hb[0];
    for (int i = 0; i < 8; ++i)
      fprintf(stderr, "  tile %2d  mma1 %7lld..%7lld | mma2 %7lld..%7lld | epi1 %7lld..%7lld | epi2 %7lld..%7lld\n", i, hb[8 * i] - t0,
              hb[8 * i + 1] - t0, hb[8 * i + 2] - t0, hb[8 * i + 3] - t0, hb[8 * i + 4] - t0, hb[8 * i + 5] - t0, hb[8 * i + 6] - t0,
              hb[8 * i + 7] - t0);
    cudaMemset(dbg, 0, sizeof(hb));
  }
  return e;
}

}  // namespace mbv
