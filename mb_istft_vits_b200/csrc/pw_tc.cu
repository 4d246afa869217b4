// pw_tc.cu -- pointwise (1x1) convs of the WaveNet stacks with TIME on the accumulator lane (sm_100a).
//
//   D[128 rows, N channels] = X[128 rows, K] . W[N, K]^T          rows = the flattened (utterance, time) axis, R = B * L
//
// The 1x1 convs of the flow / posterior encoder (ResidualCouplingLayer.pre, modules.py:328,341; the residual half of
// WN.res_skip_layers, modules.py:135-146,169-175) are a few GFLOP each: they are bound by their epilogue, not by MMAs.
// In conv_tc_kernel (output CHANNEL on the lane) an epilogue thread owns one channel of 32 time steps, so every load /
// store is a 2-byte access per thread (64 bytes per warp instruction), the 192 output channels fill one and a half
// 128-row channel tiles, and the per-row mask costs 32 loads per chunk.  Here the roles of the operands are swapped:
//   * A = the activation tile (128 consecutive rows x 64 channels per k-block, TMA, 128B swizzle), B = ALL weight rows of
//     the layer (N <= 256, resident in shared memory for the life of the persistent CTA), tcgen05.mma M 128, N = C_out.
//   * an epilogue thread owns one ROW: its 32 channels of a chunk are 64 contiguous bytes of the channels-last tensors, so
//     the residual input arrives as two 256-bit loads and every output leaves as two 256-bit stores (full 32-byte
//     sectors, 16 x fewer memory instructions), the mask is ONE value per thread and tile, the bias a broadcast read.
//   * a 1x1 conv has no halo: tiles are 128 rows of the flattened [B * L] axis and may straddle utterances (the mask and
//     every address are per row), so 55 168 rows make 431 full tiles instead of 64 x 7 ragged ones.
//   * warp roles as in conv_tc.cu: 8 epilogue warps (lane quarter x column half), TMA producer, MMA issuer; two
//     accumulator stages in TMEM, 2-4 activation stages.
// Epilogues (same arithmetic and operation order as conv_tc.cu): ACT  y = (acc + bias) [* mask] -> fp16 stream copy +
// leaky-relu'd operand copy;  RS  x = (xin + acc + bias) * mask -> fp16 stream + operand copy (or, single stream, the
// fp16 operand tensor only).  Anything else (per-utterance bias, fp32 streams, K > 256, tf32) stays on conv_tc_kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mbistft.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace mbv {

constexpr int PW_ROWS = 128;                       // rows per tile (UMMA M)
constexpr int PW_EPI_WARPS = 8;
constexpr int PW_WARP_TMA = PW_EPI_WARPS, PW_WARP_MMA = PW_EPI_WARPS + 1;
constexpr int PW_THREADS = 32 * (PW_EPI_WARPS + 2);
constexpr int PW_ACC_STRIDE = 256;                 // TMEM columns per accumulator stage
constexpr int PW_KB_BYTES = PW_ROWS * TC_ROW_BYTES;  // one k-block of an activation tile (16 KB)
constexpr int PW_MAX_CHUNKS = 4;                   // 32-channel chunks per epilogue warp (N <= 256)

struct PwRt {
  int R;          // rows in total
  int kblocks;    // Cp_in / 64
  int N;          // output channels = UMMA N (multiple of 32, <= 256)
  int n_tiles;
  int n_a_stages, w_bytes;     // ring of 16 KB k-block stages; bytes of the resident weights
  int a_off, bias_off, bar_off;
  int res_c0, res_cols, res_box;   // L2 prefetch of the residual rows: first channel, channels, box width (elements)
};

__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
  const __half2 h = __halves2half2(to_half_sat(a), to_half_sat(b));
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <typename Op> __device__ __forceinline__ uint32_t pack_op2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack_op2<OpBF16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack_op2<OpF16>(float a, float b) { return pack_half2_sat(a, b); }

// MODE: EPI_ACT or EPI_RS.  RH: 1 = fp16 stream (xin / xout) next to the operand copy, 2 = single stream (fp16 operands: xin is
// the operand tensor, no xout), 0 = ACT without a stream copy.
template <typename Op, int MODE, int RH>
__global__ void __launch_bounds__(PW_THREADS, 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmR,
             const EpiParams p, const PwRt rt) {
  using T = typename Op::T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smW = smem;                  // kblocks tiles of N rows x 128 B
  uint8_t* smA = smem + rt.a_off;       // n_a_stages x kblocks x 16 KB
  float* s_bias = reinterpret_cast<float*>(smem + rt.bias_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + rt.bar_off);
  const int iWF = 0, iAF = 1, iAE = iAF + rt.n_a_stages, iCF = iAE + rt.n_a_stages, iCE = iCF + 2, nBars = iCE + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + nBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    if constexpr (MODE != EPI_ACT) tma_prefetch_desc(&tmR);
    mbar_init(BAR(iWF), 1);
    for (int i = 0; i < rt.n_a_stages; ++i) { mbar_init(BAR(iAF + i), 1); mbar_init(BAR(iAE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(iCF + i), 1); mbar_init(BAR(iCE + i), PW_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == PW_WARP_MMA) tmem_alloc(smem_u32(tmem_ptr_smem), 512u);
  // weights and bias are constants of the model (never written by a kernel of the step): fetched before the
  // programmatic-dependency wait, so they overlap the tail of the previous kernel
  for (int i = threadIdx.x; i < rt.N; i += PW_THREADS) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == PW_WARP_TMA) {
    if (elect_one()) {
      mbar_expect_tx(BAR(iWF), (uint32_t)rt.w_bytes);
      for (int kb = 0; kb < rt.kblocks; ++kb)
        tma_load_2d(smem_u32(smW) + (uint32_t)(kb * rt.N * TC_ROW_BYTES), &tmW, BAR(iWF), kb * 64, 0);
    }
    __syncwarp();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == PW_WARP_TMA) {
    // ===================== TMA producer: one k-block (128 rows x 64 channels, 16 KB) of an activation tile per ring stage
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < rt.n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < rt.kblocks; ++kb) {
        mbar_wait(BAR(iAE + s), ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(BAR(iAF + s), (uint32_t)PW_KB_BYTES);
          tma_load_2d(smem_u32(smA) + (uint32_t)(s * PW_KB_BYTES), &tmX, BAR(iAF + s), kb * 64, tile * PW_ROWS);
          if (MODE != EPI_ACT && kb == 0) {
            // the residual rows of this tile: asked into L2 now (the producer runs ahead of the epilogue), so the epilogue's
            // 256-bit loads find them there instead of paying a DRAM round trip per tile
            for (int c = 0; c < rt.res_cols; c += rt.res_box) tma_prefetch_l2_2d(&tmR, rt.res_c0 + c, tile * PW_ROWS);
          }
        }
        __syncwarp();
        if (++s == rt.n_a_stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == PW_WARP_MMA) {
    // ===================== MMA issuer =====================
    constexpr uint32_t fmt = (Op::kPrec == 3) ? 0u : 1u;  // F16 / BF16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(rt.N >> 3) << 17) | ((uint32_t)(PW_ROWS >> 4) << 24);
    int s = 0, sc = 0;
    uint32_t ph = 0, pc = 0;
    mbar_wait(BAR(iWF), 0);
    tc_fence_after();
    for (int tile = blockIdx.x; tile < rt.n_tiles; tile += gridDim.x) {
      mbar_wait(BAR(iCE + sc), pc ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(sc * PW_ACC_STRIDE);
      for (int kb = 0; kb < rt.kblocks; ++kb) {
        mbar_wait(BAR(iAF + s), ph);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem_u32(smA) + (uint32_t)(s * PW_KB_BYTES));
        const uint32_t w_lo = desc_lo(smem_u32(smW) + (uint32_t)(kb * rt.N * TC_ROW_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma<2>(tmem_d, desc64(a_lo + 2 * k), desc64(w_lo + 2 * k), idesc, (kb | k) ? 1u : 0u);
          tc_commit(BAR(iAE + s));
        }
        __syncwarp();
        if (++s == rt.n_a_stages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) tc_commit(BAR(iCF + sc));
      __syncwarp();
      if (++sc == 2) { sc = 0; pc ^= 1; }
    }
  } else {
    // ===================== epilogue warps: one row per thread =====================
    const int q = warp & 3;       // TMEM lane quarter = rows q*32 .. q*32+31 of the tile
    const int half = warp >> 2;   // this warp takes the chunks at columns half*32 + 64*j
    const size_t ld = (size_t)p.ld;
    const uint32_t bias_s = smem_u32(s_bias);
    int sc = 0;
    uint32_t pc = 0;
    for (int tile = blockIdx.x; tile < rt.n_tiles; tile += gridDim.x) {
      const int row = tile * PW_ROWS + q * 32 + lane;
      const bool valid = row < rt.R;
      float m = 1.f;
      if (p.mask != nullptr) m = valid ? p.mask[row] : 0.f;
      if constexpr (MODE == EPI_POST) {
        // coupling update (modules.py:338-352): z[:, ch_off + n] = (z - sign * ((acc + bias) * mask)) * mask on the fp32 latent, plus
        // its operand copy.  N <= 128: at most two 32-channel chunks per warp, 128 bytes of fp32 z each.
        const size_t zoff = (size_t)row * ld + (size_t)p.ch_off + (size_t)(half * 32);
        uint32_t zr[2][32];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (valid && half * 32 + 64 * j < rt.N) {
            const char* zin = reinterpret_cast<const char*>(p.xin) + (zoff + (size_t)(64 * j)) * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) ldg256(zin + k * 32, zr[j] + 8 * k);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) zr[j][i] = 0u;
          }
        }
        mbar_wait(BAR(iCF + sc), pc);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sc * PW_ACC_STRIDE);
        const float sign = p.post_sign;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = half * 32 + 64 * j;
          if (c < rt.N) {
            float acc[32];
            tmem_ld32(taddr + (uint32_t)c, acc);
            tmem_ld_wait();
            uint32_t zo[32], u[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 b4;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(bias_s + (uint32_t)(c + i) * 4u));
              const float z0 = (__uint_as_float(zr[j][i]) - sign * ((acc[i] + b4.x) * m)) * m;
              const float z1 = (__uint_as_float(zr[j][i + 1]) - sign * ((acc[i + 1] + b4.y) * m)) * m;
              const float z2 = (__uint_as_float(zr[j][i + 2]) - sign * ((acc[i + 2] + b4.z) * m)) * m;
              const float z3 = (__uint_as_float(zr[j][i + 3]) - sign * ((acc[i + 3] + b4.w) * m)) * m;
              zo[i] = __float_as_uint(z0); zo[i + 1] = __float_as_uint(z1); zo[i + 2] = __float_as_uint(z2); zo[i + 3] = __float_as_uint(z3);
              u[i / 2] = pack_op2<Op>(z0, z1);
              u[i / 2 + 1] = pack_op2<Op>(z2, z3);
            }
            if (valid) {
              char* zout = reinterpret_cast<char*>(p.xout) + (zoff + (size_t)(64 * j)) * 4;
#pragma unroll
              for (int k = 0; k < 4; ++k) stg256(zout + k * 32, zo + 8 * k);
              char* ao = reinterpret_cast<char*>(p.act[0]) + (zoff + (size_t)(64 * j)) * 2;
              stg256(ao, u);
              stg256(ao + 32, u + 8);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(iCE + sc));
        if (++sc == 2) { sc = 0; pc ^= 1; }
        continue;
      }
      // the residual input does not depend on the accumulator: all of this thread's chunks are requested before the wait
      uint32_t res[PW_MAX_CHUNKS][16];
      if constexpr (MODE == EPI_RS) {
        const char* xin = reinterpret_cast<const char*>(p.xin) + ((size_t)row * ld + (size_t)(half * 32)) * 2;
#pragma unroll
        for (int j = 0; j < PW_MAX_CHUNKS; ++j) {
          if (valid && half * 32 + 64 * j < rt.N) {
            ldg256(xin + (size_t)j * 128, res[j]);
            ldg256(xin + (size_t)j * 128 + 32, res[j] + 8);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) res[j][i] = 0u;
          }
        }
      }
      mbar_wait(BAR(iCF + sc), pc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sc * PW_ACC_STRIDE);
#pragma unroll
      for (int j = 0; j < PW_MAX_CHUNKS; ++j) {
        const int c = half * 32 + 64 * j;
        if (c < rt.N) {  // warp-uniform
          float acc[32];
          tmem_ld32(taddr + (uint32_t)c, acc);
          tmem_ld_wait();
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(bias_s + (uint32_t)(c + i) * 4u));
            if constexpr (MODE == EPI_RS) {
              const __half2 h0 = *reinterpret_cast<const __half2*>(&res[j][i / 2]);
              const __half2 h1 = *reinterpret_cast<const __half2*>(&res[j][i / 2 + 1]);
              const float2 r0 = __half22float2(h0), r1 = __half22float2(h1);
              x[i] = (r0.x + acc[i] + b4.x) * m;
              x[i + 1] = (r0.y + acc[i + 1] + b4.y) * m;
              x[i + 2] = (r1.x + acc[i + 2] + b4.z) * m;
              x[i + 3] = (r1.y + acc[i + 3] + b4.w) * m;
            } else {
              x[i] = acc[i] + b4.x; x[i + 1] = acc[i + 1] + b4.y; x[i + 2] = acc[i + 2] + b4.z; x[i + 3] = acc[i + 3] + b4.w;
              if (p.mask != nullptr) { x[i] *= m; x[i + 1] *= m; x[i + 2] *= m; x[i + 3] *= m; }
            }
          }
          if (valid) {
            const size_t off = ((size_t)row * ld + (size_t)c) * 2;
            if constexpr (RH == 1) {  // fp16 stream copy
              uint32_t u[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) u[i] = pack_half2_sat(x[2 * i], x[2 * i + 1]);
              char* xo = reinterpret_cast<char*>(p.xout) + off;
              stg256(xo, u);
              stg256(xo + 32, u + 8);
            }
            if (MODE == EPI_RS || p.n_act > 0) {  // operand copy (leaky-relu'd for ACT; the WN stream is stored as is)
              uint32_t u[16];
              if constexpr (MODE == EPI_ACT) {
                const float slope = p.slope;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  u[i] = pack_op2<Op>(fmaxf(x[2 * i], x[2 * i] * slope), fmaxf(x[2 * i + 1], x[2 * i + 1] * slope));
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) u[i] = pack_op2<Op>(x[2 * i], x[2 * i + 1]);
              }
              char* ao = reinterpret_cast<char*>(p.act[0]) + off;
              stg256(ao, u);
              stg256(ao + 32, u + 8);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(iCE + sc));
      if (++sc == 2) { sc = 0; pc ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// WN gate conv (in_layers[l], k taps, 2H rows -> tanh(.) * sigmoid(.), commons.py:100-107) with TIME on the accumulator lane
// and cta_group::2 MMAs over CTA pairs  (gate_tm_kernel).
//
// On conv_tc_kernel the gate conv has three 128-row channel tiles (odd: no CTA pairs), single-CTA MMAs at ~175 cycles per
// N = 224 step (tensor pipe 49 %) and an epilogue in which the tanh and the sigmoid row of a channel live in different
// warps (shared-memory exchange).  Here
//   * A = the activation slab of ONE row tile per CTA (128 time steps of one utterance + the taps' halo, all k-blocks
//     resident while the tile's units run; a tap is a row offset of the A descriptor), B = the packed weight rows, split
//     between the two CTAs of a cluster: one tcgen05.mma.cta_group::2 (M = 256) multiplies BOTH CTAs' row tiles -- any two
//     row tiles, they need not be neighbours -- with the same weights, so every weight tile is fetched once per pair and
//     each CTA reads 4 KB of A + at most 4 KB of B per MMA.
//   * the packed row order [64 tanh | 64 sigmoid] per 64 channels (mbistft.cu pack_wn) puts both pre-activations of a
//     channel in the SAME thread (columns c and c + 64 of its row): no exchange.  Units of N = 256 (two packed tiles) while
//     two remain, then N = 128; consecutive units alternate between two 256-column TMEM slots, so the epilogue of one unit
//     overlaps the MMAs of the next.
//   * an epilogue thread owns one row: 32 gated channels = 64 contiguous bytes = two 256-bit stores.
// ------------------------------------------------------------------------------------------------
struct GtRt {
  int B, T, t_tiles, total_tiles, n_pair_tiles;
  int kblocks, taps, dil, shift0, N_total, n_ct;
  int slab_kb_bytes, slab_stage_bytes, n_w_stages;
  int w_off, gb_off, bar_off;
  int narrow_steps;   // (k-block, tap) steps of an N = 128 unit per weight-ring stage (1 or 2)
};
constexpr int GT_SLAB_STAGES = 2;
constexpr int GT_WARP_SLAB = PW_EPI_WARPS + 2;   // slab producer (warps 8 / 9: weight producer / MMA issuer)
constexpr int GT_THREADS = PW_THREADS + 32;
constexpr int GT_W_STAGE_BYTES = 128 * TC_ROW_BYTES;  // one CTA's half of an N = 256 weight tile

__device__ __forceinline__ float gt_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <typename Op>
__global__ void __launch_bounds__(GT_THREADS, 1)
gate_tm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWh,
               const EpiParams p, const GtRt rt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                  // GT_SLAB_STAGES x kblocks x slab_kb_bytes
  uint8_t* smW = smem + rt.w_off;       // n_w_stages x 16 KB
  float* s_gb = reinterpret_cast<float*>(smem + rt.gb_off);  // 2 x N_total: bias + cond_layer(g) of the row tile's utterance
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + rt.bar_off);
  const int iAF = 0, iAE = iAF + GT_SLAB_STAGES, iWF = iAE + GT_SLAB_STAGES, iWE = iWF + rt.n_w_stages, iCF = iWE + rt.n_w_stages,
            iCE = iCF + 2, nBars = iCE + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + nBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t crank = blockIdx.x & 1u;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // Schedule: the work is the sequence of (pair tile, packed 128-row weight tile) entries; every pair takes an equal
  // CONTIGUOUS share of it (55 168 rows = 224 pair tiles x 3 packed tiles = 672 entries = 9 or 10 per pair, where whole
  // pair tiles would be 3 or 4 per pair: 4 rounds for 3.03 rounds of work).  Consecutive entries of one pair tile run as one
  // N = 256 unit when two are left in the share, else as an N = 128 unit; a slab is loaded when the pair tile changes.
  const long long total_e = (long long)rt.n_pair_tiles * rt.n_ct;
  const int e_begin = (int)((long long)pair * total_e / n_pairs), e_end = (int)((long long)(pair + 1) * total_e / n_pairs);
  struct Unit { int pt, j, e_next; bool wide, first, last; };
  auto unit_at = [&](int e) {
    Unit u;
    u.pt = e / rt.n_ct;
    u.j = e - u.pt * rt.n_ct;
    u.wide = (u.j + 1 < rt.n_ct) && (e + 1 < e_end);
    u.e_next = e + (u.wide ? 2 : 1);
    u.first = (e == e_begin) || (u.j == 0);
    u.last = (u.e_next == e_end) || (u.e_next % rt.n_ct == 0);
    return u;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmWh);
    for (int i = 0; i < GT_SLAB_STAGES; ++i) { mbar_init(BAR(iAF + i), 1); mbar_init(BAR(iAE + i), 1); }
    for (int i = 0; i < rt.n_w_stages; ++i) { mbar_init(BAR(iWF + i), 1); mbar_init(BAR(iWE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(iCF + i), 1); mbar_init(BAR(iCE + i), 2 * PW_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == PW_WARP_MMA) tmem_alloc2(smem_u32(tmem_ptr_smem), 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp != PW_WARP_TMA) {   // (the weight producer does not wait: weights are constants of the model)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }

  if (warp == GT_WARP_SLAB) {
    // ===================== slab producer (both CTAs): the slab of the next pair tile is requested as soon as its stage is free,
    // a whole tile ahead of the weight producer's position
    int ss = 0;
    uint32_t ps = 0;
    for (int e = e_begin; e < e_end;) {
      const Unit u = unit_at(e);
      if (u.first) {
        const int ri = 2 * u.pt + (int)crank;
        const int b = ri < rt.total_tiles ? ri / rt.t_tiles : rt.B;  // no tile for this CTA: rows of utterance B do not exist -> zeros
        const int t0 = (ri % rt.t_tiles) * PW_ROWS + rt.shift0;
        mbar_wait(BAR(iAE + ss), ps ^ 1);
        if (elect_one()) {
          if (crank == 0) mbar_expect_tx(BAR(iAF + ss), (uint32_t)(2 * rt.slab_stage_bytes));
          const uint32_t af = mapa_shared(BAR(iAF + ss), 0);
          const uint32_t dst = smem_u32(smA) + (uint32_t)(ss * rt.slab_stage_bytes);
          for (int kb = 0; kb < rt.kblocks; ++kb) tma_load_3d_2sm(dst + (uint32_t)(kb * rt.slab_kb_bytes), &tmX, af, kb * 64, t0, b);
        }
        __syncwarp();
        if (++ss == GT_SLAB_STAGES) { ss = 0; ps ^= 1; }
      }
      e = u.e_next;
    }
  } else if (warp == PW_WARP_TMA) {
    // ===================== weight producer (both CTAs): own half of every weight tile =====================
    int sw = 0;
    uint32_t pw = 0;
    // one ring stage = one (k-block, tap) step of an N = 256 unit (16 KB per CTA) or TWO steps of an N = 128 unit (2 x 8 KB):
    // either way a stage lasts ~512 tensor-core cycles, so the ring covers the same TMA round trip
    const int n_steps = rt.kblocks * rt.taps;
    for (int e = e_begin; e < e_end;) {
      const Unit u = unit_at(e);
      const int half_rows = u.wide ? 128 : 64;   // N = 256: packed tiles j (even CTA) and j+1 (odd CTA); N = 128: halves of tile j
      const int row0 = 128 * u.j + (int)crank * half_rows;
      const int per_stage = u.wide ? 1 : rt.narrow_steps;
      for (int i = 0; i < n_steps; i += per_stage) {
        const int n_here = min(per_stage, n_steps - i);
        mbar_wait(BAR(iWE + sw), pw ^ 1);
        if (elect_one()) {
          if (crank == 0) mbar_expect_tx(BAR(iWF + sw), (uint32_t)(2 * n_here * half_rows * TC_ROW_BYTES));
          const uint32_t wf = mapa_shared(BAR(iWF + sw), 0);
          const uint32_t wdst = smem_u32(smW) + (uint32_t)(sw * GT_W_STAGE_BYTES);
          for (int d = 0; d < n_here; ++d) {
            const int kb = (i + d) / rt.taps, tap = (i + d) - kb * rt.taps;
            tma_load_2d_2sm(wdst + (uint32_t)(d * half_rows * TC_ROW_BYTES), u.wide ? &tmW : &tmWh, wf, kb * 64, tap * rt.N_total + row0);
          }
        }
        __syncwarp();
        if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
      }
      e = u.e_next;
    }
  } else if (warp == PW_WARP_MMA) {
    // ===================== MMA issuer (even CTA of the pair) =====================
    if (crank == 0) {
      constexpr uint32_t fmt = (Op::kPrec == 3) ? 0u : 1u;
      const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)((2 * PW_ROWS) >> 4) << 24);
      const uint32_t tap_step = (uint32_t)(rt.dil * TC_ROW_BYTES) >> 4;
      int ss = 0, sw = 0;
      uint32_t ps = 0, pw = 0, useq = 0;
      for (int e = e_begin; e < e_end;) {
        const Unit u = unit_at(e);
        if (u.first) {
          mbar_wait(BAR(iAF + ss), ps);
          tc_fence_after();
        }
        const uint32_t idesc = idesc0 | ((uint32_t)((u.wide ? 256 : 128) >> 3) << 17);
        const uint32_t slot = useq & 1u;
        mbar_wait(BAR(iCE + slot), ((useq >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + slot * (uint32_t)PW_ACC_STRIDE;
        uint32_t accum = 0;
        const int n_steps = rt.kblocks * rt.taps, per_stage = u.wide ? 1 : rt.narrow_steps;
        const uint32_t half_bytes = (uint32_t)((u.wide ? 128 : 64) * TC_ROW_BYTES);
        for (int i = 0; i < n_steps; i += per_stage) {
          const int n_here = min(per_stage, n_steps - i);
          mbar_wait(BAR(iWF + sw), pw);
          tc_fence_after();
          const uint32_t w_stage = smem_u32(smW) + (uint32_t)(sw * GT_W_STAGE_BYTES);
          for (int d = 0; d < n_here; ++d) {
            const int kb = (i + d) / rt.taps, tap = (i + d) - kb * rt.taps;
            const uint32_t a_lo = desc_lo(smem_u32(smA) + (uint32_t)(ss * rt.slab_stage_bytes + kb * rt.slab_kb_bytes)) + (uint32_t)tap * tap_step;
            const uint32_t w_lo = desc_lo(w_stage + (uint32_t)d * half_bytes);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma2<2>(tmem_d, desc64(a_lo + 2 * k), desc64(w_lo + 2 * k), idesc, (k == 0) ? accum : 1u);
              if (d == n_here - 1) tc_commit2_mc(BAR(iWE + sw), (uint16_t)3);
            }
            __syncwarp();
            accum = 1;
          }
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
        if (elect_one()) {
          tc_commit2_mc(BAR(iCF + slot), (uint16_t)3);
          if (u.last) tc_commit2_mc(BAR(iAE + ss), (uint16_t)3);  // both CTAs' slabs are free once the tile's last unit is done
        }
        __syncwarp();
        ++useq;
        if (u.last && ++ss == GT_SLAB_STAGES) { ss = 0; ps ^= 1; }
        e = u.e_next;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs): one row per thread, tanh and sigmoid columns of a channel side by side
    using T = typename Op::T;
    const int q = warp & 3, half = warp >> 2;
    const int tid_epi = threadIdx.x;  // epilogue warps are warps 0..7
    uint32_t useq = 0;
    int it = 0;
    bool valid = false;
    uint32_t gb_s = 0;
    char* out_row = nullptr;
    for (int e = e_begin; e < e_end;) {
      const Unit u = unit_at(e);
      if (u.first) {  // a new pair tile: this thread's row, bias + cond_layer(g) of its utterance
        const int ri = 2 * u.pt + (int)crank;
        const bool tile_ok = ri < rt.total_tiles;
        const int b = tile_ok ? ri / rt.t_tiles : 0;
        const int t = (ri % rt.t_tiles) * PW_ROWS + q * 32 + lane;
        valid = tile_ok && t < rt.T;
        float* gb = s_gb + (it & 1) * rt.N_total;
        ++it;
        for (int i = tid_epi; i < rt.N_total; i += 32 * PW_EPI_WARPS) {
          float v = p.bias[(size_t)b * p.bias_bs + i];
          if (p.add2 != nullptr) v += p.add2[(size_t)b * p.add2_bs + i];
          gb[i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * PW_EPI_WARPS) : "memory");
        gb_s = smem_u32(gb);
        out_row = reinterpret_cast<char*>(p.act[0]) + (((size_t)b * p.rows_out + (size_t)t) * p.ld + (size_t)p.ch_off) * sizeof(T);
      }
      const uint32_t slot = useq & 1u;
      mbar_wait(BAR(iCF + slot), (useq >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)PW_ACC_STRIDE;
      const int ptile = u.wide ? u.j + half : u.j;  // packed 128-row tile this warp drains
      const int col0 = u.wide ? 128 * half : 0;
      const int nsub = u.wide ? 2 : 1;
      for (int sub = 0; sub < nsub; ++sub) {
        const int cb = u.wide ? 32 * sub : 32 * half;   // first of the 32 channels (within the packed tile's 64)
        float ta[32], sg[32];
        tmem_ld32(taddr + (uint32_t)(col0 + cb), ta);
        tmem_ld32(taddr + (uint32_t)(col0 + 64 + cb), sg);
        tmem_ld_wait();
        const uint32_t gsrc = gb_s + (uint32_t)(128 * ptile + cb) * 4u;
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 bt, bs;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bt.x), "=f"(bt.y), "=f"(bt.z), "=f"(bt.w) : "r"(gsrc + (uint32_t)i * 4u));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bs.x), "=f"(bs.y), "=f"(bs.z), "=f"(bs.w) : "r"(gsrc + (uint32_t)(64 + i) * 4u));
          const float g0 = gt_tanh(ta[i] + bt.x) * fmaf(gt_tanh(0.5f * (sg[i] + bs.x)), 0.5f, 0.5f);
          const float g1 = gt_tanh(ta[i + 1] + bt.y) * fmaf(gt_tanh(0.5f * (sg[i + 1] + bs.y)), 0.5f, 0.5f);
          const float g2 = gt_tanh(ta[i + 2] + bt.z) * fmaf(gt_tanh(0.5f * (sg[i + 2] + bs.z)), 0.5f, 0.5f);
          const float g3 = gt_tanh(ta[i + 3] + bt.w) * fmaf(gt_tanh(0.5f * (sg[i + 3] + bs.w)), 0.5f, 0.5f);
          o[i / 2] = pack_op2<Op>(g0, g1);
          o[i / 2 + 1] = pack_op2<Op>(g2, g3);
        }
        if (valid) {
          char* dst = out_row + (size_t)(64 * ptile + cb) * sizeof(T);
          stg256(dst, o);
          stg256(dst + 32, o + 8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (crank != 0) mbar_arrive_remote(BAR(iCE + slot), 0u);  // the accumulator-free barriers live in the even CTA
        else mbar_arrive(BAR(iCE + slot));
      }
      ++useq;
      e = u.e_next;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == PW_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused ResBlock1 conv pair with TIME on the accumulator lane  (pair_tm_kernel):  x' = x + c2(lrelu(c1(a) + b1)) + b2
// for the k = 3 pairs of a 128-channel stage (modules.py:217-224), the HBM-bound third of that stage: two launches move
// 1356 MB per pair (the intermediate h = lrelu(c1(a)) goes out and comes back), one kernel 904 MB.
//
// Two earlier designs with the output CHANNEL on the lane (conv_pair_kernel) were slower than two launches: a
// thread owned one channel, so the h tile had to be written to shared memory with 2-byte scattered stores (5 K cycles per
// tile) and conv 1 -> h-tile epilogue -> conv 2 ran serially inside a CTA.  With time on the lane
//   * a thread owns one ROW of D1: its 128 channels are the two 128-byte rows (k-blocks) of the K-major, 128B-swizzled
//     operand tile conv 2 reads -- sixteen 16-byte shared-memory stores per row, conflict-free;
//   * the epilogue of conv 2 is the row-per-thread residual add of pw_tc_kernel: 256-bit loads of x (requested before the
//     accumulator wait) and 256-bit stores of x' (fp16 stream) and lrelu(x') (operand copy);
//   * cta_group::2: each CTA of a pair owns its own row tile (slab, h tile, accumulators) and HALF of every weight tile;
//     all 12 weight tiles of the two k = 3 convs stay resident (96 KB per CTA) for the life of the persistent pair;
//   * two D1 and two D2 accumulator slots (4 x 128 TMEM columns): the issuer runs conv 1 of tile i+1 while the epilogue-1
//     warps turn tile i into its h tile, then conv 2 of tile i; epilogue 2 of tile i-1 streams to HBM meanwhile.
// A row tile yields 128 - (taps - 1) = 126 output rows: D1 covers rows [t0 - 1, t0 + 127), conv 2's outer taps need one row
// on either side, and the two rows that would need h rows past D1 are simply not stored (1.6 % of the MMA work).
// ------------------------------------------------------------------------------------------------
struct PtRt {
  int B, L, t_tiles, total_tiles, n_pair_tiles, out_rows;
  int taps, dil1, shift0;
  int slab_kb_bytes, slab_stage_bytes, h_kb_bytes;
  int slab_off, h_off, bias_off, bar_off;
  float slope_h;
  const float* bias_h;
  long long* dbg;   // MBV_TIMELINE=10: clock stamps of pair 0 (both CTAs) [cta][tile < 16][16] (debug only)
  int res_pf;       // 1: L2 prefetch of the next tile's residual rows (MBV_NO_RES_PF=1 clears it: A/B only)
};
constexpr int PT_KB = 2;                                  // 128 channels = two 64-channel k-blocks
constexpr int PT_W_TILE = 64 * TC_ROW_BYTES;              // one CTA's half (64 rows) of a weight tile
constexpr int PT_WARP_TMA = 16, PT_WARP_MMA = 17;         // warps 0-7 epilogue 1, 8-15 epilogue 2
constexpr int PT_THREADS = 32 * 18;

__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  // acquire at cluster scope: the arrivals come from both CTAs of the pair and order their shared-memory writes
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if ((unsigned long long)(clock64() - t0) > TC_TIMEOUT_CYCLES) {
      printf("mbistft pair_tm: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

template <typename Op>
__global__ void __launch_bounds__(PT_THREADS, 1)
pair_tm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
               const EpiParams p, const PtRt rt) {
  using T = typename Op::T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smW = smem;                     // [conv][k-block][tap] half tiles of 8 KB
  uint8_t* smA = smem + rt.slab_off;       // 2 stages x 2 k-blocks
  uint8_t* smH = smem + rt.h_off;          // 2 k-blocks x 136 rows x 128 B
  float* s_b1 = reinterpret_cast<float*>(smem + rt.bias_off);   // [128] c1 bias
  float* s_b2 = s_b1 + 128;                                     // [2][128] c2 bias (+ per-utterance cond) of the tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + rt.bar_off);
  const int iWF = 0, iAF = 1, iAE = 3, iD1F = 5, iD1E = 7, iHF = 9, iHE = 10, iD2F = 11, iD2E = 13, nBars = 15;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + nBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t crank = blockIdx.x & 1u;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int pt_begin = (int)((long long)pair * rt.n_pair_tiles / n_pairs), pt_end = (int)((long long)(pair + 1) * rt.n_pair_tiles / n_pairs);
  const int n_my = pt_end - pt_begin;
  const int n_w = 2 * PT_KB * rt.taps;     // resident weight half tiles per CTA
  auto STAMP = [&](int i, int slot) {
    if (rt.dbg != nullptr && pair == 0 && lane == 0 && i < 16) rt.dbg[((size_t)crank * 16 + i) * 16 + slot] = clock64();
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(BAR(iWF), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(iAF + i), 1); mbar_init(BAR(iAE + i), 1);
      mbar_init(BAR(iD1F + i), 1); mbar_init(BAR(iD1E + i), 16);
      mbar_init(BAR(iD2F + i), 1); mbar_init(BAR(iD2E + i), 16);
    }
    mbar_init(BAR(iHF), 16);
    mbar_init(BAR(iHE), 1);
    fence_barrier_init();
  }
  if (warp == PT_WARP_MMA) tmem_alloc2(smem_u32(tmem_ptr_smem), 512u);
  for (int i = threadIdx.x; i < 128; i += PT_THREADS) s_b1[i] = rt.bias_h[i];
  const bool shared_bias = p.bias_bs == 0;   // (per-utterance only for the first pair of a ResBlock with speaker conditioning)
  if (shared_bias) for (int i = threadIdx.x; i < 128; i += PT_THREADS) s_b2[i] = p.bias[i];
  // rows 128..135 of the h tile are never written (D1 has 128 rows); they only feed the two output rows nobody stores, but
  // they must not hold NaN patterns
  for (int i = threadIdx.x; i < PT_KB * 8 * 8; i += PT_THREADS) {
    const int kb = i / 64, r = 128 + (i % 64) / 8, ch = i % 8;
    *reinterpret_cast<uint4*>(smH + (size_t)kb * rt.h_kb_bytes + (size_t)r * TC_ROW_BYTES + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == PT_WARP_TMA) {  // resident weights: constants of the model, fetched before the dependency wait
    if (elect_one()) {
      if (crank == 0) mbar_expect_tx(BAR(iWF), (uint32_t)(2 * n_w * PT_W_TILE));
      const uint32_t wf = mapa_shared(BAR(iWF), 0);
      for (int c = 0; c < 2; ++c)
        for (int kb = 0; kb < PT_KB; ++kb)
          for (int tap = 0; tap < rt.taps; ++tap)
            tma_load_2d_2sm(smem_u32(smW) + (uint32_t)(((c * PT_KB + kb) * rt.taps + tap) * PT_W_TILE), c ? &tmW2 : &tmW1, wf, kb * 64,
                            tap * 128 + (int)crank * 64);
    }
    __syncwarp();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == PT_WARP_TMA) {
    // ===================== TMA producer (both CTAs): the slab of its own row tile =====================
    for (int i = 0; i < n_my; ++i) {
      const int ri = 2 * (pt_begin + i) + (int)crank;
      const int b = ri < rt.total_tiles ? ri / rt.t_tiles : rt.B;
      const int t0 = (ri % rt.t_tiles) * rt.out_rows;
      const int ss = i & 1;
      mbar_wait(BAR(iAE + ss), (((uint32_t)i >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        if (crank == 0) mbar_expect_tx(BAR(iAF + ss), (uint32_t)(2 * rt.slab_stage_bytes));
        const uint32_t af = mapa_shared(BAR(iAF + ss), 0);
        const uint32_t dst = smem_u32(smA) + (uint32_t)(ss * rt.slab_stage_bytes);
        for (int kb = 0; kb < PT_KB; ++kb) tma_load_3d_2sm(dst + (uint32_t)(kb * rt.slab_kb_bytes), &tmX, af, kb * 64, t0 - 1 + rt.shift0, b);
      }
      __syncwarp();
    }
  } else if (warp == PT_WARP_MMA) {
    // ===================== MMA issuer (even CTA): conv 1 of tile i+1, then conv 2 of tile i =====================
    if (crank == 0) {
      constexpr uint32_t fmt = (Op::kPrec == 3) ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)((2 * PW_ROWS) >> 4) << 24);
      const uint32_t tap_step1 = (uint32_t)(rt.dil1 * TC_ROW_BYTES) >> 4, tap_step2 = (uint32_t)TC_ROW_BYTES >> 4;
      mbar_wait(BAR(iWF), 0);
      tc_fence_after();
      auto conv = [&](int c, uint32_t a_base, uint32_t a_kb_bytes, uint32_t tap_step, uint32_t tmem_d) {
        uint32_t accum = 0;
        for (int kb = 0; kb < PT_KB; ++kb) {
          uint32_t a_lo = desc_lo(a_base + (uint32_t)kb * a_kb_bytes);
          for (int tap = 0; tap < rt.taps; ++tap) {
            const uint32_t w_lo = desc_lo(smem_u32(smW) + (uint32_t)(((c * PT_KB + kb) * rt.taps + tap) * PT_W_TILE));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma2<2>(tmem_d, desc64(a_lo + 2 * k), desc64(w_lo + 2 * k), idesc, (k == 0) ? accum : 1u);
            }
            __syncwarp();
            accum = 1;
            a_lo += tap_step;
          }
        }
      };
      auto conv1 = [&](int i) {
        const int s = i & 1;
        const uint32_t par = ((uint32_t)i >> 1) & 1u;
        STAMP(i, 0);
        mbar_wait(BAR(iAF + s), par);
        STAMP(i, 1);
        mbar_wait(BAR(iD1E + s), par ^ 1u);
        tc_fence_after();
        STAMP(i, 2);
        conv(0, smem_u32(smA) + (uint32_t)(s * rt.slab_stage_bytes), (uint32_t)rt.slab_kb_bytes, tap_step1, tmem_base + (uint32_t)(s * 128));
        if (elect_one()) { tc_commit2_mc(BAR(iAE + s), (uint16_t)3); tc_commit2_mc(BAR(iD1F + s), (uint16_t)3); }
        __syncwarp();
      };
      auto conv2 = [&](int i) {
        const int s = i & 1;
        const uint32_t par = ((uint32_t)i >> 1) & 1u;
        STAMP(i, 3);
        mbar_wait_cluster(BAR(iHF), (uint32_t)i & 1u);   // both CTAs' h tiles are written
        STAMP(i, 4);
        mbar_wait(BAR(iD2E + s), par ^ 1u);
        tc_fence_after();
        STAMP(i, 5);
        conv(1, smem_u32(smH), (uint32_t)rt.h_kb_bytes, tap_step2, tmem_base + 256u + (uint32_t)(s * 128));
        if (elect_one()) { tc_commit2_mc(BAR(iHE), (uint16_t)3); tc_commit2_mc(BAR(iD2F + s), (uint16_t)3); }
        __syncwarp();
      };
      if (n_my > 0) conv1(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) conv1(i + 1);
        conv2(i);
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue-1 warps (8): D1 -> h tile, one row per thread, one k-block (two 32-channel chunks) per warp
    const int q = warp & 3, r = q * 32 + lane, kbw = warp >> 2;
    const float slope_h = rt.slope_h;
    const uint32_t b1_s = smem_u32(s_b1);
    const uint32_t h_row = smem_u32(smH) + (uint32_t)r * TC_ROW_BYTES;
    for (int i = 0; i < n_my; ++i) {
      const int ri = 2 * (pt_begin + i) + (int)crank;
      const int t_row = (ri % rt.t_tiles) * rt.out_rows - 1 + r;
      const bool inside = ri < rt.total_tiles && t_row >= 0 && t_row < rt.L;   // outside the utterance h is conv 2's zero padding
      const int s = i & 1;
      mbar_wait(BAR(iD1F + s), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      if (warp == 0) STAMP(i, 6);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128);
      // Two halves of two 32-channel chunks (= one k-block of the h tile each), so that only 32 packed registers are live: the first
      // half is converted while conv 2 of the previous tile still reads the h tile and stored as soon as that conv is done; the
      // second half follows.  (Holding all four chunks spilled and put local-memory loads in front of every shared-memory store.)
      auto convert2 = [&](int c0, uint32_t (*hp)[16]) {
        float acc0[32], acc1[32];
        tmem_ld32(taddr + (uint32_t)(32 * c0), acc0);
        tmem_ld32(taddr + (uint32_t)(32 * c0 + 32), acc1);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float* acc = h ? acc1 : acc0;
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(b1_s + (uint32_t)(32 * (c0 + h) + k) * 4u));
            float v0 = acc[k] + b4.x, v1 = acc[k + 1] + b4.y, v2 = acc[k + 2] + b4.z, v3 = acc[k + 3] + b4.w;
            v0 = fmaxf(v0, v0 * slope_h); v1 = fmaxf(v1, v1 * slope_h); v2 = fmaxf(v2, v2 * slope_h); v3 = fmaxf(v3, v3 * slope_h);
            hp[h][k / 2] = pack_op2<Op>(v0, v1);
            hp[h][k / 2 + 1] = pack_op2<Op>(v2, v3);
          }
        }
        if (!inside) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 16; ++k) hp[h][k] = 0u;
        }
      };
      auto store2 = [&](int kb, uint32_t (*hp)[16]) {   // chunks 2*kb, 2*kb+1 = the 128-byte row of k-block kb
        const uint32_t base = h_row + (uint32_t)kb * (uint32_t)rt.h_kb_bytes;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t ci = (uint32_t)(h * 4 + j);
            const uint32_t addr = base + ((ci ^ (uint32_t)(r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(hp[h][4 * j]), "r"(hp[h][4 * j + 1]), "r"(hp[h][4 * j + 2]),
                         "r"(hp[h][4 * j + 3]) : "memory");
          }
      };
      uint32_t hp[2][16];
      convert2(2 * kbw, hp);          // converted while conv 2 of the previous tile still reads the h tile
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // this warp's half of the D1 slot is drained
        if (crank != 0) mbar_arrive_remote(BAR(iD1E + s), 0u);
        else mbar_arrive(BAR(iD1E + s));
      }
      if (warp == 0) STAMP(i, 7);
      mbar_wait(BAR(iHE), ((uint32_t)i & 1u) ^ 1u);   // conv 2 of the previous tile has finished reading the h tile
      if (warp == 0) STAMP(i, 8);
      store2(kbw, hp);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) {
        if (crank != 0) mbar_arrive_remote(BAR(iHF), 0u);   // (relaxed: the st.shared + fence.proxy.async above have completed in this SM before the arrive issues)
        else mbar_arrive(BAR(iHF));
      }
      if (warp == 0) STAMP(i, 9);
    }
  } else {
    // ===================== epilogue-2 warps (8): the residual add on D2, one row per thread, two 32-channel chunks per warp
    const int q = warp & 3, r = q * 32 + lane, half = (warp - 8) >> 2;
    const int tid2 = threadIdx.x - 256;
    const float slope = p.slope;
    for (int i = 0; i < n_my; ++i) {
      const int ri = 2 * (pt_begin + i) + (int)crank;
      const bool tile_ok = ri < rt.total_tiles;
      const int b = tile_ok ? ri / rt.t_tiles : 0;
      const int t = (ri % rt.t_tiles) * rt.out_rows + r;
      const bool valid = tile_ok && r < rt.out_rows && t < rt.L;
      const size_t row_off = ((size_t)b * rt.L + (size_t)t) * (size_t)p.ld * 2 + (size_t)half * 128;
      uint32_t res[32];
      if (valid) {
        const char* xin = reinterpret_cast<const char*>(p.xin) + row_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) ldg256(xin + j * 32, res + 8 * j);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) res[j] = 0u;
      }
      {  // the next tile's residual row: asked into L2 a tile ahead (more bytes in flight: 191 -> 179 us per pair)
        const int rn = ri + 2;
        if (rt.res_pf && i + 1 < n_my && rn < rt.total_tiles && r < rt.out_rows) {
          const int tn = (rn % rt.t_tiles) * rt.out_rows + r;
          if (tn < rt.L)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.xin) +
                                                          ((size_t)(rn / rt.t_tiles) * rt.L + (size_t)tn) * (size_t)p.ld * 2 + (size_t)half * 128));
        }
      }
      const float* b2 = s_b2;
      if (!shared_bias) {
        float* b2w = s_b2 + (i & 1) * 128;
        if (tid2 < 128) b2w[tid2] = p.bias[(size_t)b * p.bias_bs + tid2];
        asm volatile("bar.sync 2, 256;" ::: "memory");
        b2 = b2w;
      }
      const uint32_t b2_s = smem_u32(b2) + (uint32_t)half * 256u;
      const int s = i & 1;
      if (warp == 8) STAMP(i, 10);
      mbar_wait(BAR(iD2F + s), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      if (warp == 8) STAMP(i, 11);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(s * 128) + (uint32_t)half * 64u;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float acc[32];
        tmem_ld32(taddr + (uint32_t)(32 * c), acc);
        tmem_ld_wait();
        if (c == 1) {  // this warp's half of the D2 slot is drained
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (crank != 0) mbar_arrive_remote(BAR(iD2E + s), 0u);
            else mbar_arrive(BAR(iD2E + s));
          }
        }
        uint32_t xo[16], ao[16];
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          float4 b4;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(b2_s + (uint32_t)(32 * c + k) * 4u));
          const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&res[16 * c + k / 2]));
          const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&res[16 * c + k / 2 + 1]));
          const float x0 = r0.x + acc[k] + b4.x, x1 = r0.y + acc[k + 1] + b4.y, x2 = r1.x + acc[k + 2] + b4.z, x3 = r1.y + acc[k + 3] + b4.w;
          xo[k / 2] = pack_half2_sat(x0, x1);
          xo[k / 2 + 1] = pack_half2_sat(x2, x3);
          ao[k / 2] = pack_op2<Op>(fmaxf(x0, x0 * slope), fmaxf(x1, x1 * slope));
          ao[k / 2 + 1] = pack_op2<Op>(fmaxf(x2, x2 * slope), fmaxf(x3, x3 * slope));
        }
        if (valid) {
          char* xout = reinterpret_cast<char*>(p.sum_mode == 1 ? p.xs : p.xout) + row_off + c * 64;
          stg256(xout, xo);
          stg256(xout + 32, xo + 8);
          if (p.n_act > 0) {
            char* aout = reinterpret_cast<char*>(p.act[0]) + row_off + c * 64;
            stg256(aout, ao);
            stg256(aout + 32, ao + 8);
          }
        }
      }
      if (warp == 8) STAMP(i, 12);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == PT_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// Multi-tap convs of a 128-channel ResBlock stage with TIME on the accumulator lane  (conv_tm_kernel): the k = 7 / k = 11
// convs of ResBlock1 (c1: bias -> lrelu -> operand copy; c2: the residual add in its four running-sum flavours,
// modules.py:213-228, models.py:355-361).
//
// On conv_tc_kernel these convs have ONE 128-row channel tile, so they run as single-CTA MMAs (M 128, N 256: 12 KB of shared
// memory operands per MMA, ~250 cycles against 128 nominal) -- the largest block of the step.  With time on the lane the
// operand roles swap (A = activation rows, B = the 128 weight rows of a tap) and a CTA PAIR shares B: one
// tcgen05.mma.cta_group::2 (M = 256, N = 128) multiplies one 128-row sub-tile of EACH CTA with the same weight tile, half of
// whose rows each CTA holds: 4 KB of A + 2 KB of B per CTA and MMA, measured ~88 cycles per MMA in pair_tm_kernel (N = 128,
// i.e. ~176 per N = 256 equivalent).
//   * a CTA tile is 256 consecutive time steps of one utterance = two 128-row sub-tiles, so every weight stage (two
//     (k-block, tap) steps = 16 KB per CTA) feeds 16 MMAs and the L2 -> shared-memory weight traffic per output row is that of
//     the 256-column tiles of conv_tc_kernel; the slab (256 + halo rows, both k-blocks, two TMA boxes per k-block) is loaded
//     once per tile, a tap is a row offset of the A descriptor; the next tile's slab is requested in the middle of the current
//     tile's weight stages;
//   * accumulators: 2 slots x 2 sub-tiles x 128 TMEM columns; 16 epilogue warps (8 per sub-tile: lane quarter x 64-channel half);
//   * an epilogue thread owns one ROW: 64 channels = 128 contiguous bytes of the channels-last tensors per stream: 256-bit
//     loads of the residual row (requested before the accumulator wait) and 256-bit stores, as in pair_tm_kernel.
// ------------------------------------------------------------------------------------------------
struct CtRt {
  int B, L, t_tiles, total_tiles, n_pair_tiles;
  int taps, dil, shift0, box_rows;
  int slab_kb_bytes, slab_stage_bytes, n_w_stages;
  int w_off, bias_off, bar_off;
  int res_pf;     // 1: L2 prefetch of the next tile's residual rows (MBV_NO_RES_PF=1 clears it: A/B only)
};
constexpr int CT_ROWS = 256;                               // rows per CTA tile: two 128-row accumulators
constexpr int CT_KB = 2;                                   // 128 input channels = two 64-channel k-blocks
constexpr int CT_W_STEP = 64 * TC_ROW_BYTES;               // one CTA's half (64 rows) of a (k-block, tap) weight tile
constexpr int CT_W_STAGE_BYTES = 2 * CT_W_STEP;            // two steps per ring stage
constexpr int CT_SLAB_STAGES = 2;
constexpr int CT_EPI_WARPS = 16;
constexpr int CT_WARP_TMA = 16, CT_WARP_MMA = 17;
constexpr int CT_THREADS = 32 * 18;

// SM = the residual add's running-sum flavour (EpiParams::sum_mode 0..3), compile-time so that each variant holds only its own streams in registers
template <typename Op, int MODE, int SM>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const EpiParams p, const CtRt rt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                  // CT_SLAB_STAGES x 2 k-blocks x 2 boxes
  uint8_t* smW = smem + rt.w_off;       // n_w_stages x 16 KB
  float* s_bias = reinterpret_cast<float*>(smem + rt.bias_off);   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + rt.bar_off);
  const int iAF = 0, iAE = iAF + CT_SLAB_STAGES, iWF = iAE + CT_SLAB_STAGES, iWE = iWF + rt.n_w_stages, iCF = iWE + rt.n_w_stages,
            iCE = iCF + 2, nBars = iCE + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + nBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t crank = blockIdx.x & 1u;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int pt_begin = (int)((long long)pair * rt.n_pair_tiles / n_pairs), pt_end = (int)((long long)(pair + 1) * rt.n_pair_tiles / n_pairs);
  const int n_my = pt_end - pt_begin;
  const int n_stages_tile = rt.taps;    // CT_KB * taps steps, two per stage

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < CT_SLAB_STAGES; ++i) { mbar_init(BAR(iAF + i), 1); mbar_init(BAR(iAE + i), 1); }
    for (int i = 0; i < rt.n_w_stages; ++i) { mbar_init(BAR(iWF + i), 1); mbar_init(BAR(iWE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(iCF + i), 1); mbar_init(BAR(iCE + i), 2 * CT_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == CT_WARP_MMA) tmem_alloc2(smem_u32(tmem_ptr_smem), 512u);
  const bool shared_bias = p.bias_bs == 0;
  if (shared_bias) for (int i = threadIdx.x; i < 128; i += CT_THREADS) s_bias[i] = p.bias[i];   // weights: not produced by the previous kernel
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp != CT_WARP_TMA) {   // (the producer requests the first weight stages first, see below)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }

  if (warp == CT_WARP_TMA) {
    // ===================== TMA producer (both CTAs): own slab, own half of every weight tile =====================
    auto load_slab = [&](int i) {
      const int ri = 2 * (pt_begin + i) + (int)crank;
      const int b = ri < rt.total_tiles ? ri / rt.t_tiles : rt.B;   // no tile for this CTA: rows of utterance B do not exist -> zeros
      const int t0 = (ri % rt.t_tiles) * CT_ROWS + rt.shift0;
      const int ss = i & 1;
      mbar_wait(BAR(iAE + ss), (((uint32_t)i >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        if (crank == 0) mbar_expect_tx(BAR(iAF + ss), (uint32_t)(2 * rt.slab_stage_bytes));
        const uint32_t af = mapa_shared(BAR(iAF + ss), 0);
        const uint32_t dst = smem_u32(smA) + (uint32_t)(ss * rt.slab_stage_bytes);
        for (int kb = 0; kb < CT_KB; ++kb)
          for (int bx = 0; bx < 2; ++bx)
            tma_load_3d_2sm(dst + (uint32_t)(kb * rt.slab_kb_bytes + bx * rt.box_rows * TC_ROW_BYTES), &tmX, af, kb * 64, t0 + bx * rt.box_rows, b);
      }
      __syncwarp();
    };
    int sw = 0;
    uint32_t pw = 0;
    auto load_w = [&](int g) {   // weight stage g of this CTA's flat (tile, stage) sequence
      const int st = g % n_stages_tile;
      mbar_wait(BAR(iWE + sw), pw ^ 1);
      if (elect_one()) {
        if (crank == 0) mbar_expect_tx(BAR(iWF + sw), (uint32_t)(2 * CT_W_STAGE_BYTES));
        const uint32_t wf = mapa_shared(BAR(iWF + sw), 0);
        const uint32_t wdst = smem_u32(smW) + (uint32_t)(sw * CT_W_STAGE_BYTES);
        for (int d = 0; d < 2; ++d) {
          const int step = 2 * st + d, kb = step / rt.taps, tap = step - kb * rt.taps;
          tma_load_2d_2sm(wdst + (uint32_t)(d * CT_W_STEP), &tmW, wf, kb * 64, tap * 128 + (int)crank * 64);
        }
      }
      __syncwarp();
      if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
    };
    // weights are constants of the model: the first turn of the ring is requested BEFORE the programmatic-dependency wait
    const int total_g = n_my * n_stages_tile;
    const int pf = rt.n_w_stages < total_g ? rt.n_w_stages : total_g;
    int g = 0;
    for (; g < pf; ++g) load_w(g);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // the next tile's slab is requested once the ring has turned over inside the current tile (its stage is free by then)
    const int pre = rt.n_w_stages < n_stages_tile - 1 ? rt.n_w_stages : n_stages_tile - 1;
    int next_slab = 0;
    if (n_my > 0) load_slab(next_slab++);
    for (; g < total_g; ++g) {
      while (next_slab < n_my && g >= (next_slab - 1) * n_stages_tile + pre) load_slab(next_slab++);
      load_w(g);
    }
    while (next_slab < n_my) load_slab(next_slab++);
  } else if (warp == CT_WARP_MMA) {
    // ===================== MMA issuer (even CTA of the pair) =====================
    if (crank == 0) {
      constexpr uint32_t fmt = (Op::kPrec == 3) ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)((2 * PW_ROWS) >> 4) << 24);
      const uint32_t tap_step = (uint32_t)(rt.dil * TC_ROW_BYTES) >> 4;
      constexpr uint32_t sub_step = (uint32_t)(PW_ROWS * TC_ROW_BYTES) >> 4;   // second sub-tile: 128 rows further down the slab
      int sw = 0;
      uint32_t pw = 0;
      for (int i = 0; i < n_my; ++i) {
        const int ss = i & 1;
        const uint32_t par = ((uint32_t)i >> 1) & 1u;
        mbar_wait(BAR(iAF + ss), par);
        mbar_wait(BAR(iCE + ss), par ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(ss * 256);
        const uint32_t a_stage = smem_u32(smA) + (uint32_t)(ss * rt.slab_stage_bytes);
        uint32_t accum = 0;
        uint32_t a_kb = desc_lo(a_stage), a_lo = a_kb;   // (k-block, tap) walked incrementally: no division on the issue path
        int tap = 0;
        for (int st = 0; st < n_stages_tile; ++st) {
          mbar_wait(BAR(iWF + sw), pw);
          tc_fence_after();
          const uint32_t w_stage = smem_u32(smW) + (uint32_t)(sw * CT_W_STAGE_BYTES);
          for (int d = 0; d < 2; ++d) {
            const uint32_t w_lo = desc_lo(w_stage + (uint32_t)(d * CT_W_STEP));
            if (elect_one()) {
#pragma unroll
              for (int sub = 0; sub < 2; ++sub)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc_mma2<2>(tmem_d + (uint32_t)(sub * 128), desc64(a_lo + (uint32_t)sub * sub_step + 2 * k), desc64(w_lo + 2 * k), idesc, (k == 0) ? accum : 1u);
              if (d == 1) tc_commit2_mc(BAR(iWE + sw), (uint16_t)3);
            }
            __syncwarp();
            accum = 1;
            a_lo += tap_step;
            if (++tap == rt.taps) { tap = 0; a_kb += (uint32_t)rt.slab_kb_bytes >> 4; a_lo = a_kb; }
          }
          if (++sw == rt.n_w_stages) { sw = 0; pw ^= 1; }
        }
        if (elect_one()) {
          tc_commit2_mc(BAR(iCF + ss), (uint16_t)3);
          tc_commit2_mc(BAR(iAE + ss), (uint16_t)3);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs): one row per thread, 64 channels (two 32-channel chunks) per warp
    const int sub = warp >> 3, q = warp & 3, half = (warp >> 2) & 1;
    const int r = sub * PW_ROWS + q * 32 + lane;
    const float slope = p.slope, scale = p.scale;
    constexpr int sum_mode = SM;
    for (int i = 0; i < n_my; ++i) {
      const int ri = 2 * (pt_begin + i) + (int)crank;
      const bool tile_ok = ri < rt.total_tiles;
      const int b = tile_ok ? ri / rt.t_tiles : 0;
      const int t = (ri % rt.t_tiles) * CT_ROWS + r;
      const bool valid = tile_ok && t < rt.L;
      const size_t res_off = ((size_t)b * p.rows_res + (size_t)t) * (size_t)p.ld * 2 + (size_t)half * 128;
      const int mrow = t + p.row_add;
      const size_t map_off = ((size_t)b * p.rows_out + (size_t)mrow) * (size_t)p.ld * 2 + (size_t)half * 128;
      uint32_t res[32];
      if constexpr (MODE == EPI_RES) {
        if (valid) {
          const char* xin = reinterpret_cast<const char*>(p.xin) + res_off;
#pragma unroll
          for (int j = 0; j < 4; ++j) ldg256(xin + j * 32, res + 8 * j);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) res[j] = 0u;
        }
      }
      if constexpr (MODE == EPI_RES && SM >= 2) {
        // The running ResBlock sum has no registers to wait in (res[] already takes 32): it is loaded after the accumulator wait,
        // chunk by chunk -- from L2, into which it is asked here so that those loads do not see a DRAM access on a saturated memory
        // system.  SM 2 (the k = 7 ResBlock's last conv, epilogue-bound): a tile AHEAD, together with the residual row (236 -> 193 us);
        // SM 3 (k = 11, MMA-bound) and the plain residual adds measured no gain from the tile-ahead requests.
        if ((SM == 3 || i == 0 || !rt.res_pf) && valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.xs) + res_off));
        const int rn = ri + 2;
        if (SM == 2 && rt.res_pf && i + 1 < n_my && rn < rt.total_tiles) {
          const int tn = (rn % rt.t_tiles) * CT_ROWS + r;
          if (tn < rt.L) {
            const size_t off_n = ((size_t)(rn / rt.t_tiles) * p.rows_res + (size_t)tn) * (size_t)p.ld * 2 + (size_t)half * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.xin) + off_n));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.xs) + off_n));
          }
        }
      }
      const float* bias = s_bias;
      if (!shared_bias) {
        float* bw = s_bias + (i & 1) * 128;
        if (threadIdx.x < 128) bw[threadIdx.x] = p.bias[(size_t)b * p.bias_bs + threadIdx.x];
        asm volatile("bar.sync 1, %0;" ::"n"(32 * CT_EPI_WARPS) : "memory");
        bias = bw;
      }
      const uint32_t b_s = smem_u32(bias) + (uint32_t)half * 256u;
      const int ss = i & 1;
      mbar_wait(BAR(iCF + ss), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ss * 256 + sub * 128 + half * 64);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t xsr[16];
        if constexpr (MODE == EPI_RES) {
          if (sum_mode >= 2) {
            if (valid) {
              const char* xs = reinterpret_cast<const char*>(p.xs) + res_off + c * 64;
              ldg256(xs, xsr);
              ldg256(xs + 32, xsr + 8);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) xsr[j] = 0u;
            }
          }
        }
        float acc[32];
        tmem_ld32(taddr + (uint32_t)(32 * c), acc);
        tmem_ld_wait();
        if (c == 1) {  // this warp's part of the accumulator slot is drained
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (crank != 0) mbar_arrive_remote(BAR(iCE + ss), 0u);   // the accumulator-free barriers live in the even CTA
            else mbar_arrive(BAR(iCE + ss));
          }
        }
        // 16 channels (one 256-bit store per stream) at a time: o1 = 2-byte stream output (x' / xs'), o2 = operand copy
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t o1[8], o2[8];
#pragma unroll
          for (int kk = 0; kk < 16; kk += 4) {
            const int k = 16 * g + kk;
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(b_s + (uint32_t)(32 * c + k) * 4u));
            float x0, x1, x2, x3;   // operation order of conv_tc.cu's epi_act / epi_res: results are bit-identical to the generic kernel
            if constexpr (MODE == EPI_RES) {
              const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&res[16 * c + k / 2]));
              const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&res[16 * c + k / 2 + 1]));
              x0 = r0.x; x1 = r0.y; x2 = r1.x; x3 = r1.y;
              if (sum_mode >= 2) {
                const float2 s0 = __half22float2(*reinterpret_cast<const __half2*>(&xsr[k / 2]));
                const float2 s1 = __half22float2(*reinterpret_cast<const __half2*>(&xsr[k / 2 + 1]));
                x0 += s0.x; x1 += s0.y; x2 += s1.x; x3 += s1.y;
              }
              x0 = x0 + acc[k] + b4.x; x1 = x1 + acc[k + 1] + b4.y; x2 = x2 + acc[k + 2] + b4.z; x3 = x3 + acc[k + 3] + b4.w;
              o1[kk / 2] = pack_half2_sat(x0, x1);
              o1[kk / 2 + 1] = pack_half2_sat(x2, x3);
              if (sum_mode == 3) { x0 *= scale; x1 *= scale; x2 *= scale; x3 *= scale; }   // mean over the parallel ResBlocks (models.py:361)
            } else {
              x0 = acc[k] + b4.x; x1 = acc[k + 1] + b4.y; x2 = acc[k + 2] + b4.z; x3 = acc[k + 3] + b4.w;
            }
            o2[kk / 2] = pack_op2<Op>(fmaxf(x0, x0 * slope), fmaxf(x1, x1 * slope));
            o2[kk / 2 + 1] = pack_op2<Op>(fmaxf(x2, x2 * slope), fmaxf(x3, x3 * slope));
          }
          if (valid) {
            const size_t co = (size_t)(c * 64 + g * 32);
            if constexpr (MODE == EPI_RES) {
              if (sum_mode <= 2)   // 0: x' -> residual stream;  1: x' starts the running ResBlock sum;  2: xs + x' -> xs
                stg256(reinterpret_cast<char*>(sum_mode == 0 ? p.xout : p.xs) + res_off + co, o1);
            }
            if (MODE == EPI_ACT || sum_mode == 0 || sum_mode == 3) {
              stg256(reinterpret_cast<char*>(p.act[0]) + map_off + co, o2);
              if (mrow == p.dup_src)   // ReflectionPad1d((1, 0)) in front of conv_post
                stg256(reinterpret_cast<char*>(p.act[0]) + ((size_t)b * p.rows_out + (size_t)p.dup_dst) * (size_t)p.ld * 2 + (size_t)half * 128 + co, o2);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == CT_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool pw_eligible(int prec, const ConvArgs& a, int flags) {
  const EpiParams& e = a.epi;
  if (flags & MBV_FLAG_NO_PW) return false;
  if (prec < 2 || a.taps != 1 || a.n_phases != 1 || a.shift0[0] != 0 || a.gate) return false;
  if (a.L_in != a.L_out || a.Cp_in % 64 != 0) return false;
  if (e.bias == nullptr || e.bias_bs != 0) return false;
  if (e.n_valid % 32 != 0 || e.n_valid < 32 || e.n_valid > 256 || e.n_valid > a.N_total || e.ld < e.n_valid || e.ld % 16 != 0) return false;
  if (e.row_mul != 1 || e.row_add != 0 || e.rows_out != a.L_out || e.rows_res != a.L_out || e.dup_src >= 0) return false;
  // resident weights + at least three 16 KB k-block stages
  if ((a.Cp_in / 64) * e.n_valid * TC_ROW_BYTES + 3 * PW_KB_BYTES + 4096 > 224 * 1024) return false;
  if (e.mode == EPI_RS) {
    if (e.res_half != 1 && e.res_half != 2) return false;
    if (e.res_half == 2 && prec != 3) return false;
    if (e.n_split < a.N_total || e.xin == nullptr || e.act[0] == nullptr || e.n_act != 1) return false;
    if (e.res_half == 1 && e.xout == nullptr) return false;
    if (e.res_half == 2 && e.inv_slope != 1.f) return false;
    return true;
  }
  if (e.mode == EPI_ACT) {
    if (e.n_act > 1 || (e.n_act == 1 && (e.act[0] == nullptr || e.act_add[0] != nullptr))) return false;
    if (e.xout != nullptr && e.res_half != 1) return false;
    if (e.xout == nullptr && e.n_act == 0) return false;
    return true;
  }
  if (e.mode == EPI_POST) {  // coupling update on the fp32 latent (K = n_layers * hidden: the weights still fit, N = half the latent)
    if (e.n_valid > 128 || e.xin == nullptr || e.xout == nullptr || e.act[0] == nullptr || e.n_act != 1 || e.mask == nullptr) return false;
    if (e.ch_off % 16 != 0 || e.ld % 16 != 0 || e.ch_off + e.n_valid > e.ld) return false;
    return true;
  }
  return false;
}

const char* pw_make_plan(int prec, const ConvArgs& a, int num_sms, TcPlan* plan) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tc_tensormap_encoder());
  if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
  const int kblocks = a.Cp_in / 64, N = a.epi.n_valid;
  const long long R = (long long)a.B * a.L_out;
  if (R > 0x7fffffff - PW_ROWS) return "pointwise conv: too many rows";
  plan->pw = 1;
  plan->pw_N = N; plan->pw_kblocks = kblocks; plan->pw_R = (int)R;
  plan->pw_tiles = (int)((R + PW_ROWS - 1) / PW_ROWS);
  plan->pw_w_bytes = kblocks * N * TC_ROW_BYTES;
  plan->pw_a_stage_bytes = PW_KB_BYTES;    // ring of k-block stages
  const int fixed = 2048;  // bias + barriers
  int stages = (224 * 1024 - plan->pw_w_bytes - fixed) / PW_KB_BYTES;
  if (stages > 12) stages = 12;
  if (stages < 3) return "pointwise conv: not enough shared memory";
  plan->pw_a_stages = stages;
  plan->pw_a_off = plan->pw_w_bytes;
  plan->pw_bias_off = plan->pw_a_off + stages * plan->pw_a_stage_bytes;
  plan->pw_bar_off = plan->pw_bias_off + 1024;
  plan->smem_bytes = 1024 + plan->pw_bar_off + (2 * stages + 5) * 8 + 16;
  if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns
  plan->grid = plan->pw_tiles < num_sms ? plan->pw_tiles : num_sms;
  if (plan->grid < 1) plan->grid = 1;
  plan->n_time = PW_ROWS;
  plan->cluster = 0;
  const CUtensorMapDataType dt = prec == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)a.x_ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)PW_ROWS};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&plan->tmA, dt, 2, const_cast<void*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the pointwise activation map";
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)N};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&plan->tmB, dt, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the pointwise weight map";
  }
  plan->tmR = plan->tmA; plan->tmS = plan->tmA; plan->tmBh = plan->tmB;  // unused
  if (a.epi.mode == EPI_RS || a.epi.mode == EPI_POST) {  // residual rows: L2 prefetch boxes of 128 bytes x 128 rows
    const bool f32 = a.epi.mode == EPI_POST;
    cuuint64_t dims[2] = {(cuuint64_t)a.epi.ld, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)a.epi.ld * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)PW_ROWS};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&plan->tmR, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a.epi.xin), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the pointwise residual map";
  }
  return nullptr;
}


// ---- pair_tm_kernel: eligibility, plan, launch
bool ptm_eligible(int prec, const ConvArgs& a, int flags, int num_sms) {
  const EpiParams& e = a.epi;
  if (flags & (MBV_FLAG_NO_PW | MBV_FLAG_NO_PAIR_TM)) return false;
  if (prec != 2 || num_sms < 2) return false;                       // bf16 operands + fp16 residual stream
  if (a.Cp_in != 128 || a.N_total != 128 || a.x_ld != 128 || a.n_phases != 1 || a.taps != 3 || a.L_in != a.L_out) return false;
  if (a.w2 == nullptr || a.bias_h == nullptr) return false;
  if (e.mode != EPI_RES || e.res_half != 1 || e.xin == nullptr || e.bias == nullptr) return false;
  // sum_mode 0: x' -> fp16 stream + operand copy;  1 (last pair of the first ResBlock): x' starts the running ResBlock sum, nothing else
  if (e.sum_mode == 0) { if (e.xout == nullptr || e.n_act != 1 || e.act[0] == nullptr || e.act_add[0] != nullptr) return false; }
  else if (e.sum_mode == 1) { if (e.xs == nullptr || e.n_act != 0 || e.xout != nullptr) return false; }
  else return false;
  if (e.ld != 128 || e.n_valid != 128 || e.row_mul != 1 || e.row_add != 0 || e.rows_out != a.L_out || e.rows_res != a.L_out || e.dup_src >= 0) return false;
  return PW_ROWS + 2 * a.dil <= 256;
}

const char* ptm_make_plan(int prec, const ConvArgs& a, int num_sms, TcPairPlan* plan) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tc_tensormap_encoder());
  if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
  const int halo1 = (a.taps - 1) * a.dil;
  const int box_rows = (PW_ROWS + halo1 + 7) / 8 * 8;
  plan->tm = 1;
  plan->tm_out_rows = PW_ROWS - (a.taps - 1);
  plan->t_tiles = (a.L_out + plan->tm_out_rows - 1) / plan->tm_out_rows;
  plan->total_tiles = a.B * plan->t_tiles;
  plan->box_rows = box_rows;
  plan->tm_slab_kb_bytes = box_rows * TC_ROW_BYTES;
  plan->tm_slab_stage_bytes = PT_KB * plan->tm_slab_kb_bytes;
  plan->tm_h_kb_bytes = (PW_ROWS + 8) * TC_ROW_BYTES;
  plan->tm_slab_off = 2 * PT_KB * a.taps * PT_W_TILE;
  plan->h_off = plan->tm_slab_off + 2 * plan->tm_slab_stage_bytes;
  plan->tm_bias_off = plan->h_off + PT_KB * plan->tm_h_kb_bytes;
  plan->bar_off = plan->tm_bias_off + 3 * 128 * 4;
  plan->smem_bytes = 1024 + plan->bar_off + 15 * 8 + 16;
  if (plan->smem_bytes > 227 * 1024) return "conv pair (time on lane): shared memory budget exceeded";
  const int n_pair_tiles = (plan->total_tiles + 1) / 2;
  const int pairs = n_pair_tiles < num_sms / 2 ? n_pair_tiles : num_sms / 2;
  plan->grid = 2 * (pairs < 1 ? 1 : pairs);
  const CUtensorMapDataType dt = prec == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.L_in, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.x_ld * 2, (cuuint64_t)a.L_in * a.x_ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&plan->tmA, dt, 3, const_cast<void*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the pair activation map";
  }
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.taps * a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    if (enc(which ? &plan->tmB2 : &plan->tmB, dt, 2, const_cast<void*>(which ? a.w2 : a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for a pair weight map";
  }
  return nullptr;
}

template <typename Op>
static cudaError_t ptm_launch_one(const ConvArgs& a, const TcPairPlan& p, const PtRt& rt, cudaStream_t st, int pdl, bool set_attr) {
  auto k = pair_tm_kernel<Op>;
  if (set_attr) return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(PT_THREADS);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, k, p.tmA, p.tmB, p.tmB2, a.epi, rt);
}

cudaError_t launch_ptm(int prec, const ConvArgs& a, const TcPairPlan& p, cudaStream_t st, int pdl) {
  PtRt rt;
  rt.B = a.B; rt.L = a.L_out; rt.t_tiles = p.t_tiles; rt.total_tiles = p.total_tiles; rt.n_pair_tiles = (p.total_tiles + 1) / 2;
  rt.out_rows = p.tm_out_rows; rt.taps = a.taps; rt.dil1 = a.dil; rt.shift0 = a.shift0[0];
  rt.slab_kb_bytes = p.tm_slab_kb_bytes; rt.slab_stage_bytes = p.tm_slab_stage_bytes; rt.h_kb_bytes = p.tm_h_kb_bytes;
  rt.slab_off = p.tm_slab_off; rt.h_off = p.h_off; rt.bias_off = p.tm_bias_off; rt.bar_off = p.bar_off;
  rt.slope_h = a.slope_h; rt.bias_h = a.bias_h;
  static const int no_res_pf = getenv("MBV_NO_RES_PF") ? atoi(getenv("MBV_NO_RES_PF")) : 0;  // A/B measurements only
  rt.res_pf = no_res_pf ? 0 : 1;
  if (prec != 2) return cudaErrorInvalidValue;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("MBV_TIMELINE");
    dbg_on = (e && atoi(e) == 10) ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, 2 * 16 * 16 * sizeof(long long)); cudaMemset(dbg, 0, 2 * 16 * 16 * sizeof(long long)); }
  }
  rt.dbg = dbg_on ? dbg : nullptr;
  cudaError_t e = ptm_launch_one<OpBF16>(a, p, rt, st, pdl, false);
  if (dbg_on && e == cudaSuccess) {
    cudaStreamSynchronize(st);
    long long hb[2 * 16 * 16];
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    const long long t0 = hb[0];
    fprintf(stderr, "[pair_tm timeline] d%d  columns: conv1 issue | slab ok | D1 free | conv2 issue | h ok | D2 free || epi1: D1 full | loaded | h free | h written || epi2: top | D2 full | stored\n", a.dil);
    for (int c = 0; c < 2; ++c)
      for (int i = 0; i < 12; ++i) {
        const long long* r = &hb[(c * 16 + i) * 16];
        fprintf(stderr, "  cta %d tile %2d  %7lld %7lld %7lld | %7lld %7lld %7lld || %7lld %7lld %7lld %7lld || %7lld %7lld %7lld\n", c, i,
                r[0] ? r[0] - t0 : 0, r[1] ? r[1] - t0 : 0, r[2] ? r[2] - t0 : 0, r[3] ? r[3] - t0 : 0, r[4] ? r[4] - t0 : 0, r[5] ? r[5] - t0 : 0,
                r[6] - t0, r[7] - t0, r[8] - t0, r[9] - t0, r[10] - t0, r[11] - t0, r[12] - t0);
      }
    cudaMemset(dbg, 0, sizeof(hb));
  }
  return e;
}

// ---- gate_tm_kernel: eligibility, plan, launch
bool gt_eligible(int prec, const ConvArgs& a, int flags, int num_sms) {
  const EpiParams& e = a.epi;
  if (flags & MBV_FLAG_NO_PW) return false;
  if (prec < 2 || e.mode != EPI_GATE || !a.gate || a.n_phases != 1 || num_sms < 2) return false;
  if (a.N_total % 128 != 0 || a.N_total > 1024 || a.Cp_in % 64 != 0 || a.L_in != a.L_out || a.x_ld < a.Cp_in) return false;
  if (e.bias == nullptr || e.act[0] == nullptr || e.n_act != 1 || e.row_mul != 1 || e.row_add != 0) return false;
  if (e.ld % 16 != 0 || e.ch_off % 16 != 0 || e.n_valid != a.N_total / 2) return false;
  const int halo = (a.taps - 1) * a.dil;
  if (PW_ROWS + halo > 256) return false;
  const int box_rows = (PW_ROWS + halo + 7) / 8 * 8;
  const int slab = GT_SLAB_STAGES * (a.Cp_in / 64) * box_rows * TC_ROW_BYTES;
  return slab + 4 * GT_W_STAGE_BYTES + 2 * a.N_total * 4 + 1024 <= 222 * 1024;
}

const char* gt_make_plan(int prec, const ConvArgs& a, int num_sms, TcPlan* plan) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tc_tensormap_encoder());
  if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
  const int kblocks = a.Cp_in / 64, halo = (a.taps - 1) * a.dil;
  const int box_rows = (PW_ROWS + halo + 7) / 8 * 8;
  plan->pw = 2;
  plan->pw_kblocks = kblocks;
  plan->pw_tiles = (a.L_out + PW_ROWS - 1) / PW_ROWS;                 // row tiles per utterance
  plan->pw_R = a.B * plan->pw_tiles;                                   // row tiles in total
  plan->pw_a_stage_bytes = kblocks * box_rows * TC_ROW_BYTES;
  plan->pw_w_bytes = box_rows * TC_ROW_BYTES;                          // one k-block of a slab
  plan->pw_a_off = GT_SLAB_STAGES * plan->pw_a_stage_bytes;            // weight ring starts here
  int stages = (222 * 1024 - plan->pw_a_off - 2 * a.N_total * 4 - 1024) / GT_W_STAGE_BYTES;
  if (stages > 8) stages = 8;
  if (stages < 3) return "gate conv: not enough shared memory for the weight ring";
  plan->pw_a_stages = stages;
  plan->pw_bias_off = plan->pw_a_off + stages * GT_W_STAGE_BYTES;
  plan->pw_bar_off = plan->pw_bias_off + (2 * a.N_total * 4 + 127) / 128 * 128;
  plan->smem_bytes = 1024 + plan->pw_bar_off + (2 * GT_SLAB_STAGES + 2 * stages + 4) * 8 + 16;
  if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;
  const int n_pair_tiles = (plan->pw_R + 1) / 2;
  const long long entries = (long long)n_pair_tiles * (a.N_total / 128);
  const int pairs = entries < num_sms / 2 ? (int)entries : num_sms / 2;
  plan->grid = 2 * (pairs < 1 ? 1 : pairs);
  plan->n_time = PW_ROWS;
  plan->cluster = 2;
  const CUtensorMapDataType dt = prec == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.L_in, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.x_ld * 2, (cuuint64_t)a.L_in * a.x_ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&plan->tmA, dt, 3, const_cast<void*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the gate activation map";
  }
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.taps * a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(which ? 64 : 128)};
    cuuint32_t estr[2] = {1, 1};
    if (enc(which ? &plan->tmBh : &plan->tmB, dt, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for a gate weight map";
  }
  plan->tmR = plan->tmA; plan->tmS = plan->tmA;  // unused
  return nullptr;
}

template <typename Op>
static cudaError_t gt_launch_one(const ConvArgs& a, const TcPlan& p, const GtRt& rt, cudaStream_t st, int pdl, bool set_attr) {
  auto k = gate_tm_kernel<Op>;
  if (set_attr) return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(GT_THREADS);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, k, p.tmA, p.tmB, p.tmBh, a.epi, rt);
}

static cudaError_t launch_gt(int prec, const ConvArgs& a, const TcPlan& p, cudaStream_t st, int pdl) {
  GtRt rt;
  rt.B = a.B; rt.T = a.L_out; rt.t_tiles = p.pw_tiles; rt.total_tiles = p.pw_R; rt.n_pair_tiles = (p.pw_R + 1) / 2;
  rt.kblocks = p.pw_kblocks; rt.taps = a.taps; rt.dil = a.dil; rt.shift0 = a.shift0[0]; rt.N_total = a.N_total; rt.n_ct = a.N_total / 128;
  rt.slab_kb_bytes = p.pw_w_bytes; rt.slab_stage_bytes = p.pw_a_stage_bytes; rt.n_w_stages = p.pw_a_stages;
  rt.w_off = p.pw_a_off; rt.gb_off = p.pw_bias_off; rt.bar_off = p.pw_bar_off;
  static const int narrow_env = getenv("MBV_GT_NARROW") ? atoi(getenv("MBV_GT_NARROW")) : 2;  // A/B measurements only
  rt.narrow_steps = narrow_env == 1 ? 1 : 2;
  if (prec == 3) return gt_launch_one<OpF16>(a, p, rt, st, pdl, false);
  return gt_launch_one<OpBF16>(a, p, rt, st, pdl, false);
}

// ---- conv_tm_kernel: eligibility, plan, launch (plan->pw = 3)
bool ct_eligible(int prec, const ConvArgs& a, int flags, int num_sms) {
  const EpiParams& e = a.epi;
  if (flags & (MBV_FLAG_NO_PW | MBV_FLAG_NO_CONV_TM)) return false;
  if (prec != 2 || num_sms < 2 || a.gate) return false;
  if (a.Cp_in != 128 || a.N_total != 128 || a.x_ld != 128 || a.n_phases != 1 || a.taps < 5 || a.L_in != a.L_out) return false;
  if (e.bias == nullptr || e.mask != nullptr || e.ld != 128 || e.n_valid != 128 || e.row_mul != 1 || e.rows_res != a.L_out) return false;
  const bool plain_rows = e.rows_out == a.L_out && e.row_add == 0 && e.dup_src < 0;
  const bool one_act = e.n_act == 1 && e.act[0] != nullptr && e.act_add[0] == nullptr;
  if (e.mode == EPI_ACT) {
    if (!one_act || e.xout != nullptr || !plain_rows) return false;
  } else if (e.mode == EPI_RES) {
    if (e.res_half != 1 || e.xin == nullptr) return false;
    if (e.sum_mode == 0) { if (e.xout == nullptr || !one_act || !plain_rows) return false; }
    else if (e.sum_mode == 1 || e.sum_mode == 2) { if (e.xs == nullptr || e.n_act != 0 || e.xout != nullptr) return false; }
    else if (e.sum_mode == 3) {
      if (e.xs == nullptr || !one_act || e.xout != nullptr || e.row_add < 0 || e.rows_out < a.L_out + e.row_add) return false;
      if (e.dup_src >= 0 && (e.dup_dst < 0 || e.dup_dst >= e.rows_out)) return false;
    } else return false;
  } else return false;
  const int halo = (a.taps - 1) * a.dil;
  const int box_rows = ((CT_ROWS + halo + 1) / 2 + 7) / 8 * 8;
  if (box_rows > 256) return false;
  const int slab = CT_SLAB_STAGES * CT_KB * 2 * box_rows * TC_ROW_BYTES;
  return slab + 3 * CT_W_STAGE_BYTES + 2048 <= 224 * 1024;
}

const char* ct_make_plan(int prec, const ConvArgs& a, int num_sms, TcPlan* plan) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tc_tensormap_encoder());
  if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
  const int halo = (a.taps - 1) * a.dil;
  const int box_rows = ((CT_ROWS + halo + 1) / 2 + 7) / 8 * 8;
  plan->pw = 3;
  plan->pw_kblocks = CT_KB;
  plan->box_rows = box_rows;
  plan->pw_tiles = (a.L_out + CT_ROWS - 1) / CT_ROWS;                  // 256-row tiles per utterance
  plan->pw_R = a.B * plan->pw_tiles;                                    // tiles in total
  plan->pw_w_bytes = 2 * box_rows * TC_ROW_BYTES;                       // one k-block of a slab (two boxes)
  plan->pw_a_stage_bytes = CT_KB * plan->pw_w_bytes;
  plan->pw_a_off = CT_SLAB_STAGES * plan->pw_a_stage_bytes;             // weight ring starts here
  int stages = (227 * 1024 - 1024 - plan->pw_a_off - 2 * 128 * 4 - 256) / CT_W_STAGE_BYTES;   // all of the 227 KB: the k = 11, dilation 5 slabs leave room for 4
  if (stages > 6) stages = 6;
  if (stages < 3) return "conv (time on lane): not enough shared memory for the weight ring";
  plan->pw_a_stages = stages;
  plan->pw_bias_off = plan->pw_a_off + stages * CT_W_STAGE_BYTES;
  plan->pw_bar_off = plan->pw_bias_off + 2 * 128 * 4;
  plan->smem_bytes = 1024 + plan->pw_bar_off + (2 * CT_SLAB_STAGES + 2 * stages + 4) * 8 + 16;
  if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;    // one CTA per SM: each CTA owns all 512 TMEM columns
  const int n_pair_tiles = (plan->pw_R + 1) / 2;
  const int pairs = n_pair_tiles < num_sms / 2 ? n_pair_tiles : num_sms / 2;
  plan->grid = 2 * (pairs < 1 ? 1 : pairs);
  plan->n_time = CT_ROWS;
  plan->cluster = 2;
  const CUtensorMapDataType dt = prec == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.L_in, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.x_ld * 2, (cuuint64_t)a.L_in * a.x_ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&plan->tmA, dt, 3, const_cast<void*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the time-on-lane conv activation map";
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.taps * a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&plan->tmB, dt, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the time-on-lane conv weight map";
  }
  plan->tmR = plan->tmA; plan->tmS = plan->tmA; plan->tmBh = plan->tmB;  // unused
  return nullptr;
}

template <typename Op, int MODE, int SM>
static cudaError_t ct_launch_one(const ConvArgs& a, const TcPlan& p, const CtRt& rt, cudaStream_t st, int pdl, bool set_attr) {
  auto k = conv_tm_kernel<Op, MODE, SM>;
  if (set_attr) return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(CT_THREADS);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, k, p.tmA, p.tmB, a.epi, rt);
}

static cudaError_t launch_ct(int prec, const ConvArgs& a, const TcPlan& p, cudaStream_t st, int pdl) {
  CtRt rt;
  rt.B = a.B; rt.L = a.L_out; rt.t_tiles = p.pw_tiles; rt.total_tiles = p.pw_R; rt.n_pair_tiles = (p.pw_R + 1) / 2;
  rt.taps = a.taps; rt.dil = a.dil; rt.shift0 = a.shift0[0]; rt.box_rows = p.box_rows;
  rt.slab_kb_bytes = p.pw_w_bytes; rt.slab_stage_bytes = p.pw_a_stage_bytes; rt.n_w_stages = p.pw_a_stages;
  rt.w_off = p.pw_a_off; rt.bias_off = p.pw_bias_off; rt.bar_off = p.pw_bar_off;
  static const int no_res_pf = getenv("MBV_NO_RES_PF") ? atoi(getenv("MBV_NO_RES_PF")) : 0;  // A/B measurements only
  rt.res_pf = no_res_pf ? 0 : 1;
  if (prec != 2) return cudaErrorInvalidValue;
  if (a.epi.mode == EPI_ACT) return ct_launch_one<OpBF16, EPI_ACT, 0>(a, p, rt, st, pdl, false);
  switch (a.epi.sum_mode) {
    case 0: return ct_launch_one<OpBF16, EPI_RES, 0>(a, p, rt, st, pdl, false);
    case 1: return ct_launch_one<OpBF16, EPI_RES, 1>(a, p, rt, st, pdl, false);
    case 2: return ct_launch_one<OpBF16, EPI_RES, 2>(a, p, rt, st, pdl, false);
    case 3: return ct_launch_one<OpBF16, EPI_RES, 3>(a, p, rt, st, pdl, false);
    default: return cudaErrorInvalidValue;
  }
}

template <typename Op, int MODE, int RH>
static cudaError_t pw_launch_one(const ConvArgs& a, const TcPlan& p, const PwRt& rt, cudaStream_t st, int pdl, bool set_attr) {
  auto k = pw_tc_kernel<Op, MODE, RH>;
  if (set_attr) return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(PW_THREADS);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, p.tmA, p.tmB, p.tmR, a.epi, rt);
}

static cudaError_t pw_dispatch(int prec, const ConvArgs& a, const TcPlan& p, const PwRt& rt, cudaStream_t st, int pdl, bool set_attr,
                               int mode, int rh) {
  if (prec == 2) {
    if (mode == EPI_RS && rh == 1) return pw_launch_one<OpBF16, EPI_RS, 1>(a, p, rt, st, pdl, set_attr);
    if (mode == EPI_ACT && rh == 1) return pw_launch_one<OpBF16, EPI_ACT, 1>(a, p, rt, st, pdl, set_attr);
    if (mode == EPI_ACT && rh == 0) return pw_launch_one<OpBF16, EPI_ACT, 0>(a, p, rt, st, pdl, set_attr);
    if (mode == EPI_POST) return pw_launch_one<OpBF16, EPI_POST, 0>(a, p, rt, st, pdl, set_attr);
  } else if (prec == 3) {
    if (mode == EPI_POST) return pw_launch_one<OpF16, EPI_POST, 0>(a, p, rt, st, pdl, set_attr);
    if (mode == EPI_RS && rh == 1) return pw_launch_one<OpF16, EPI_RS, 1>(a, p, rt, st, pdl, set_attr);
    if (mode == EPI_RS && rh == 2) return pw_launch_one<OpF16, EPI_RS, 2>(a, p, rt, st, pdl, set_attr);
    if (mode == EPI_ACT && rh == 1) return pw_launch_one<OpF16, EPI_ACT, 1>(a, p, rt, st, pdl, set_attr);
    if (mode == EPI_ACT && rh == 0) return pw_launch_one<OpF16, EPI_ACT, 0>(a, p, rt, st, pdl, set_attr);
  }
  return cudaErrorInvalidValue;
}

cudaError_t pw_set_attributes() {
  ConvArgs a{};
  TcPlan p{};
  PwRt rt{};
  const int combos[9][3] = {{2, EPI_RS, 1}, {2, EPI_ACT, 1}, {2, EPI_ACT, 0}, {3, EPI_RS, 1}, {3, EPI_RS, 2}, {3, EPI_ACT, 1}, {3, EPI_ACT, 0},
                            {2, EPI_POST, 0}, {3, EPI_POST, 0}};
  for (auto& c : combos) {
    cudaError_t e = pw_dispatch(c[0], a, p, rt, nullptr, 0, true, c[1], c[2]);
    if (e != cudaSuccess) return e;
  }
  GtRt g{};
  cudaError_t e = gt_launch_one<OpBF16>(a, p, g, nullptr, 0, true);
  if (e == cudaSuccess) e = gt_launch_one<OpF16>(a, p, g, nullptr, 0, true);
  if (e == cudaSuccess) { TcPairPlan pp{}; PtRt pr{}; e = ptm_launch_one<OpBF16>(a, pp, pr, nullptr, 0, true); }
  CtRt cr{};
  if (e == cudaSuccess) e = ct_launch_one<OpBF16, EPI_ACT, 0>(a, p, cr, nullptr, 0, true);
  if (e == cudaSuccess) e = ct_launch_one<OpBF16, EPI_RES, 0>(a, p, cr, nullptr, 0, true);
  if (e == cudaSuccess) e = ct_launch_one<OpBF16, EPI_RES, 1>(a, p, cr, nullptr, 0, true);
  if (e == cudaSuccess) e = ct_launch_one<OpBF16, EPI_RES, 2>(a, p, cr, nullptr, 0, true);
  if (e == cudaSuccess) e = ct_launch_one<OpBF16, EPI_RES, 3>(a, p, cr, nullptr, 0, true);
  return e;
}

cudaError_t launch_pw(int prec, const ConvArgs& a, const TcPlan& p, cudaStream_t st, int pdl) {
  if (p.pw == 2) return launch_gt(prec, a, p, st, pdl);
  if (p.pw == 3) return launch_ct(prec, a, p, st, pdl);
  PwRt rt;
  rt.R = p.pw_R; rt.kblocks = p.pw_kblocks; rt.N = p.pw_N; rt.n_tiles = p.pw_tiles;
  rt.n_a_stages = p.pw_a_stages; rt.w_bytes = p.pw_w_bytes;
  rt.a_off = p.pw_a_off; rt.bias_off = p.pw_bias_off; rt.bar_off = p.pw_bar_off;
  rt.res_c0 = a.epi.mode == EPI_POST ? a.epi.ch_off : 0; rt.res_cols = p.pw_N; rt.res_box = a.epi.mode == EPI_POST ? 32 : 64;
  const int rh = (a.epi.mode == EPI_RS) ? a.epi.res_half : (a.epi.mode == EPI_POST ? 0 : (a.epi.xout != nullptr ? 1 : 0));
  return pw_dispatch(prec, a, p, rt, st, pdl, false, a.epi.mode, rh);
}

}  // namespace mbv
