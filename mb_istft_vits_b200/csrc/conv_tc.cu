// conv_tc.cu -- implicit-GEMM Conv1d / polyphase ConvTranspose1d on the 5th-gen tensor cores (sm_100a).
//
//   D[time 128, N] (+)= sum_tap sum_kblock  A_tap[time 128, 64ch] * W_tap[N, 64ch]^T
//
// * activations are channels-last, so BOTH operands are K-major: A = a 128-row time tile of the
//   activation (rows shifted by the tap offset), B = one tap of the packed weight [N][C_in].
// * TMA (cp.async.bulk.tensor, 128B swizzle) stages operands; per k-block ONE activation slab with the
//   halo of all taps (128 + (taps-1)*dil rows) is loaded and every tap's MMA reads it through a
//   row-offset shared-memory descriptor, so activations cross L2->SMEM once, not `taps` times.
//   Out-of-range rows (utterance edges = the conv's zero padding) are zero-filled by TMA itself
//   through a 3-D (C, time, utterance) tensor map.
// * tcgen05.mma (cta_group::1, M=128, N<=256, kind::f16 bf16 or kind::tf32) accumulates in TMEM;
//   two accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
//   (tcgen05.ld -> registers -> fused bias / residual / leaky-relu / gate / mask -> global).
// * persistent: grid = min(#tiles, #SMs), static round-robin tile order.
//
// Reference semantics implemented: F.conv1d 'same' (commons.py:14-15, modules.py:191-206,220-224),
// F.conv_transpose1d as S polyphase branches (models.py:320-323), WN gate (commons.py:100-107).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "common.cuh"
#include "kernels.h"

namespace mbv {

constexpr int TC_M = 128;            // time rows per tile (UMMA M)
constexpr int TC_THREADS = 192;      // 6 warps
constexpr int TC_ROW_BYTES = 128;    // one swizzle row = 64 bf16 / 32 tf32 channels
constexpr uint64_t TC_TIMEOUT_CYCLES = 4000000000ull;  // ~2 s: a stuck pipeline traps instead of hanging the box

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if ((unsigned long long)(clock64() - t0) > TC_TIMEOUT_CYCLES) {
      printf("mbistft conv_tc: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int KIND>  // 2 = bf16 (kind::f16), 1 = tf32
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (KIND == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  base_offset [49,52), layout SWIZZLE_128B=2 [61,64).)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

struct TcRt {  // runtime scalars the kernel needs beyond ConvArgs
  int slab_rows, a_stage_bytes, b_stage_bytes, n_a_stages, n_b_stages, per_tap;
  int m_tiles, n_tiles, total_tiles, tmem_cols;
};

template <typename Op>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvArgs a, const TcRt rt) {
  using T = typename Op::T;
  constexpr int KB = TC_ROW_BYTES / (int)sizeof(T);  // channels per k-block
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][barriers][tmem ptr]; dynamic smem base is 1024-aligned by the launch
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smA + (size_t)rt.n_a_stages * rt.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)rt.n_b_stages * rt.b_stage_bytes);
  // barrier indices
  const int iAF = 0, iAE = iAF + rt.n_a_stages, iBF = iAE + rt.n_a_stages, iBE = iBF + rt.n_b_stages;
  const int iCF = iBE + rt.n_b_stages, iCE = iCF + 2, nBars = iCE + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + nBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < rt.n_a_stages; ++i) { mbar_init(BAR(iAF + i), 1); mbar_init(BAR(iAE + i), 1); }
    for (int i = 0; i < rt.n_b_stages; ++i) { mbar_init(BAR(iBF + i), 1); mbar_init(BAR(iBE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(iCF + i), 1); mbar_init(BAR(iCE + i), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), (uint32_t)rt.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int kblocks = a.Cp_in / KB;
  const int box_n = a.gate ? a.N_tile / 2 : a.N_tile;     // weight rows per TMA box
  const int cols_logical = a.gate ? a.N_tile / 2 : a.N_tile;  // logical output columns per tile
  const uint32_t a_box_bytes = (uint32_t)rt.slab_rows * TC_ROW_BYTES;
  const uint32_t b_box_bytes = (uint32_t)a.N_tile * TC_ROW_BYTES;
  const int a_pt_off = (int)b_box_bytes;  // per-tap mode: the A tile sits after the weights in a B stage

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = blockIdx.x; tile < rt.total_tiles; tile += gridDim.x) {
      int rest = tile;
      const int nt = rest % rt.n_tiles; rest /= rt.n_tiles;
      const int phase = rest % a.n_phases; rest /= a.n_phases;
      const int mt = rest % rt.m_tiles;
      const int b = rest / rt.m_tiles;
      const int t0 = mt * TC_M;
      const int n0 = nt * cols_logical;
      const int wrow0 = phase * a.taps * a.N_total;
      for (int kb = 0; kb < kblocks; ++kb) {
        if (!rt.per_tap) {
          mbar_wait(BAR(iAE + sa), pa ^ 1);
          mbar_expect_tx(BAR(iAF + sa), a_box_bytes);
          tma_load_3d(smem_u32(smA + (size_t)sa * rt.a_stage_bytes), &tmA, BAR(iAF + sa), kb * KB,
                      t0 + a.shift0[phase], b);
          if (++sa == rt.n_a_stages) { sa = 0; pa ^= 1; }
        }
        for (int tap = 0; tap < a.taps; ++tap) {
          mbar_wait(BAR(iBE + sb), pb ^ 1);
          uint8_t* stage = smB + (size_t)sb * rt.b_stage_bytes;
          mbar_expect_tx(BAR(iBF + sb), b_box_bytes + (rt.per_tap ? a_box_bytes : 0u));
          if (rt.per_tap)
            tma_load_3d(smem_u32(stage + a_pt_off), &tmA, BAR(iBF + sb), kb * KB,
                        t0 + a.shift0[phase] + tap * a.dil, b);
          const int wrow = wrow0 + tap * a.N_total + n0;
          tma_load_2d(smem_u32(stage), &tmB, BAR(iBF + sb), kb * KB, wrow);
          if (a.gate)
            tma_load_2d(smem_u32(stage + (size_t)box_n * TC_ROW_BYTES), &tmB, BAR(iBF + sb), kb * KB,
                        wrow + a.N_total / 2);
          if (++sb == rt.n_b_stages) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16 or tf32, K-major both, N, M=128
    constexpr uint32_t fmt = (Op::kPrec == 2) ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.N_tile >> 3) << 17) |
                           ((uint32_t)(TC_M >> 4) << 24);
    int sa = 0, sb = 0, sc = 0;
    uint32_t pa = 0, pb = 0, pc = 0;
    for (int tile = blockIdx.x; tile < rt.total_tiles; tile += gridDim.x) {
      mbar_wait(BAR(iCE + sc), pc ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(sc * a.N_tile);
      uint32_t accum = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        uint32_t a_base = 0;
        if (!rt.per_tap) {
          mbar_wait(BAR(iAF + sa), pa);
          tc_fence_after();
          a_base = smem_u32(smA + (size_t)sa * rt.a_stage_bytes);
        }
        for (int tap = 0; tap < a.taps; ++tap) {
          mbar_wait(BAR(iBF + sb), pb);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smB + (size_t)sb * rt.b_stage_bytes);
          // Tap offset = row offset into the slab.  The 128B swizzle is a function of the absolute shared-memory
          // address bits (measured on B200: base_offset 0 is exact for any row offset, (row & 7) is wrong), so a
          // descriptor that merely starts `roff` rows later addresses exactly the rows TMA wrote.
          const uint32_t a_addr = rt.per_tap ? b_addr + (uint32_t)a_pt_off
                                             : a_base + (uint32_t)(tap * a.dil) * TC_ROW_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_smem_desc(a_addr + k * 32, 0);
            const uint64_t bd = make_smem_desc(b_addr + k * 32, 0);
            tc_mma<(Op::kPrec == 2) ? 2 : 1>(tmem_d, ad, bd, idesc, accum);
            accum = 1;
          }
          tc_commit(BAR(iBE + sb));
          if (++sb == rt.n_b_stages) { sb = 0; pb ^= 1; }
        }
        if (!rt.per_tap) {
          tc_commit(BAR(iAE + sa));
          if (++sa == rt.n_a_stages) { sa = 0; pa ^= 1; }
        }
      }
      tc_commit(BAR(iCF + sc));
      if (++sc == 2) { sc = 0; pc ^= 1; }
    }
  } else if (warp >= 2) {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int sc = 0;
    uint32_t pc = 0;
    for (int tile = blockIdx.x; tile < rt.total_tiles; tile += gridDim.x) {
      int rest = tile;
      const int nt = rest % rt.n_tiles; rest /= rt.n_tiles;
      const int phase = rest % a.n_phases; rest /= a.n_phases;
      const int mt = rest % rt.m_tiles;
      const int b = rest / rt.m_tiles;
      const int row = mt * TC_M + q * 32 + lane;
      const int n0 = nt * cols_logical;
      mbar_wait(BAR(iCF + sc), pc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sc * a.N_tile);
      for (int c = 0; c < cols_logical; c += 16) {
        float acc[16], acc2[16];
        tmem_ld16(taddr + (uint32_t)c, acc);
        if (a.gate) tmem_ld16(taddr + (uint32_t)(cols_logical + c), acc2);
        tmem_ld_wait();
        if (row < a.L_out) epilogue_chunk<Op, 16>(a.epi, b, row, phase, n0 + c, acc, acc2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(iCE + sc));
      if (++sc == 2) { sc = 0; pc ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)rt.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

static int next_pow2_cols(int c) {
  int p = 32;
  while (p < c) p <<= 1;
  return p;
}

const char* tc_make_plan(int prec, const ConvArgs& a, int flags, int num_sms, TcPlan* plan) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
  const int esize = prec == 2 ? 2 : 4;
  const int KB = TC_ROW_BYTES / esize;
  if (a.Cp_in % 64 != 0) return "tcgen05 conv: padded input channels must be a multiple of 64";
  if (a.N_tile % 16 != 0 || a.N_tile < 16 || a.N_tile > 256) return "tcgen05 conv: N tile must be 16..256, multiple of 16";
  if (a.gate && (a.N_tile % 32 != 0)) return "tcgen05 conv: gate tile must be a multiple of 32";
  const int cols_logical = a.gate ? a.N_tile / 2 : a.N_tile;
  const int NL = a.gate ? a.N_total / 2 : a.N_total;
  if (NL % cols_logical != 0) return "tcgen05 conv: N tile must divide the padded output channels";
  plan->per_tap = (flags & 1) ? 1 : 0;
  const int halo = (a.taps - 1) * a.dil;
  plan->slab_rows = plan->per_tap ? TC_M : TC_M + halo;
  if (plan->slab_rows > 256) return "tcgen05 conv: activation slab exceeds the 256-row TMA box limit";
  const int a_bytes = ((plan->slab_rows * TC_ROW_BYTES + 1023) / 1024) * 1024;
  const int b_bytes = a.N_tile * TC_ROW_BYTES;
  const int budget = 200 * 1024;
  if (plan->per_tap) {
    plan->a_stage_bytes = 0;
    plan->n_a_stages = 0;
    plan->b_stage_bytes = b_bytes + a_bytes;
    plan->n_b_stages = budget / plan->b_stage_bytes;
  } else {
    plan->a_stage_bytes = a_bytes;
    plan->n_a_stages = 3;
    plan->b_stage_bytes = b_bytes;
    plan->n_b_stages = (budget - 3 * a_bytes) / b_bytes;
  }
  if (plan->n_b_stages > 8) plan->n_b_stages = 8;
  if (plan->n_b_stages < 2) return "tcgen05 conv: not enough shared memory for two weight stages";
  plan->m_tiles = (a.L_out + TC_M - 1) / TC_M;
  plan->n_tiles = NL / cols_logical;
  plan->total_tiles = a.B * a.n_phases * plan->m_tiles * plan->n_tiles;
  plan->tmem_cols = next_pow2_cols(2 * a.N_tile);
  if (plan->tmem_cols > 512) return "tcgen05 conv: accumulators exceed TMEM";
  const int nbars = 2 * plan->n_a_stages + 2 * plan->n_b_stages + 4;
  plan->smem_bytes = 1024 + plan->n_a_stages * plan->a_stage_bytes + plan->n_b_stages * plan->b_stage_bytes +
                     nbars * 8 + 16;
  // keep one CTA per SM (each CTA wants up to all 512 TMEM columns)
  if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;
  plan->grid = plan->total_tiles < num_sms ? plan->total_tiles : num_sms;
  if (plan->grid < 1) plan->grid = 1;

  const CUtensorMapDataType dt = prec == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.L_in, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.Cp_in * esize, (cuuint64_t)a.L_in * a.Cp_in * esize};
    cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)plan->slab_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&plan->tmA, dt, 3, const_cast<void*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the activation map";
  }
  {
    const int box_n = a.gate ? a.N_tile / 2 : a.N_tile;
    cuuint64_t dims[2] = {(cuuint64_t)a.Cp_in, (cuuint64_t)a.n_phases * a.taps * a.N_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cp_in * esize};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&plan->tmB, dt, 2, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the weight map";
  }
  return nullptr;
}

cudaError_t tc_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<OpBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_tc_kernel<OpTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_conv_tc(int prec, const ConvArgs& a, const TcPlan& p, cudaStream_t st) {
  TcRt rt;
  rt.slab_rows = p.slab_rows; rt.a_stage_bytes = p.a_stage_bytes; rt.b_stage_bytes = p.b_stage_bytes;
  rt.n_a_stages = p.n_a_stages; rt.n_b_stages = p.n_b_stages; rt.per_tap = p.per_tap;
  rt.m_tiles = p.m_tiles; rt.n_tiles = p.n_tiles;
  rt.total_tiles = p.total_tiles; rt.tmem_cols = p.tmem_cols;
  if (prec == 2)
    conv_tc_kernel<OpBF16><<<p.grid, TC_THREADS, p.smem_bytes, st>>>(p.tmA, p.tmB, a, rt);
  else
    conv_tc_kernel<OpTF32><<<p.grid, TC_THREADS, p.smem_bytes, st>>>(p.tmA, p.tmB, a, rt);
  return cudaGetLastError();
}

}  // namespace mbv
