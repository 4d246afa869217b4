// tail.cu -- fused decoder tail: exp / pi*sin head -> 16-point inverse real DFT (registers) -> periodic Hann
// window -> overlap-add (hop 4) with the exact edge envelope -> {nothing | 4-band synthesis FIR with the x4
// zero-stuffing folded into a polyphase form} -> fp32 waveform.  HBM-bound by design: per latent frame it reads
// 16 frames x 72 logits x 4 B = 4608 B and writes 256 samples x 4 B = 1024 B.
//
// Reference semantics: models.py:366-375 (head, reshape), stft.py:197-202 (torch.istft,
// n_fft 16, hop 4, center=True, periodic Hann), pqmf.py:105-116 (zero-stuff x4 with gain 4, pad 31, 63-tap
// cross-correlation), models.py:463-465 (MS: same with the trainable multistream_conv_post).
//
// Two kernels: tail_mb3_kernel (4 sub-bands + synthesis FIR: MB / MS decoders) and tail_sb3_kernel (single band).
// Both keep the frames in registers (lane = STFT frame, overlap-add by warp shuffles); see the comments above each.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace mbv {

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

__device__ __forceinline__ float ptx_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <bool PRECISE>
__device__ __forceinline__ void head(float xm, float xp, float& mag, float& ph, float& re, float& im) {
  // spec = exp(x), phase = pi * sin(x)  (models.py:368-369); re/im = spec * (cos, sin)(phase)
  float s, c;
  if (PRECISE) {
    mag = expf(xm);
    ph = 3.14159265358979323846f * sinf(xp);
    sincosf(ph, &s, &c);
  } else {
    // MUFU.EX2 / MUFU.SIN / MUFU.COS; the sine's own argument reduction keeps the absolute error ~1e-6 for the
    // O(1) logits this head sees (|x| < ~100)
    mag = ptx_ex2(xm * 1.4426950408889634f);
    ph = 3.14159265358979323846f * ptx_sin(xp);
    s = ptx_sin(ph);
    c = ptx_cos(ph);
  }
  re = mag * c;
  im = mag * s;
}

// Two fp32 values in one 64-bit register pair: sm_100 executes add / mul / fma on both halves with ONE instruction
// (FADD2 / FMUL2 / FFMA2, IEEE round-to-nearest like the scalar forms).  The multi-band tail puts two sub-bands of the
// same frame in the two halves, which halves the instruction count of everything between the head and the FIR.
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 dup2(float a) { return mk2(a, a); }
__device__ __forceinline__ float lo2(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); (void)y; return x; }
__device__ __forceinline__ float hi2(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); (void)x; return y; }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator*(float a, f2 b) { return dup2(a) * b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

// 16-point inverse real DFT (imaginary parts of bins 0 and 8 ignored, like torch.istft's irfft) times the periodic
// Hann window / 16, written as an even/odd-bin split so that it costs ~100 flops instead of 16 x 15 MACs:
//   x[n] = E[n] + O[n], x[n+8] = E[n] - O[n];  E = bins {0,4,8} (period 4) +- bins {2,6};  O = odd bins with the
//   n <-> 8-n symmetry of cos / antisymmetry of sin.   T = float, or f2 for two sub-bands at once.
template <typename T>
__device__ __forceinline__ void idft16_windowed(const T* re, const T* im, T* fr) {
  constexpr float c1 = 0.92387953251128674f, c2 = 0.70710678118654752f, c3 = 0.38268343236508977f;
  const T a0 = re[0] + re[8], a1 = re[0] - re[8];
  const T A0 = a0 + 2.f * re[4], A1 = a1 - 2.f * im[4], A2 = a0 - 2.f * re[4], A3 = a1 + 2.f * im[4];
  const T ims = im[2] + im[6];
  const T B0 = 2.f * (re[2] + re[6]);
  const T B1 = (2.f * c2) * ((re[2] - re[6]) - ims);
  const T B2 = 2.f * (im[6] - im[2]);
  const T B3 = (2.f * c2) * ((re[6] - re[2]) - ims);
  T E[8] = {A0 + B0, A1 + B1, A2 + B2, A3 + B3, A0 - B0, A1 - B1, A2 - B2, A3 - B3};
  const T r17 = re[1] - re[7], r35 = re[3] - re[5], i17 = im[1] + im[7], i35 = im[3] + im[5];
  const T oc0 = (re[1] + re[7]) + (re[3] + re[5]);
  const T oc1 = c1 * r17 + c3 * r35, os1 = c3 * i17 + c1 * i35;
  const T oc2 = c2 * ((re[1] + re[7]) - (re[3] + re[5])), os2 = c2 * ((im[1] - im[7]) + (im[3] - im[5]));
  const T oc3 = c3 * r17 - c1 * r35, os3 = c1 * i17 - c3 * i35;
  const T os4 = (im[1] - im[3]) + (im[5] - im[7]);
  T O[8];
  O[0] = 2.f * oc0;
  O[1] = 2.f * (oc1 - os1); O[7] = -2.f * (oc1 + os1);
  O[2] = 2.f * (oc2 - os2); O[6] = -2.f * (oc2 + os2);
  O[3] = 2.f * (oc3 - os3); O[5] = -2.f * (oc3 + os3);
  O[4] = -2.f * os4;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    fr[n] = kWin16[n] * (E[n] + O[n]);
    fr[n + 8] = kWin16[n + 8] * (E[n] - O[n]);
  }
}

__device__ __forceinline__ void t2_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && (unsigned long long)(clock64() - t0) > 4000000000ull) __trap();
  }
}

// ------------------------------------------------------------------------------------------------
// Multi-band / multi-stream tail.  The round-1 kernel it replaced was bound by shared-memory wavefronts (17.6 M per
// launch, 63 % L1 pipe, profiles/r01_ncu_full_final.txt): every intermediate (frames, sub-band signal, modulated
// signal) made a round trip through shared memory.  This one keeps them in registers:
//   * lane = STFT frame, all four bands in the same thread.  The logits tile arrives by TMA with the 128-byte
//     swizzle, so "one 288-byte row per lane" reads are bank-conflict-free 16-byte loads (18 per frame).
//   * overlap-add = 12 warp shuffles per band (lane l owns hop block l: frames l, l+1, l+2, l+3); only the 6 vectors
//     per band that cross a warp boundary go through shared memory.
//   * the thread then holds its hop block of all four bands: envelope, optional o_mb store and the PQMF cosine
//     modulation (8 rows) happen in registers; only U[8][512] is written to shared memory.
//   * synthesis FIR: thread = (pair of hop blocks, residue pair): two 24-float windows serve 8 outputs (before: 20-float
//     windows for 4), rows of the two residue classes 516 floats apart so the 16-byte loads of a quarter warp hit
//     disjoint banks.  Outputs are staged (XOR-swizzled) in the dead logits buffer and leave as coalesced 16-byte stores.
//   * persistent CTAs of 128 threads (tiles of 128 frames, <= 121 owned hop blocks), 4 per SM, each owning an EQUAL
//     contiguous range of hop blocks (tiles never straddle an utterance): no wave-quantisation tail.  (-DMBV_T3_NF=256,
//     2 CTAs of 8 warps per SM, measures 8 % slower: 91.6 vs 84.3 us.)
// ------------------------------------------------------------------------------------------------
#ifndef MBV_T3_NF
#define MBV_T3_NF 128
#endif
constexpr int T3_NF = MBV_T3_NF;           // frames per tile = threads per CTA (128 or 256)
constexpr int T3_CTAS = 512 / T3_NF;       // CTAs per SM: 16 resident warps either way
constexpr int T3_NW = T3_NF / 32;
constexpr int T3_NQ = T3_NF - 7;           // owned hop blocks per tile (max)
constexpr int T3_UP = 4 * T3_NF + 4;       // floats per U / Y row; = 4 mod 32 (bank spread between adjacent rows)
constexpr int T3_BOX01 = T3_NF * 128;      // bytes of one 32-float box
constexpr int T3_BOX2 = T3_NF * 32;        // bytes of the 8-float box
constexpr int T3_OFF_U = 2 * T3_BOX01 + T3_BOX2;              // 36864
constexpr int T3_OFF_HALO = T3_OFF_U + 8 * T3_UP * 4;         // + 16512
constexpr int T3_OFF_TAB = T3_OFF_HALO + T3_NW * 4 * 6 * 16;
constexpr int T3_OFF_BAR = T3_OFF_TAB + 1024;
constexpr int T3_SMEM = T3_OFF_BAR + 16;

__device__ __forceinline__ float4 shfl_down4(const float* v, int delta) {
  float4 r;
  r.x = __shfl_down_sync(0xffffffffu, v[0], delta);
  r.y = __shfl_down_sync(0xffffffffu, v[1], delta);
  r.z = __shfl_down_sync(0xffffffffu, v[2], delta);
  r.w = __shfl_down_sync(0xffffffffu, v[3], delta);
  return r;
}
__device__ __forceinline__ void add4(float4& y, const float4& v) { y.x += v.x; y.y += v.y; y.z += v.z; y.w += v.w; }

// 16-byte chunk c (0..17) of frame row r of the swizzled logits tile
template <int C>
__device__ __forceinline__ float4 t3_chunk(const uint8_t* sm, int r) {
  if constexpr (C < 16) {
    constexpr int box = C >> 3, cc = C & 7;
    return *reinterpret_cast<const float4*>(sm + box * T3_BOX01 + r * 128 + ((cc ^ (r & 7)) << 4));
  } else {
    constexpr int cc = C - 16;
    return *reinterpret_cast<const float4*>(sm + 2 * T3_BOX01 + r * 32 + ((cc ^ ((r >> 2) & 1)) << 4));
  }
}
template <int C0, int N>
struct T3Load {
  static __device__ __forceinline__ void run(const uint8_t* sm, int r, float* buf) {
    const float4 v = t3_chunk<C0>(sm, r);
    buf[0] = v.x; buf[1] = v.y; buf[2] = v.z; buf[3] = v.w;
    T3Load<C0 + 1, N - 1>::run(sm, r, buf + 4);
  }
};
template <int C0> struct T3Load<C0, 0> { static __device__ __forceinline__ void run(const uint8_t*, int, float*) {} };

template <bool PRECISE, bool EMIT>
__global__ void __launch_bounds__(T3_NF, T3_CTAS)
tail_mb3_kernel(const __grid_constant__ CUtensorMap tm32, const __grid_constant__ CUtensorMap tm8,
                const __grid_constant__ TailArgs a) {
  constexpr int S = 4;
  extern __shared__ __align__(1024) uint8_t sm3[];
  float* s_u = reinterpret_cast<float*>(sm3 + T3_OFF_U);          // [8][T3_UP]  (generic filter: rows 0..3 = Y)
  f2* s_halo = reinterpret_cast<f2*>(sm3 + T3_OFF_HALO);           // [warp][band pair][6 parts][4 samples]
  float* s_tab = reinterpret_cast<float*>(sm3 + T3_OFF_TAB);      // fast: (g2, g2)[4][16] duplicated pairs; generic: coef[4][64]
  float* s_out = s_u;  // output staging [64 rows of 2 hop blocks][32], chunks XOR-swizzled; reuses U once the FIR has read it
  const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(sm3 + T3_OFF_BAR));
  const uint32_t sm_base = static_cast<uint32_t>(__cvta_generic_to_shared(sm3));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, F = L + 1;

  // this CTA's equal share of the B*L hop blocks.  The walk keeps (utterance, block, blocks left) as 32-bit state: the
  // first version recomputed them from a 64-bit global position with two 64-bit divisions per tile in every thread,
  // ~3 K cycles on the critical path between the phase-A barrier and the next tile's TMA issue (MBV_TAIL_TIMELINE).
  const long long total = (long long)a.B * L;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long g0 = per * blockIdx.x;
  const long long g_end = (g0 + per < total) ? g0 + per : total;
  long long g = g0;  // only used by the debug stamps
  int left_blocks = (int)(g_end > g0 ? g_end - g0 : 0);

  struct Tile { int b, q0, nq; };
  auto make_tile = [&](int b, int q0, int rem) {
    Tile t;
    t.b = b; t.q0 = q0;
    const int tiles_left = (rem + T3_NQ - 1) / T3_NQ;
    int nq = (rem + tiles_left - 1) / tiles_left;
    if (nq > L - q0) nq = L - q0;
    t.nq = nq;
    return t;
  };
  // one thread: three boxes of the logits rows [q0-3, q0-3+128) (OOB rows -> 0).  The tile is only ever READ through the
  // generic proxy, and those reads are ordered before this point by the block barrier, so no proxy fence is needed for
  // the write-after-read (same discipline as a TMA pipeline's consumer-release / producer-acquire).
  auto issue_load = [&](const Tile& t) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)T3_OFF_U) : "memory");
    const int f0 = t.q0 - 3;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(sm_base), "l"(reinterpret_cast<uint64_t>(&tm32)), "r"(bar), "r"(0), "r"(f0), "r"(t.b) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(sm_base + T3_BOX01), "l"(reinterpret_cast<uint64_t>(&tm32)), "r"(bar), "r"(32), "r"(f0), "r"(t.b) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(sm_base + 2 * T3_BOX01), "l"(reinterpret_cast<uint64_t>(&tm8)), "r"(bar), "r"(64), "r"(f0), "r"(t.b) : "memory");
  };

  if (a.fast_pqmf) { if (tid < 128) s_tab[tid] = a.g2[tid >> 5][(tid >> 1) & 15]; }
  else { for (int i = tid; i < 256; i += T3_NF) s_tab[i] = a.coef[i >> 6][i & 63]; }
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm32)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm8)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (left_blocks <= 0) return;
  if (a.dbg && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[64 + 3 * blockIdx.x] = (long long)t;
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.dbg[64 + 3 * blockIdx.x + 2] = smid;
  }
  Tile cur = make_tile((int)(g0 / L), (int)(g0 % L), left_blocks);
  if (tid == 0) issue_load(cur);
  uint32_t parity = 0;

  while (true) {
    const int b = cur.b, Q0 = cur.q0, nq = cur.nq;
    const int QY0 = Q0 - 2, F0 = Q0 - 3;
    const bool last_tile = (Q0 + nq == L);
    long long* dbg = nullptr;
    if (a.dbg && blockIdx.x == 0 && tid == 0) {
      const long long k = (g - per * blockIdx.x) / T3_NQ;
      if (k < 8) dbg = a.dbg + k * 8;
    }
    if (dbg) dbg[0] = clock64();
    t2_mbar_wait(bar, parity);
    parity ^= 1;
    if (dbg) dbg[1] = clock64();

    // ---- phase A: head + inverse DFT + window (registers), overlap-add by shuffles.  Lane = frame F0 + tid = hop block
    // QY0 + tid.  Two sub-bands share every arithmetic instruction (f2): yp[P][i] = sample i of bands (2P, 2P+1).
    f2 yp[2][4];
    {
      const int f = F0 + tid;
      // Frames outside the utterance read TMA's zero fill (finite head values) and get magnitude 0, so the band body is
      // branch-free (v3.0 spent 12 % of its instructions and 40 % of its stall samples on BSSY/BRA/BSYNC).
      const bool live = (f >= 0) && (f < F);
      const bool emit = EMIT && live && (f >= Q0) && (f < Q0 + nq || (last_tile && f == L));
      const f2 keep = dup2(live ? 1.f : 0.f);
      const f2 k1 = dup2(lane < 31 ? 1.f : 0.f), k2 = dup2(lane < 30 ? 1.f : 0.f), k3 = dup2(lane < 29 ? 1.f : 0.f);
      auto band_pair = [&](auto p_tag) {
        constexpr int P = decltype(p_tag)::value, s0 = 2 * P;
        // floats [36P, 36P + 36) of the frame row = chunks 9P .. 9P+8: band s0 then band s0 + 1
        float x[36];
        T3Load<9 * P, 9>::run(sm3, tid, x);
        f2 re[9], im[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          float mag0, mag1, sn0, sn1;
          if (PRECISE) {
            mag0 = expf(x[k]); mag1 = expf(x[18 + k]);
            sn0 = sinf(x[9 + k]); sn1 = sinf(x[27 + k]);
          } else {
            mag0 = ptx_ex2(x[k] * 1.4426950408889634f); mag1 = ptx_ex2(x[18 + k] * 1.4426950408889634f);
            sn0 = ptx_sin(x[9 + k]); sn1 = ptx_sin(x[27 + k]);
          }
          const f2 ph = 3.14159265358979323846f * mk2(sn0, sn1);   // phase = pi * sin(x)  (models.py:369)
          const float ph0 = lo2(ph), ph1 = hi2(ph);
          float s0v, c0v, s1v, c1v;
          if (PRECISE) { sincosf(ph0, &s0v, &c0v); sincosf(ph1, &s1v, &c1v); }
          else { s0v = ptx_sin(ph0); c0v = ptx_cos(ph0); s1v = ptx_sin(ph1); c1v = ptx_cos(ph1); }
          if (EMIT) {
            if (emit) {
              const size_t o = (((size_t)b * S + s0) * 9 + k) * F + f;
              a.spec[o] = mag0; a.phase[o] = ph0;
              a.spec[o + (size_t)9 * F] = mag1; a.phase[o + (size_t)9 * F] = ph1;
            }
          }
          const f2 mag = keep * mk2(mag0, mag1);
          re[k] = mag * mk2(c0v, c1v);
          im[k] = mag * mk2(s0v, s1v);
        }
        f2 fr[16];
        idft16_windowed<f2>(re, im, fr);
        // hop block of this lane = part 3 of its own frame + part 2 / 1 / 0 of the next three frames
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f2 acc = fr[12 + i], v;
          v.v = __shfl_down_sync(0xffffffffu, fr[8 + i].v, 1); acc = fma2(k1, v, acc);
          v.v = __shfl_down_sync(0xffffffffu, fr[4 + i].v, 2); acc = fma2(k2, v, acc);
          v.v = __shfl_down_sync(0xffffffffu, fr[i].v, 3);     acc = fma2(k3, v, acc);
          yp[P][i] = acc;
        }
        // what the previous warp's lanes 29..31 are missing: slots [lane 0: p0 p1 p2 | lane 1: p0 p1 | lane 2: p0].
        // Predicated stores, no branch: both band pairs stay in one basic block and the scheduler can interleave them.
        {
          f2* h = s_halo + ((warp * 2 + P) * 6) * 4 + (lane == 0 ? 0 : (lane == 1 ? 3 : 5)) * 4;
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            const bool on = (warp > 0) && (lane + part < 3);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (on) h[part * 4 + i] = fr[4 * part + i];
          }
        }
      };
      band_pair(std::integral_constant<int, 0>{});
      band_pair(std::integral_constant<int, 1>{});
    }
    if (dbg) dbg[2] = clock64();
    __syncthreads();  // halo visible; the logits tile is dead from here on: refill it with the next tile while B/C run
    if (dbg) dbg[3] = clock64();
    left_blocks -= nq;
    const bool has_next = left_blocks > 0;
    Tile nxt = cur;
    if (has_next) {
      const bool wrap = (Q0 + nq == L);
      nxt = make_tile(wrap ? b + 1 : b, wrap ? 0 : Q0 + nq, left_blocks);
      if (tid == 0) issue_load(nxt);
    }
    if (dbg) dbg[7] = clock64();

    // ---- phase B: finish the blocks that straddle a warp boundary, envelope, optional o_mb, modulation -> U
    {
      const int q = QY0 + tid;
      const bool inside = (q >= 0) && (q < L) && (tid < T3_NF - 3);
#pragma unroll
      for (int P = 0; P < 2; ++P) {
        {
          // lanes 29 / 30 / 31 of warps 0..2 add the 1 / 2 / 3 vectors the next warp left for them.  Unconditional loads
          // from valid addresses + predicated adds: the first version branched three ways here and the serialised
          // LDS -> FADD2 chains of three lanes held every warp (and the block barrier) for ~1 K cycles per tile.
          const bool fix = (warp < T3_NW - 1) && (lane >= 29);
          const bool on1 = fix && (lane >= 30), on2 = fix && (lane == 31);
          const int s0 = lane == 30 ? 1 : (lane == 31 ? 2 : 0), s1 = lane == 30 ? 3 : 4;
          const f2* h = s_halo + ((((fix ? warp + 1 : warp)) * 2 + P) * 6) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const f2 a0 = h[s0 * 4 + i], a1 = h[s1 * 4 + i], a2 = h[5 * 4 + i];
            f2 v = yp[P][i];
            const f2 v0 = v + a0;
            if (fix) v = v0;
            const f2 v1 = v + a1;
            if (on1) v = v1;
            const f2 v2 = v + a2;
            if (on2) v = v2;
            yp[P][i] = v;
          }
        }
        if (!inside) {
#pragma unroll
          for (int i = 0; i < 4; ++i) yp[P][i] = dup2(0.f);
        } else if (PRECISE || q == 0 || q == L - 1) {
          float e[4] = {1.5f, 1.5f, 1.5f, 1.5f};
          if (q == 0) { e[0] -= win_sq(12); e[1] -= win_sq(13); e[2] -= win_sq(14); e[3] -= win_sq(15); }
          if (q == L - 1) { e[0] -= win_sq(0); e[1] -= win_sq(1); e[2] -= win_sq(2); e[3] -= win_sq(3); }
#pragma unroll
          for (int i = 0; i < 4; ++i) yp[P][i] = mk2(lo2(yp[P][i]) / e[i], hi2(yp[P][i]) / e[i]);
        } else {  // steady-state envelope 1.5: multiply by the reciprocal (<= 1 ulp from the division)
#pragma unroll
          for (int i = 0; i < 4; ++i) yp[P][i] = 0.66666666666666667f * yp[P][i];
        }
        if (a.o_mb != nullptr && inside && tid >= 2 && tid < 2 + nq) {
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const int s = 2 * P + hb;
            const float4 v = hb ? make_float4(hi2(yp[P][0]), hi2(yp[P][1]), hi2(yp[P][2]), hi2(yp[P][3]))
                                : make_float4(lo2(yp[P][0]), lo2(yp[P][1]), lo2(yp[P][2]), lo2(yp[P][3]));
            if (a.variant == 1) {
              *reinterpret_cast<float4*>(a.o_mb + ((size_t)b * S + s) * 4 * L + 4 * (size_t)q) = v;
            } else {
              float4* o = reinterpret_cast<float4*>(a.o_mb + ((size_t)b * S + s) * 16 * L + 16 * (size_t)q);
              o[0] = make_float4(4.f * v.x, 0.f, 0.f, 0.f);
              o[1] = make_float4(4.f * v.y, 0.f, 0.f, 0.f);
              o[2] = make_float4(4.f * v.z, 0.f, 0.f, 0.f);
              o[3] = make_float4(4.f * v.w, 0.f, 0.f, 0.f);
            }
          }
        }
      }
      if (a.fast_pqmf) {
        // U[m][j] = sum_c mod[m][c] y_c[j]: (mod[m][0], mod[m][1]) x (y_0, y_1) + (mod[m][2], mod[m][3]) x (y_2, y_3), halves added
        const unsigned long long* modp = reinterpret_cast<const unsigned long long*>(&a.mod[0][0]);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          f2 m01, m23;
          m01.v = modp[2 * m]; m23.v = modp[2 * m + 1];
          float u[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const f2 t = fma2(m23, yp[1][i], m01 * yp[0][i]);
            u[i] = lo2(t) + hi2(t);
          }
          *reinterpret_cast<float4*>(s_u + m * T3_UP + 4 * tid) = make_float4(u[0], u[1], u[2], u[3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int P = c >> 1;
          const float4 v = (c & 1) ? make_float4(hi2(yp[P][0]), hi2(yp[P][1]), hi2(yp[P][2]), hi2(yp[P][3]))
                                   : make_float4(lo2(yp[P][0]), lo2(yp[P][1]), lo2(yp[P][2]), lo2(yp[P][3]));
          *reinterpret_cast<float4*>(s_u + c * T3_UP + 4 * tid) = v;
        }
      }
    }
    __syncthreads();

    if (dbg) dbg[4] = clock64();
    // ---- phase C: synthesis FIR.  Thread = (hop blocks 2p, 2p+1; residues rh and rh+2); sub-band positions j = 8p + e.
    {
      const int p = 16 * warp + (lane >> 1), rh = lane & 1;
      const bool fir_on = (p >= 1 && 2 * p < 2 + nq);
      float accs[2][8];
      if (fir_on) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int r = rh + 2 * h2;
          float* acc = accs[h2];
          if (a.fast_pqmf) {
            // taps d = -7..8 -> prototype tap 4d+31-r; odd index d reads U[7-r], even U[3-r]; window index 1+e+d of
            // [8p-8, 8p+16).  Outputs are paired so that every FFMA2 reads an ALIGNED pair of the window: odd taps
            // accumulate outputs (0,1)(2,3)(4,5)(6,7), even taps (-1,0)(1,2)(3,4)(5,6)(7,8) (two unused), summed at the end.
            f2 we[12], wo[12];
            const ulonglong2* pe = reinterpret_cast<const ulonglong2*>(s_u + (7 - r) * T3_UP + 8 * p - 8);
            const ulonglong2* po = reinterpret_cast<const ulonglong2*>(s_u + (3 - r) * T3_UP + 8 * p - 8);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              const ulonglong2 u = pe[i], v = po[i];
              we[2 * i].v = u.x; we[2 * i + 1].v = u.y;
              wo[2 * i].v = v.x; wo[2 * i + 1].v = v.y;
            }
            const ulonglong2* gp = reinterpret_cast<const ulonglong2*>(s_tab) + r * 8;  // (g_d, g_d) pairs, two taps per load
            f2 accA[4], accB[5];
#pragma unroll
            for (int i = 0; i < 4; ++i) accA[i] = dup2(0.f);
#pragma unroll
            for (int i = 0; i < 5; ++i) accB[i] = dup2(0.f);
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) {
              const ulonglong2 gv = gp[dd];
              f2 ge, go;
              ge.v = gv.x;  // tap index d = 2 dd   (even): window index 2j + d for the output pair (2j-1, 2j)
              go.v = gv.y;  // tap index d = 2 dd+1 (odd):  window index 2j + d + 1 for the output pair (2j, 2j+1)
#pragma unroll
              for (int jj = 0; jj < 5; ++jj) accB[jj] = fma2(ge, wo[jj + dd], accB[jj]);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) accA[jj] = fma2(go, we[jj + dd + 1], accA[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              acc[2 * jj] = lo2(accA[jj]) + hi2(accB[jj]);
              acc[2 * jj + 1] = hi2(accA[jj]) + lo2(accB[jj + 1]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float v[24], gg[16];
              const float4* yq = reinterpret_cast<const float4*>(s_u + c * T3_UP + 8 * p - 8);
#pragma unroll
              for (int i = 0; i < 6; ++i) {
                const float4 u = yq[i];
                v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 gv = *reinterpret_cast<const float4*>(s_tab + c * 64 + r * 16 + 4 * i);
                gg[4 * i] = gv.x; gg[4 * i + 1] = gv.y; gg[4 * i + 2] = gv.z; gg[4 * i + 3] = gv.w;
              }
#pragma unroll
              for (int d = 0; d < 16; ++d)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(gg[d], v[1 + e + d], acc[e]);
            }
          }
        }
      }
      if (dbg) dbg[5] = clock64();
      __syncthreads();  // every FIR window has been read: U becomes the output staging buffer
      if (fir_on) {
        // row p = 32 floats (blocks 2p, 2p+1), 16-byte chunk e XOR-swizzled with p
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          float* o = s_out + 32 * p + rh + 2 * h2;
#pragma unroll
          for (int e = 0; e < 8; ++e) o[4 * (e ^ (p & 7))] = accs[h2][e];
        }
      }
    }
    __syncthreads();
    // coalesced store of the owned outputs: hop blocks [2, 2+nq) of the staging buffer = 4*nq chunks of 16 bytes
    {
      float4* dst = reinterpret_cast<float4*>(a.wav + (size_t)b * 16 * L + 16 * (size_t)Q0);
      const float4* src = reinterpret_cast<const float4*>(s_out);
      for (int c = tid; c < 4 * nq; c += T3_NF) {
        const int lc = c + 8, pr = lc >> 3, e = lc & 7;
        dst[c] = src[8 * pr + (e ^ (pr & 7))];
      }
    }
    if (dbg) dbg[6] = clock64();
    if (!has_next) {
      if (a.dbg && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.dbg[64 + 3 * blockIdx.x + 1] = (long long)t;
      }
      break;
    }
    g += nq;
    cur = nxt;  // (the barrier after the next phase A orders these staging reads before U is rewritten)
  }
}

static cudaError_t launch_tail_mb3(const TailArgs& a_in, int precise, int num_sms, cudaStream_t st) {
  TailArgs a = a_in;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    dbg_on = (getenv("MBV_TAIL_TIMELINE") && atoi(getenv("MBV_TAIL_TIMELINE"))) ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, (64 + 3 * 1024) * sizeof(long long)); cudaMemset(dbg, 0, (64 + 3 * 1024) * sizeof(long long)); }
  }
  a.dbg = dbg_on ? dbg : nullptr;
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tc_tensormap_encoder());
  if (!enc) return cudaErrorNotSupported;
  const int F = a.L + 1;
  CUtensorMap tm32, tm8;
  cuuint64_t dims[3] = {72, (cuuint64_t)F, (cuuint64_t)a.B};
  cuuint64_t strides[2] = {288, (cuuint64_t)F * 288};
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint32_t box32[3] = {32, (cuuint32_t)T3_NF, 1}, box8[3] = {8, (cuuint32_t)T3_NF, 1};
  if (enc(&tm32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.logits), dims, strides, box32, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  if (enc(&tm8, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.logits), dims, strides, box8, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tail_mb3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_mb3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_mb3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_mb3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const long long total = (long long)a.B * a.L;
  long long ctas = (total + T3_NQ - 1) / T3_NQ;
  if (ctas > (long long)T3_CTAS * num_sms) ctas = (long long)T3_CTAS * num_sms;
  const bool emit = a.spec != nullptr;
  if (precise) {
    if (emit) tail_mb3_kernel<true, true><<<(int)ctas, T3_NF, T3_SMEM, st>>>(tm32, tm8, a);
    else tail_mb3_kernel<true, false><<<(int)ctas, T3_NF, T3_SMEM, st>>>(tm32, tm8, a);
  } else {
    if (emit) tail_mb3_kernel<false, true><<<(int)ctas, T3_NF, T3_SMEM, st>>>(tm32, tm8, a);
    else tail_mb3_kernel<false, false><<<(int)ctas, T3_NF, T3_SMEM, st>>>(tm32, tm8, a);
  }
  if (dbg_on) {
    cudaStreamSynchronize(st);
    long long hb[64];
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tail timeline] CTA 0, thread 0: wait | phase A | barrier | B | C | stage+store  (SM clocks)\n");
    for (int i = 0; i < 8 && hb[8 * i]; ++i)
      fprintf(stderr, "  tile %d  start %7lld  wait %5lld  A %5lld  bar %5lld  issue %5lld  B %5lld  C %5lld  store %5lld\n", i, hb[8 * i] - hb[0],
              hb[8 * i + 1] - hb[8 * i], hb[8 * i + 2] - hb[8 * i + 1], hb[8 * i + 3] - hb[8 * i + 2], hb[8 * i + 7] - hb[8 * i + 3],
              hb[8 * i + 4] - hb[8 * i + 7], hb[8 * i + 5] - hb[8 * i + 4], hb[8 * i + 6] - hb[8 * i + 5]);
    {
      std::vector<long long> cb(3 * 1024);
      cudaMemcpy(cb.data(), dbg + 64, cb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      long long t0 = 0, t1 = 0;
      int n = 0;
      for (int c = 0; c < 1024; ++c) if (cb[3 * c]) { if (!t0 || cb[3 * c] < t0) t0 = cb[3 * c]; if (cb[3 * c + 1] > t1) t1 = cb[3 * c + 1]; ++n; }
      fprintf(stderr, "[tail timeline] %d CTAs, first start -> last end %.1f us\n", n, (t1 - t0) * 1e-3);
      for (int c = 0; c < n; c += 37)
        fprintf(stderr, "  CTA %3d sm %3lld  start +%6.1f us  end +%6.1f us\n", c, cb[3 * c + 2], (cb[3 * c] - t0) * 1e-3, (cb[3 * c + 1] - t0) * 1e-3);
    }
    cudaMemset(dbg, 0, (64 + 3 * 1024) * sizeof(long long));
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// conv_post INSIDE the tail (16-bit operand paths, MB / MS decoders):  models.py:363-375 as ONE kernel.
//
// The two-kernel path writes the fp32 logits [B][16T+1][72] (254 MB at BASELINE size) and reads them back.  Here the
// logits never leave the SM:
//   * conv_post (Conv1d C -> 72, k 7, on the reflect-padded lrelu_0.01 operand tensor the last ResBlock wrote) runs on
//     the tensor cores with the FRAME on the accumulator lane:  D[128 frames, 80] = sum_tap sum_kb X[frames + tap, 64] .
//     W[tap][kb][80, 64]^T  (tcgen05.mma M 128, N 80 = 72 channels + padding, K 16; the activation slab is the A
//     operand, tap = row offset of its descriptor).  A consumer thread then owns one frame and reads its 72 logits with
//     tcgen05.ld -- exactly the "lane = frame" layout the tail arithmetic wants, no transpose, no shared-memory tile.
//   * all conv_post weights (7 taps x C/64 k-blocks x 72 rows x 128 B = 126 KB for C = 128) stay resident in shared
//     memory for the life of the (persistent) CTA; the 8 padding rows of a tile are simply the first rows of the next one
//     (they only feed the 8 unused accumulator columns).
//   * warp roles: 1 TMA producer (activation slabs, one k-block per stage), 1 MMA issuer, TF_G consumer groups of 4
//     warps.  Each group walks its own equal share of the hop blocks (like one CTA of tail_mb3_kernel) with its own
//     double-buffered 80-column accumulator pair in TMEM, so head / iDFT / FIR of one group overlap the MMAs and the
//     tile of the others.  The tail arithmetic is the one of tail_mb3_kernel (same operation order).
// Algorithmic HBM bytes per latent frame: 16 frames x C x 2 B in (4096 for C = 128) + 1024 B of samples out.
// ------------------------------------------------------------------------------------------------
#ifndef MBV_TF_G
#define MBV_TF_G 3
#endif
constexpr int TF_G = MBV_TF_G;                 // consumer groups per CTA
constexpr int TF_THREADS = 128 * TF_G + 64;
constexpr int TF_NCOL = 80;                    // UMMA N
constexpr int TF_WROWS = 72;                   // weight rows kept per (tap, k-block) tile
constexpr int TF_WTILE = TF_WROWS * 128;       // 9216 B (a multiple of the 1024-byte swizzle atom)
constexpr int TF_TAPS = 7;
constexpr int TF_SLAB_ROWS = 136;              // 128 frames + 6 rows of tap halo, rounded up to 8
constexpr int TF_SLAB = TF_SLAB_ROWS * 128;
#ifndef MBV_TF_NSLAB
#define MBV_TF_NSLAB 2
#endif
constexpr int TF_NSLAB = MBV_TF_NSLAB;
constexpr int TF_OFF_X = TF_TAPS * 2 * TF_WTILE;                        // weights first (129024 B)
constexpr int TF_OFF_G = TF_OFF_X + TF_NSLAB * TF_SLAB;
constexpr int TF_GROUP_BYTES = 8 * T3_UP * 4 + T3_NW * 4 * 6 * 16;      // U + halo of one group
constexpr int TF_OFF_TAB = TF_OFF_G + TF_G * TF_GROUP_BYTES;
constexpr int TF_OFF_BAR = TF_OFF_TAB + 1024;
constexpr int TF_NBAR = 1 + 2 * TF_NSLAB + 4 * TF_G;
constexpr int TF_SMEM = TF_OFF_BAR + TF_NBAR * 8 + 16 + 1024;          // + slack for the 1024-byte alignment
static_assert(TF_SMEM <= 227 * 1024, "fused tail: shared memory budget");
static_assert(T3_NF == 128, "fused tail: one consumer group = 128 frames = 128 TMEM lanes");

struct FusedTailArgs {
  TailArgs t;          // outputs, filter tables, B, L (t.logits unused)
  float bias[72];      // conv_post bias (by value: read as constant-bank operands)
  int kblocks;         // input channels / 64 (1 or 2)
  int f16;             // operand element type: 0 bf16, 1 fp16
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// the hop-block walk of one consumer group (identical in the producer, the MMA issuer and the group itself)
struct TfWalk {
  int b, q0, left, L;
  __device__ __forceinline__ void init(long long total, int n_v, int v, int L_) {
    L = L_;
    const long long per = (total + n_v - 1) / n_v;
    const long long g0 = per * v;
    const long long g_end = (g0 + per < total) ? g0 + per : total;
    left = (int)(g_end > g0 ? g_end - g0 : 0);
    b = (int)(g0 / L); q0 = (int)(g0 % L);
  }
  __device__ __forceinline__ int nq() const {
    const int tiles_left = (left + T3_NQ - 1) / T3_NQ;
    int n = (left + tiles_left - 1) / tiles_left;
    if (n > L - q0) n = L - q0;
    return n;
  }
  __device__ __forceinline__ void advance(int n) {
    left -= n;
    if (q0 + n == L) { b += 1; q0 = 0; } else { q0 += n; }
  }
};

template <bool EMIT>
__global__ void __launch_bounds__(TF_THREADS, 1)
tail_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ FusedTailArgs fa) {
  constexpr int S = 4;
  constexpr bool PRECISE = false;
  const TailArgs& a = fa.t;
  extern __shared__ __align__(1024) uint8_t tf_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tf_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smW = sm;
  uint8_t* smX = sm + TF_OFF_X;
  float* s_tab = reinterpret_cast<float*>(sm + TF_OFF_TAB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TF_OFF_BAR);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + TF_NBAR);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int iWF = 0, iXF = 1, iXE = iXF + TF_NSLAB, iDF = iXE + TF_NSLAB, iDE = iDF + 2 * TF_G;
  const int tid_cta = threadIdx.x, warp_cta = tid_cta >> 5, lane = tid_cta & 31;
  const int L = a.L, F = L + 1;
  const int kblocks = fa.kblocks;
  const long long total = (long long)a.B * L;
  const int n_v = (int)gridDim.x * TF_G;

  if (a.fast_pqmf) { if (tid_cta < 128) s_tab[tid_cta] = a.g2[tid_cta >> 5][(tid_cta >> 1) & 15]; }
  else { for (int i = tid_cta; i < 256; i += TF_THREADS) s_tab[i] = a.coef[i >> 6][i & 63]; }
  if (tid_cta == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    mbar_init(BAR(iWF), 1);
    for (int i = 0; i < TF_NSLAB; ++i) { mbar_init(BAR(iXF + i), 1); mbar_init(BAR(iXE + i), 1); }
    for (int i = 0; i < 2 * TF_G; ++i) { mbar_init(BAR(iDF + i), 1); mbar_init(BAR(iDE + i), 4); }
    fence_barrier_init();
  }
  if (warp_cta == 4 * TF_G + 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp_cta == 4 * TF_G) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(BAR(iWF), (uint32_t)(TF_TAPS * kblocks * TF_WTILE));
      for (int tap = 0; tap < TF_TAPS; ++tap)
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_2d(smem_u32(smW + (size_t)(tap * kblocks + kb) * TF_WTILE), &tmW, BAR(iWF), kb * 64, tap * 128);
    }
    __syncwarp();
    TfWalk w[TF_G];
#pragma unroll
    for (int g = 0; g < TF_G; ++g) w[g].init(total, n_v, (int)blockIdx.x * TF_G + g, L);
    int sx = 0;
    uint32_t px = 0;
    bool any = true;
    while (any) {
      any = false;
#pragma unroll
      for (int g = 0; g < TF_G; ++g) {
        if (w[g].left <= 0) continue;
        any = true;
        const int nq = w[g].nq();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(BAR(iXE + sx), px ^ 1);
          if (elect_one()) {
            mbar_expect_tx(BAR(iXF + sx), (uint32_t)TF_SLAB);
            // frames [q0 - 3, +128) need the operand rows [q0 - 6, +134): rows outside [0, F) are conv_post's zero padding
            tma_load_3d(smem_u32(smX + (size_t)sx * TF_SLAB), &tmX, BAR(iXF + sx), kb * 64, w[g].q0 - 6, w[g].b);
          }
          __syncwarp();
          if (++sx == TF_NSLAB) { sx = 0; px ^= 1; }
        }
        w[g].advance(nq);
      }
    }
  } else if (warp_cta == 4 * TF_G + 1) {
    // ===================== MMA issuer =====================
    const uint32_t fmt = fa.f16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TF_NCOL >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    TfWalk w[TF_G];
    int n_done[TF_G];
#pragma unroll
    for (int g = 0; g < TF_G; ++g) { w[g].init(total, n_v, (int)blockIdx.x * TF_G + g, L); n_done[g] = 0; }
    mbar_wait(BAR(iWF), 0);
    tc_fence_after();
    int sx = 0;
    uint32_t px = 0;
    bool any = true;
    while (any) {
      any = false;
#pragma unroll
      for (int g = 0; g < TF_G; ++g) {
        if (w[g].left <= 0) continue;
        any = true;
        const int nq = w[g].nq();
        const int buf = n_done[g] & 1;
        mbar_wait(BAR(iDE + 2 * g + buf), (((uint32_t)n_done[g] >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)((2 * g + buf) * TF_NCOL);
        uint32_t accum = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(BAR(iXF + sx), px);
          tc_fence_after();
          const uint32_t x_lo = desc_lo(smem_u32(smX + (size_t)sx * TF_SLAB));
          if (elect_one()) {
            for (int tap = 0; tap < TF_TAPS; ++tap) {
              const uint32_t w_lo = desc_lo(smem_u32(smW + (size_t)(tap * kblocks + kb) * TF_WTILE));
              const uint32_t a_lo = x_lo + (uint32_t)tap * (TC_ROW_BYTES >> 4);   // tap = `tap` rows further into the slab
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                tc_mma<2>(tmem_d, desc64(a_lo + 2 * k), desc64(w_lo + 2 * k), idesc, accum);
                accum = 1;
              }
            }
            tc_commit(BAR(iXE + sx));
          }
          __syncwarp();
          accum = 1;
          if (++sx == TF_NSLAB) { sx = 0; px ^= 1; }
        }
        if (elect_one()) tc_commit(BAR(iDF + 2 * g + buf));
        __syncwarp();
        n_done[g]++;
        w[g].advance(nq);
      }
    }
  } else {
    // ===================== consumer groups: head + iDFT + overlap-add + synthesis FIR on TMEM-resident logits =====================
    const int grp = warp_cta >> 2;
    const int tid = tid_cta & 127, warp = warp_cta & 3;
    uint8_t* gsm = sm + TF_OFF_G + (size_t)grp * TF_GROUP_BYTES;
    float* s_u = reinterpret_cast<float*>(gsm);                      // [8][T3_UP]
    f2* s_halo = reinterpret_cast<f2*>(gsm + 8 * T3_UP * 4);          // [warp][band pair][6 parts][4 samples]
    float* s_out = s_u;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };
    TfWalk wk;
    wk.init(total, n_v, (int)blockIdx.x * TF_G + grp, L);
    int n_done = 0;
    group_sync();   // s_tab written
    while (wk.left > 0) {
      const int b = wk.b, Q0 = wk.q0, nq = wk.nq();
      const int QY0 = Q0 - 2, F0 = Q0 - 3;
      const bool last_tile = (Q0 + nq == L);
      const int buf = n_done & 1;
      mbar_wait(BAR(iDF + 2 * grp + buf), ((uint32_t)n_done >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((2 * grp + buf) * TF_NCOL);

      // ---- phase A (see tail_mb3_kernel): lane = frame F0 + tid = hop block QY0 + tid
      f2 yp[2][4];
      {
        const int f = F0 + tid;
        const bool live = (f >= 0) && (f < F);
        const bool emit = EMIT && live && (f >= Q0) && (f < Q0 + nq || (last_tile && f == L));
        const f2 keep = dup2(live ? 1.f : 0.f);
        const f2 k1 = dup2(lane < 31 ? 1.f : 0.f), k2 = dup2(lane < 30 ? 1.f : 0.f), k3 = dup2(lane < 29 ? 1.f : 0.f);
        auto band_pair = [&](auto p_tag) {
          constexpr int P = decltype(p_tag)::value, s0 = 2 * P;
          // logits [36P, 36P + 36) of this frame: band s0 then band s0 + 1 (9 magnitudes, 9 phases each)
          float x[36];
          tmem_ld32(taddr + 36 * P, x);
          tmem_ld4(taddr + 36 * P + 32, x + 32);
          tmem_ld_wait();
          if (P == 1) {  // the accumulator is in registers: hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(iDE + 2 * grp + buf));
          }
#pragma unroll
          for (int i = 0; i < 36; ++i) x[i] += fa.bias[36 * P + i];
          f2 re[9], im[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const float mag0 = ptx_ex2(x[k] * 1.4426950408889634f), mag1 = ptx_ex2(x[18 + k] * 1.4426950408889634f);
            const float sn0 = ptx_sin(x[9 + k]), sn1 = ptx_sin(x[27 + k]);
            const f2 ph = 3.14159265358979323846f * mk2(sn0, sn1);   // phase = pi * sin(x)  (models.py:369)
            const float ph0 = lo2(ph), ph1 = hi2(ph);
            const float s0v = ptx_sin(ph0), c0v = ptx_cos(ph0), s1v = ptx_sin(ph1), c1v = ptx_cos(ph1);
            if (EMIT) {
              if (emit) {
                const size_t o = (((size_t)b * S + s0) * 9 + k) * F + f;
                a.spec[o] = mag0; a.phase[o] = ph0;
                a.spec[o + (size_t)9 * F] = mag1; a.phase[o + (size_t)9 * F] = ph1;
              }
            }
            const f2 mag = keep * mk2(mag0, mag1);
            re[k] = mag * mk2(c0v, c1v);
            im[k] = mag * mk2(s0v, s1v);
          }
          f2 fr[16];
          idft16_windowed<f2>(re, im, fr);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f2 acc = fr[12 + i], v;
            v.v = __shfl_down_sync(0xffffffffu, fr[8 + i].v, 1); acc = fma2(k1, v, acc);
            v.v = __shfl_down_sync(0xffffffffu, fr[4 + i].v, 2); acc = fma2(k2, v, acc);
            v.v = __shfl_down_sync(0xffffffffu, fr[i].v, 3);     acc = fma2(k3, v, acc);
            yp[P][i] = acc;
          }
          {
            f2* h = s_halo + ((warp * 2 + P) * 6) * 4 + (lane == 0 ? 0 : (lane == 1 ? 3 : 5)) * 4;
#pragma unroll
            for (int part = 0; part < 3; ++part) {
              const bool on = (warp > 0) && (lane + part < 3);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (on) h[part * 4 + i] = fr[4 * part + i];
            }
          }
        };
        band_pair(std::integral_constant<int, 0>{});
        band_pair(std::integral_constant<int, 1>{});
      }
      group_sync();  // halo visible; the previous tile's staged outputs have been stored (U is free)

      // ---- phase B: blocks that straddle a warp boundary, envelope, optional o_mb, modulation -> U
      {
        const int q = QY0 + tid;
        const bool inside = (q >= 0) && (q < L) && (tid < T3_NF - 3);
#pragma unroll
        for (int P = 0; P < 2; ++P) {
          {
            const bool fix = (warp < T3_NW - 1) && (lane >= 29);
            const bool on1 = fix && (lane >= 30), on2 = fix && (lane == 31);
            const int s0 = lane == 30 ? 1 : (lane == 31 ? 2 : 0), s1 = lane == 30 ? 3 : 4;
            const f2* h = s_halo + ((((fix ? warp + 1 : warp)) * 2 + P) * 6) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const f2 a0 = h[s0 * 4 + i], a1 = h[s1 * 4 + i], a2 = h[5 * 4 + i];
              f2 v = yp[P][i];
              const f2 v0 = v + a0;
              if (fix) v = v0;
              const f2 v1 = v + a1;
              if (on1) v = v1;
              const f2 v2 = v + a2;
              if (on2) v = v2;
              yp[P][i] = v;
            }
          }
          if (!inside) {
#pragma unroll
            for (int i = 0; i < 4; ++i) yp[P][i] = dup2(0.f);
          } else if (PRECISE || q == 0 || q == L - 1) {
            float e[4] = {1.5f, 1.5f, 1.5f, 1.5f};
            if (q == 0) { e[0] -= win_sq(12); e[1] -= win_sq(13); e[2] -= win_sq(14); e[3] -= win_sq(15); }
            if (q == L - 1) { e[0] -= win_sq(0); e[1] -= win_sq(1); e[2] -= win_sq(2); e[3] -= win_sq(3); }
#pragma unroll
            for (int i = 0; i < 4; ++i) yp[P][i] = mk2(lo2(yp[P][i]) / e[i], hi2(yp[P][i]) / e[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) yp[P][i] = 0.66666666666666667f * yp[P][i];
          }
          if (a.o_mb != nullptr && inside && tid >= 2 && tid < 2 + nq) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              const int s = 2 * P + hb;
              const float4 v = hb ? make_float4(hi2(yp[P][0]), hi2(yp[P][1]), hi2(yp[P][2]), hi2(yp[P][3]))
                                  : make_float4(lo2(yp[P][0]), lo2(yp[P][1]), lo2(yp[P][2]), lo2(yp[P][3]));
              if (a.variant == 1) {
                *reinterpret_cast<float4*>(a.o_mb + ((size_t)b * S + s) * 4 * L + 4 * (size_t)q) = v;
              } else {
                float4* o = reinterpret_cast<float4*>(a.o_mb + ((size_t)b * S + s) * 16 * L + 16 * (size_t)q);
                o[0] = make_float4(4.f * v.x, 0.f, 0.f, 0.f);
                o[1] = make_float4(4.f * v.y, 0.f, 0.f, 0.f);
                o[2] = make_float4(4.f * v.z, 0.f, 0.f, 0.f);
                o[3] = make_float4(4.f * v.w, 0.f, 0.f, 0.f);
              }
            }
          }
        }
        if (a.fast_pqmf) {
          const unsigned long long* modp = reinterpret_cast<const unsigned long long*>(&a.mod[0][0]);
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            f2 m01, m23;
            m01.v = modp[2 * m]; m23.v = modp[2 * m + 1];
            float u[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const f2 t = fma2(m23, yp[1][i], m01 * yp[0][i]);
              u[i] = lo2(t) + hi2(t);
            }
            *reinterpret_cast<float4*>(s_u + m * T3_UP + 4 * tid) = make_float4(u[0], u[1], u[2], u[3]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int P = c >> 1;
            const float4 v = (c & 1) ? make_float4(hi2(yp[P][0]), hi2(yp[P][1]), hi2(yp[P][2]), hi2(yp[P][3]))
                                     : make_float4(lo2(yp[P][0]), lo2(yp[P][1]), lo2(yp[P][2]), lo2(yp[P][3]));
            *reinterpret_cast<float4*>(s_u + c * T3_UP + 4 * tid) = v;
          }
        }
      }
      group_sync();

      // ---- phase C: synthesis FIR (thread = hop blocks 2p, 2p+1; residues rh and rh + 2)
      {
        const int p = 16 * warp + (lane >> 1), rh = lane & 1;
        const bool fir_on = (p >= 1 && 2 * p < 2 + nq);
        float accs[2][8];
        if (fir_on) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int r = rh + 2 * h2;
            float* acc = accs[h2];
            if (a.fast_pqmf) {
              f2 we[12], wo[12];
              const ulonglong2* pe = reinterpret_cast<const ulonglong2*>(s_u + (7 - r) * T3_UP + 8 * p - 8);
              const ulonglong2* po = reinterpret_cast<const ulonglong2*>(s_u + (3 - r) * T3_UP + 8 * p - 8);
#pragma unroll
              for (int i = 0; i < 6; ++i) {
                const ulonglong2 u = pe[i], v = po[i];
                we[2 * i].v = u.x; we[2 * i + 1].v = u.y;
                wo[2 * i].v = v.x; wo[2 * i + 1].v = v.y;
              }
              const ulonglong2* gp = reinterpret_cast<const ulonglong2*>(s_tab) + r * 8;
              f2 accA[4], accB[5];
#pragma unroll
              for (int i = 0; i < 4; ++i) accA[i] = dup2(0.f);
#pragma unroll
              for (int i = 0; i < 5; ++i) accB[i] = dup2(0.f);
#pragma unroll
              for (int dd = 0; dd < 8; ++dd) {
                const ulonglong2 gv = gp[dd];
                f2 ge, go;
                ge.v = gv.x;
                go.v = gv.y;
#pragma unroll
                for (int jj = 0; jj < 5; ++jj) accB[jj] = fma2(ge, wo[jj + dd], accB[jj]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) accA[jj] = fma2(go, we[jj + dd + 1], accA[jj]);
              }
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                acc[2 * jj] = lo2(accA[jj]) + hi2(accB[jj]);
                acc[2 * jj + 1] = hi2(accA[jj]) + lo2(accB[jj + 1]);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float v[24], gg[16];
                const float4* yq = reinterpret_cast<const float4*>(s_u + c * T3_UP + 8 * p - 8);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                  const float4 u = yq[i];
                  v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 gv = *reinterpret_cast<const float4*>(s_tab + c * 64 + r * 16 + 4 * i);
                  gg[4 * i] = gv.x; gg[4 * i + 1] = gv.y; gg[4 * i + 2] = gv.z; gg[4 * i + 3] = gv.w;
                }
#pragma unroll
                for (int d = 0; d < 16; ++d)
#pragma unroll
                  for (int e = 0; e < 8; ++e) acc[e] = fmaf(gg[d], v[1 + e + d], acc[e]);
              }
            }
          }
        }
        group_sync();  // every FIR window has been read: U becomes the output staging buffer
        if (fir_on) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float* o = s_out + 32 * p + rh + 2 * h2;
#pragma unroll
            for (int e = 0; e < 8; ++e) o[4 * (e ^ (p & 7))] = accs[h2][e];
          }
        }
      }
      group_sync();
      {
        float4* dst = reinterpret_cast<float4*>(a.wav + (size_t)b * 16 * L + 16 * (size_t)Q0);
        const float4* src = reinterpret_cast<const float4*>(s_out);
        for (int c = tid; c < 4 * nq; c += T3_NF) {
          const int lc = c + 8, pr = lc >> 3, e = lc & 7;
          dst[c] = src[8 * pr + (e ^ (pr & 7))];
        }
      }
      n_done++;
      wk.advance(nq);
      // (the group barrier after the next phase A orders these staging reads before U is rewritten)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_cta == 4 * TF_G + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

cudaError_t launch_tail_fused(const TailArgs& t, const void* act, const void* w, const float* bias, int C, int f16, int num_sms,
                              cudaStream_t st) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tc_tensormap_encoder());
  if (!enc) return cudaErrorNotSupported;
  if (C != 64 && C != 128) return cudaErrorInvalidValue;
  FusedTailArgs fa;
  fa.t = t;
  fa.t.dbg = nullptr;
  memcpy(fa.bias, bias, sizeof(fa.bias));
  fa.kblocks = C / 64;
  fa.f16 = f16;
  const int F = t.L + 1;
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)F, (cuuint64_t)t.B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)F * C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)TF_SLAB_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tmX, dt, 3, const_cast<void*>(act), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)TF_TAPS * 128};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)TF_WROWS};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tmW, dt, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tail_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const long long total = (long long)t.B * t.L;
  long long ctas = (total + (long long)TF_G * T3_NQ - 1) / ((long long)TF_G * T3_NQ);
  if (ctas > num_sms) ctas = num_sms;
  if (ctas < 1) ctas = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(TF_THREADS);
  cfg.dynamicSmemBytes = TF_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (t.spec != nullptr) return cudaLaunchKernelEx(&cfg, tail_fused_kernel<true>, tmX, tmW, fa);
  return cudaLaunchKernelEx(&cfg, tail_fused_kernel<false>, tmX, tmW, fa);
}

// ------------------------------------------------------------------------------------------------
// Single-band tail, v3: lane = STFT frame, overlap-add by warp shuffles, hop blocks stored straight from registers.
// Each WARP covers 32 consecutive frames = 29 complete hop blocks and the warps of a CTA overlap by 3 frames, so there
// is no cross-warp exchange and only ONE block barrier (after the coalesced copy of the logits rows); the 10 % of
// repeated head work buys the removal of the frame scratch round trip and the second barrier of the round-1
// kernel (96 us on the BASELINE-size problem).
// ------------------------------------------------------------------------------------------------
constexpr int SB_THREADS = 256;
constexpr int SB_NQ = 29 * (SB_THREADS / 32);   // 232 owned hop blocks per CTA
constexpr int SB_ROWS = SB_NQ + 3;              // 235 frames held in shared memory

template <bool PRECISE>
__global__ void __launch_bounds__(SB_THREADS) tail_sb3_kernel(const __grid_constant__ TailArgs a, int tiles_per_utt) {
  __shared__ __align__(16) float s_log[SB_ROWS * 18 + 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / tiles_per_utt, tile = blockIdx.x % tiles_per_utt;
  const int L = a.L, F = L + 1;
  const int Q0 = tile * SB_NQ, F0 = Q0 - 1;
  const int nq = min(SB_NQ, L - Q0);
  const bool last_tile = (Q0 + nq == L);
  // ---- coalesced copy of the logits rows [F0, F0 + 235) /\ [0, F); rows outside the utterance are zero-filled
  {
    const int f_lo = max(F0, 0), f_hi = min(F0 + SB_ROWS, F);
    const float* src = a.logits + ((size_t)b * F + f_lo) * 18;
    float* dst = s_log + (f_lo - F0) * 18;
    const int n = (f_hi - f_lo) * 18;
    for (int i = tid; i < (f_lo - F0) * 18; i += SB_THREADS) s_log[i] = 0.f;
    for (int i = (f_hi - F0) * 18 + tid; i < SB_ROWS * 18; i += SB_THREADS) s_log[i] = 0.f;
    if ((((size_t)b * F + f_lo) * 18 & 3) == 0 && (((f_lo - F0) * 18) & 3) == 0) {
      const int n4 = n >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = tid; i < n4; i += SB_THREADS) d4[i] = __ldg(s4 + i);
      for (int i = (n4 << 2) + tid; i < n; i += SB_THREADS) dst[i] = __ldg(src + i);
    } else {
      for (int i = tid; i < n; i += SB_THREADS) dst[i] = __ldg(src + i);
    }
  }
  __syncthreads();
  // ---- head + inverse DFT + window for frame F0 + r, r = 29 * warp + lane
  const int r = 29 * warp + lane;
  const int f = F0 + r;
  const bool live = (f >= 0) && (f < F);
  float fr[16];
  {
    const float2* lp = reinterpret_cast<const float2*>(s_log + r * 18);  // stride 18 words: conflict-free 8-byte loads
    float x[18];
#pragma unroll
    for (int i = 0; i < 9; ++i) { const float2 v = lp[i]; x[2 * i] = v.x; x[2 * i + 1] = v.y; }
    // this lane emits spec / phase of its frame if it is the frame's owner (rows 29w..29w+28; the last warp also its tail rows)
    const bool owner = (lane < 29) || (warp == SB_THREADS / 32 - 1);
    const bool emit = (a.spec != nullptr) && live && owner && (f >= Q0) && (f < Q0 + nq || (last_tile && f == L));
    float re[9], im[9];
    const float keep = live ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float mag, ph;
      head<PRECISE>(x[k], x[9 + k], mag, ph, re[k], im[k]);
      if (emit) {
        const size_t o = ((size_t)b * 9 + k) * F + f;
        a.spec[o] = mag;
        a.phase[o] = ph;
      }
      re[k] *= keep; im[k] *= keep;
    }
    idft16_windowed<float>(re, im, fr);
  }
  // ---- overlap-add: hop block q = f + 1 = part 3 of this frame + part 2 / 1 / 0 of the next three frames
  float4 y = make_float4(fr[12], fr[13], fr[14], fr[15]);
  add4(y, shfl_down4(fr + 8, 1));
  add4(y, shfl_down4(fr + 4, 2));
  add4(y, shfl_down4(fr, 3));
  const int q = Q0 + r;
  if (lane < 29 && r < nq) {
    if (PRECISE || q == 0 || q == L - 1) {
      float e0 = 1.5f, e1 = 1.5f, e2 = 1.5f, e3 = 1.5f;
      if (q == 0) { e0 -= win_sq(12); e1 -= win_sq(13); e2 -= win_sq(14); e3 -= win_sq(15); }
      if (q == L - 1) { e0 -= win_sq(0); e1 -= win_sq(1); e2 -= win_sq(2); e3 -= win_sq(3); }
      y.x /= e0; y.y /= e1; y.z /= e2; y.w /= e3;
    } else {
      const float inv = 0.66666666666666667f;
      y.x *= inv; y.y *= inv; y.z *= inv; y.w *= inv;
    }
    *reinterpret_cast<float4*>(a.wav + (size_t)b * 4 * L + 4 * (size_t)q) = y;
  }
}

static cudaError_t launch_tail_sb3(const TailArgs& a, int precise, cudaStream_t st) {
  const int tiles = (a.L + SB_NQ - 1) / SB_NQ;
  if (precise) tail_sb3_kernel<true><<<a.B * tiles, SB_THREADS, 0, st>>>(a, tiles);
  else tail_sb3_kernel<false><<<a.B * tiles, SB_THREADS, 0, st>>>(a, tiles);
  return cudaGetLastError();
}

cudaError_t launch_tail(const TailArgs& a, int precise, int num_sms, cudaStream_t st) {
  if (a.variant == 0) return launch_tail_sb3(a, precise, st);
  return launch_tail_mb3(a, precise, num_sms, st);
}

}  // namespace mbv
