// conv_simt.cu -- CUDA-core fp32-accumulate implicit-GEMM conv (MBV_PREC_FP32 path and the
// operand-identical cross-check of the tcgen05 kernel).  Same ConvArgs / packed weights / epilogues as
// conv_tc.cu; one CTA computes 64 rows x 64 logical columns, one thread 1 row x 16 columns.
#include "common.cuh"
#include "kernels.h"

namespace mbv {

constexpr int SIMT_ROWS = 64;
constexpr int SIMT_COLS = 64;
constexpr int SIMT_KC = 16;

template <typename Op>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvArgs a) {
  using T = typename Op::T;
  __shared__ float xs[SIMT_KC][SIMT_ROWS + 1];
  __shared__ __align__(16) float ws[SIMT_KC][SIMT_COLS];
  __shared__ __align__(16) float ws2[SIMT_KC][SIMT_COLS];

  const int tid = threadIdx.x;
  const int r = tid % SIMT_ROWS;
  const int cg = tid / SIMT_ROWS;
  const int t0 = blockIdx.x * SIMT_ROWS;
  const int nl0 = blockIdx.y * SIMT_COLS;  // logical column base
  const int b = blockIdx.z / a.n_phases;
  const int phase = blockIdx.z % a.n_phases;
  const int NL = a.gate ? a.N_total / 2 : a.N_total;  // logical columns

  float acc[16], acc2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = 0.f; acc2[i] = 0.f; }

  const T* x = reinterpret_cast<const T*>(a.x) + (size_t)b * a.L_in * a.x_ld;
  const T* w = reinterpret_cast<const T*>(a.w) + (size_t)phase * a.taps * a.N_total * a.Cp_in;

  for (int tap = 0; tap < a.taps; ++tap) {
    const int shift = a.shift0[phase] + tap * a.dil;
    const T* wt = w + (size_t)tap * a.N_total * a.Cp_in;
    for (int c0 = 0; c0 < a.Cp_in; c0 += SIMT_KC) {
      __syncthreads();
      for (int e = tid; e < SIMT_ROWS * SIMT_KC; e += 256) {
        const int rr = e / SIMT_KC, cc = e % SIMT_KC;
        const int t = t0 + rr + shift;
        float v = 0.f;
        if (t >= 0 && t < a.L_in) v = op_load<Op>(x + (size_t)t * a.x_ld + c0 + cc);
        xs[cc][rr] = v;
      }
      for (int e = tid; e < SIMT_COLS * SIMT_KC; e += 256) {
        const int nn = e / SIMT_KC, cc = e % SIMT_KC;
        const int n = nl0 + nn;
        float v = 0.f, v2 = 0.f;
        if (n < NL) {
          // gate packing: 128-row tiles of [64 tanh rows | 64 sigmoid rows] for 64 consecutive channels
          const int row = a.gate ? 128 * (n >> 6) + (n & 63) : n;
          v = op_load<Op>(wt + (size_t)row * a.Cp_in + c0 + cc);
          if (a.gate) v2 = op_load<Op>(wt + (size_t)(row + 64) * a.Cp_in + c0 + cc);
        }
        ws[cc][nn] = v;
        ws2[cc][nn] = v2;
      }
      __syncthreads();
#pragma unroll
      for (int cc = 0; cc < SIMT_KC; ++cc) {
        const float xv = xs[cc][r];
        const float4* wp = reinterpret_cast<const float4*>(&ws[cc][cg * 16]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 f = wp[q];
          acc[4 * q + 0] = fmaf(xv, f.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(xv, f.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, f.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, f.w, acc[4 * q + 3]);
        }
        if (a.gate) {
          const float4* wp2 = reinterpret_cast<const float4*>(&ws2[cc][cg * 16]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 f = wp2[q];
            acc2[4 * q + 0] = fmaf(xv, f.x, acc2[4 * q + 0]);
            acc2[4 * q + 1] = fmaf(xv, f.y, acc2[4 * q + 1]);
            acc2[4 * q + 2] = fmaf(xv, f.z, acc2[4 * q + 2]);
            acc2[4 * q + 3] = fmaf(xv, f.w, acc2[4 * q + 3]);
          }
        }
      }
    }
  }
  const int row = t0 + r;
  const int n0 = nl0 + cg * 16;
  if (row < a.L_out && n0 < NL) epilogue_chunk<Op, 16>(a.epi, b, row, phase, n0, acc, acc2);
}

template <typename Op>
static cudaError_t launch_simt(const ConvArgs& a, cudaStream_t st) {
  const int NL = a.gate ? a.N_total / 2 : a.N_total;
  dim3 grid((a.L_out + SIMT_ROWS - 1) / SIMT_ROWS, (NL + SIMT_COLS - 1) / SIMT_COLS, a.B * a.n_phases);
  conv_simt_kernel<Op><<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_conv_simt(int prec, const ConvArgs& a, cudaStream_t st) {
  switch (prec) {
    case 0: return launch_simt<OpF32>(a, st);
    case 1: return launch_simt<OpTF32>(a, st);
    case 3: return launch_simt<OpF16>(a, st);
    default: return launch_simt<OpBF16>(a, st);
  }
}

}  // namespace mbv
