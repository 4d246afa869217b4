// Microbenchmark: is the channel-on-lane epilogue's store rate bound by the number of global-store instructions, and
// does trading two 2-byte stores for one shuffle + one 4-byte store (two channels per store) help?
//   variant 0: per thread 32 x st.global.u16 (lane = channel, rows 256 B apart)       -- what conv_tc's epilogues do
//   variant 1: per thread 16 x (shfl.xor 1 + st.global.u32): even lanes write even rows, odd lanes odd rows
// 12 warps per CTA, 1 CTA per SM (like the epilogue), each warp streams 32x32 blocks.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stg_shfl stg_shfl.cu && ./stg_shfl
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

template <int V>
__global__ void __launch_bounds__(384, 1) k(__nv_bfloat16* out, int C, int rows_per_warp, float seed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = warp & 3, g = warp >> 2;
  const size_t row0 = ((size_t)blockIdx.x * 3 + g) * rows_per_warp;
  __nv_bfloat16* base = out + row0 * C + q * 32 + lane;
  for (int r = 0; r < rows_per_warp; r += 32) {
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = seed * (float)(i + lane) + (float)r;
    if (V == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) base[(size_t)(r + i) * C] = __float2bfloat16_rn(y[i]);
    } else {
      const bool odd = lane & 1;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float mine = odd ? y[2 * j + 1] : y[2 * j];
        const float send = odd ? y[2 * j] : y[2 * j + 1];
        const float got = __shfl_xor_sync(0xffffffffu, send, 1);
        const __nv_bfloat162 p = odd ? __floats2bfloat162_rn(got, mine) : __floats2bfloat162_rn(mine, got);
        __nv_bfloat16* dst = base + (size_t)(r + 2 * j + (odd ? 1 : 0)) * C - (odd ? 1 : 0);
        *reinterpret_cast<__nv_bfloat162*>(dst) = p;
      }
    }
  }
}

int main() {
  const int C = 128, rows_per_warp = 4096, ctas = 148;
  const size_t n = (size_t)ctas * 3 * rows_per_warp * C;
  __nv_bfloat16* d;
  cudaMalloc(&d, n * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int v = 0; v < 2; ++v) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (v == 0) k<0><<<ctas, 384>>>(d, C, rows_per_warp, 1.f);
      else k<1><<<ctas, 384>>>(d, C, rows_per_warp, 1.f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("variant %d: %.3f ms  %.0f GB/s  %.2f clk/store-instr/SM @1.9GHz\n", v, ms, n * 2 / ms / 1e6,
                           ms * 1e-3 * 1.9e9 / ((double)rows_per_warp * 12 / (v ? 2 : 1)));
    }
  }
  // checksum so both variants can be compared
  return 0;
}
