// Shared device/host helpers for the objgan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define OG_API extern "C" __attribute__((visibility("default")))

// Every launcher returns 0 on success, else the cudaError_t of the failed launch.
// (The reference's launcher calls exit(-1) on failure -- roi_align_kernel.cu:84-88; we never do.)
#define OG_RETURN_LAST_ERROR()                      \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    return (int)e__;                                \
  } while (0)

#define OG_CHECK(call)                              \
  do {                                              \
    cudaError_t e__ = (call);                       \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)

static inline int og_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float og_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Per-tensor power-of-two scaling of the fp16 tensor-core operands: amax_bits = float bits of max|x| over the tensor;
// the operand stores x * 2^k with k chosen so that max|x| * 2^k lies in [2^13, 2^14) (fp16 overflows at 65504).
__host__ __device__ __forceinline__ int og_scale_exp(unsigned amax_bits) {
  amax_bits &= 0x7fffffffu;
  if (amax_bits == 0) return 0;
  int k = 13 - ((int)(amax_bits >> 23) - 127);
  return k < -110 ? -110 : (k > 110 ? 110 : k);
}
__device__ __forceinline__ float og_exp2i(int k) { return __int_as_float((k + 127) << 23); }   // |k| <= 126

// activation codes shared by conv epilogues and the act-backward kernel
enum { OG_ACT_NONE = 0, OG_ACT_LRELU = 1, OG_ACT_TANH = 2, OG_ACT_SIGMOID = 3 };
// normalisation-apply fusions
enum { OG_NA_NONE = 0, OG_NA_LRELU = 1, OG_NA_GLU = 2 };
// conv input addressing modes
enum { OG_PAD_ZERO = 0, OG_PAD_REFLECT = 1, OG_UPSAMPLE2X = 2, OG_TRANSPOSED = 3 };
