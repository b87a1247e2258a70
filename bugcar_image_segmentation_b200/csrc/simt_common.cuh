// Shared device helpers of the CUDA-core ENet kernels (NHWC activations, fp32 accumulate).
#pragma once
#include "internal.h"

namespace bc {

// ---------------------------------------------------------------- vector load / store
template <int N>
__device__ __forceinline__ void ld_ch(const float* __restrict__ p, float (&v)[N]) {
  static_assert(N % 4 == 0, "N%4");
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    float4 t = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
template <int N>
__device__ __forceinline__ void ld_ch(const bf16* __restrict__ p, float (&v)[N]) {
  static_assert(N % 4 == 0, "N%4");
  if constexpr (N % 8 == 0) {
#pragma unroll
    for (int i = 0; i < N / 8; ++i) {
      uint4 t = reinterpret_cast<const uint4*>(p)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f = __bfloat1622float2(h[k]);
        v[8 * i + 2 * k] = f.x; v[8 * i + 2 * k + 1] = f.y;
      }
    }
  } else {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
    float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
  }
}
template <int N>
__device__ __forceinline__ void ld_ch(const f16* __restrict__ p, float (&v)[N]) {
  static_assert(N % 4 == 0, "N%4");
  if constexpr (N % 8 == 0) {
#pragma unroll
    for (int i = 0; i < N / 8; ++i) {
      uint4 t = reinterpret_cast<const uint4*>(p)[i];
      const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f = __half22float2(h[k]);
        v[8 * i + 2 * k] = f.x; v[8 * i + 2 * k + 1] = f.y;
      }
    }
  } else {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&t);
    float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
  }
}
template <int N>
__device__ __forceinline__ void st_ch(float* __restrict__ p, const float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N / 4; ++i)
    reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <int N>
__device__ __forceinline__ void st_ch(bf16* __restrict__ p, const float (&v)[N]) {
  if constexpr (N % 8 == 0) {
#pragma unroll
    for (int i = 0; i < N / 8; ++i) {
      uint4 t;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
      reinterpret_cast<uint4*>(p)[i] = t;
    }
  } else {
    uint2 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
}
// fp16 stores saturate at +-65504 (cvt.rn.satfinite) instead of overflowing to infinity, as in the tcgen05 kernels
__device__ __forceinline__ __half2 f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return *reinterpret_cast<__half2*>(&r);
}
template <int N>
__device__ __forceinline__ void st_ch(f16* __restrict__ p, const float (&v)[N]) {
  if constexpr (N % 8 == 0) {
#pragma unroll
    for (int i = 0; i < N / 8; ++i) {
      uint4 t;
      __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = f16x2_sat(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
      reinterpret_cast<uint4*>(p)[i] = t;
    }
  } else {
    uint2 t;
    __half2* h = reinterpret_cast<__half2*>(&t);
    h[0] = f16x2_sat(v[0], v[1]);
    h[1] = f16x2_sat(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
}
// value as it will be re-read from storage (16-bit rounding point)
template <typename T> __device__ __forceinline__ float rnd(float v);
template <> __device__ __forceinline__ float rnd<float>(float v) { return v; }
template <> __device__ __forceinline__ float rnd<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
template <> __device__ __forceinline__ float rnd<f16>(float v) { return __low2float(f16x2_sat(v, 0.f)); }

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : a * v; }

}  // namespace bc
