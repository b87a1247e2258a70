// ENet stage-5 regular bottleneck (regular5_1: 16 channels, internal width 4, 128x256 pixels)
// as ONE CUDA-core kernel.  With K = 4 the three convolutions are far too skinny for tcgen05 and
// the block is pure bandwidth: x is read once (+ a one-pixel halo) and y written once, instead of
// the six tensor passes of the unfused 1x1 -> 3x3 -> 1x1(+residual) chain.
//
//   e1 = PReLU(W1 x + b1)               on the tile + halo, zero outside the image  -> smem
//   e2 = PReLU(W2 (*) e1 + b2)          3x3, padding 1
//   y  = PReLU_out(PReLU(W3 e2 + b3) + x)
//
// Storage-rounding points (rnd<T>) are those of the unfused chain, so BC_PREC_BF16 results match
// the bf16-emulating oracle and BC_PREC_FP32 is exact.  Semantics: oracle/enet_oracle.py `regular`
// (the frozen graph the reference runs, models.py:43-44).
#include "simt_common.cuh"

#include <cstring>

namespace bc {

static constexpr int S5_TW = 32, S5_TH = 8;                     // pixels per block: 32 x 8, 2 per thread
static constexpr int S5_HW = S5_TW + 2, S5_HH = S5_TH + 2;      // with halo

// All 336 weights / biases / slopes travel BY VALUE as kernel parameters: with every loop unrolled
// their indices are compile-time constants, so each one is a constant-bank operand of its FFMA
// (the kernel was shared-memory-pipe bound, 93 % l1tex, when it fetched them with LDS).

// Packed fp32 FMA (FFMA2, sm_100): two independent round-to-nearest FMAs per instruction -- bit-identical to the
// scalar form, half the issue slots.  The kernel is issue-bound (95 %), and 43 % of its instructions were FFMAs.
__device__ __forceinline__ float2 fma2(float s, const float* w2, float2 acc) {      // acc + s * (w2[0], w2[1])
  return __ffma2_rn(make_float2(s, s), *reinterpret_cast<const float2*>(w2), acc);
}

template <typename T>
__global__ void __launch_bounds__(128)
k_stage5_bottleneck(const T* __restrict__ x, T* __restrict__ y, const __grid_constant__ Stage5Params p, int H, int W,
                    int reverse) {
  __shared__ float4 se1[S5_HH * S5_HW];        // e1 of the tile + halo, 4 channels per pixel
  const int tid = threadIdx.x;
  // reverse: blocks are dispatched in (x, y, z) order, so mirroring all three walks the batch back to front
  const int bx = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x, by = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y,
            bz = reverse ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
  const int x0 = bx * S5_TW, y0 = by * S5_TH;
  const T* xf = x + (size_t)bz * H * W * 16;
  T* yf = y + (size_t)bz * H * W * 16;

  // ---- phase 1: projection 16 -> 4 on the halo tile
  for (int i = tid; i < S5_HH * S5_HW; i += 128) {
    const int hy = i / S5_HW, hx = i % S5_HW;
    const int gy = y0 + hy - 1, gx = x0 + hx - 1;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);                // zero padding of the 3x3 conv
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      float v[16];
      ld_ch<16>(xf + ((size_t)gy * W + gx) * 16, v);
      float2 a01 = make_float2(p.b1[0], p.b1[1]), a23 = make_float2(p.b1[2], p.b1[3]);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        a01 = fma2(v[k], &p.w1[4 * k], a01);
        a23 = fma2(v[k], &p.w1[4 * k + 2], a23);
      }
      const float a[4] = {a01.x, a01.y, a23.x, a23.y};
      e = make_float4(rnd<T>(prelu(a[0], p.a1[0])), rnd<T>(prelu(a[1], p.a1[1])), rnd<T>(prelu(a[2], p.a1[2])),
                      rnd<T>(prelu(a[3], p.a1[3])));
    }
    se1[i] = e;
  }
  __syncthreads();

  // ---- phase 2: 3x3 conv 4 -> 4, expansion 4 -> 16, residual, two pixels per thread
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int ly = (tid >> 5) + 4 * r, lx = tid & 31;          // a warp covers one tile row: coalesced
    const int gy = y0 + ly, gx = x0 + lx;
    if (gy >= H || gx >= W) continue;
    float2 a01 = make_float2(p.b2[0], p.b2[1]), a23 = make_float2(p.b2[2], p.b2[3]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 e = se1[(ly + t / 3) * S5_HW + lx + t % 3];
      const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a01 = fma2(ev[k], &p.w2[(t * 4 + k) * 4], a01);
        a23 = fma2(ev[k], &p.w2[(t * 4 + k) * 4 + 2], a23);
      }
    }
    const float a[4] = {a01.x, a01.y, a23.x, a23.y};
    float e2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) e2[j] = rnd<T>(prelu(a[j], p.a2[j]));
    float o[16], xr[16];
    ld_ch<16>(xf + ((size_t)gy * W + gx) * 16, xr);
    float2 o2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o2[j] = make_float2(p.b3[2 * j], p.b3[2 * j + 1]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) o2[j] = fma2(e2[k], &p.w3[k * 16 + 2 * j], o2[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[2 * j] = o2[j].x; o[2 * j + 1] = o2[j].y; }
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = prelu(prelu(o[j], p.a3[j]) + xr[j], p.aout[j]);
    st_ch<16>(yf + ((size_t)gy * W + gx) * 16, o);
  }
}

template <typename T>
void launch_stage5(const T* x, T* y, const Bottleneck& b, int B, int H, int W, cudaStream_t s) {
  if (b.s5.size() != sizeof(Stage5Params) / sizeof(float)) return;      // built by the loader (api.cu)
  Stage5Params p;
  memcpy(&p, b.s5.data(), sizeof p);
  dim3 grid((W + S5_TW - 1) / S5_TW, (H + S5_TH - 1) / S5_TH, B);
  k_stage5_bottleneck<T><<<grid, 128, 0, s>>>(x, y, p, H, W, g_umma_reverse);
}
template void launch_stage5<float>(const float*, float*, const Bottleneck&, int, int, int, cudaStream_t);
template void launch_stage5<bf16>(const bf16*, bf16*, const Bottleneck&, int, int, int, cudaStream_t);
template void launch_stage5<f16>(const f16*, f16*, const Bottleneck&, int, int, int, cudaStream_t);

}  // namespace bc
