// ENet forward pass, CUDA-core kernels (NHWC activations, fp32 accumulate).
//
// These kernels are (a) the whole network in BC_PREC_FP32 (exact mode), and (b) in
// BC_PREC_BF16 the layers that are not GEMM-shaped enough for tcgen05: the initial block
// (K=27, N=13), the down/up-sampling bottlenecks' pooling / unpooling / strided parts,
// stage-5 internals (4 channels) and the 16->C transposed-conv head fused with the class
// argmax + LUT.  The regular/dilated/asymmetric bottlenecks of stages 1-4 run through
// enet_umma.cu when tensor cores are enabled, and through k_conv here otherwise.
//
// The network executed by the reference is a frozen graph absent from the snapshot
// (models.py:21-31,43-44); structure follows canonical ENet as tabulated in SURVEY.md 8a,
// op semantics as in oracle/enet_oracle.py (torch fp32).
#include "simt_common.cuh"

namespace bc {

// ------------------------------------------------------------------------- head (K8)
// ConvTranspose2d(16, C, 3, stride 2, padding 1, output_padding 1, no bias) at 128x256 ->
// logits (C,256,512).  One thread per INPUT pixel (i,j) producing the 2x2 output quad:
//   out(2i  ,2j  ) = in(i,j) w11
//   out(2i  ,2j+1) = in(i,j+1) w10 + in(i,j) w12
//   out(2i+1,2j  ) = in(i+1,j) w01 + in(i,j) w21
//   out(2i+1,2j+1) = in(i+1,j+1) w00 + in(i+1,j) w02 + in(i,j+1) w20 + in(i,j) w22
// LABELS = false: fp32 NCHW logits (models.py:52).  LABELS = true: class argmax (first
// max wins, models.py:55) + LUT (models.py:56-58 / 79-80) fused; 1 B/px leaves the SM.
// w: [ky*3+kx][16][CP] fp32 with CP = 16 (C <= 16) or 32 (zero padded).
template <int CP>
__device__ __forceinline__ void fma_tap(float (&acc)[CP], const float (&v)[16], const float* __restrict__ w) {
#pragma unroll
  for (int k = 0; k < 16; ++k)
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = fmaf(v[k], w[k * CP + c], acc[c]);
}

template <typename T, int CP, bool LABELS>
__global__ void __launch_bounds__(128)
k_fullconv(const T* __restrict__ x, const float* __restrict__ w, float* __restrict__ logits,
           uint8_t* __restrict__ labels, Lut256 lut, int C, int total) {
  __shared__ __align__(16) float sw[9 * 16 * CP];
  for (int i = threadIdx.x; i < 9 * 16 * CP; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int H = 128, W = 256;
  int j = p % W, i = (p / W) % H, n = p / (W * H);
  float v00[16], v01[16], v10[16], v11[16];
  const bool right = j + 1 < W, below = i + 1 < H;
  ld_ch<16>(x + (size_t)p * 16, v00);
  if (right) ld_ch<16>(x + ((size_t)p + 1) * 16, v01);
  if (below) ld_ch<16>(x + ((size_t)p + W) * 16, v10);
  if (right && below) ld_ch<16>(x + ((size_t)p + W + 1) * 16, v11);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (!right) v01[k] = 0.f;
    if (!below) v10[k] = 0.f;
    if (!(right && below)) v11[k] = 0.f;
  }
  float acc[4][CP];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[q][c] = 0.f;
  constexpr int TS = 16 * CP;   // floats per tap
  fma_tap<CP>(acc[0], v00, sw + 4 * TS);   // w11
  fma_tap<CP>(acc[1], v01, sw + 3 * TS);   // w10
  fma_tap<CP>(acc[1], v00, sw + 5 * TS);   // w12
  fma_tap<CP>(acc[2], v10, sw + 1 * TS);   // w01
  fma_tap<CP>(acc[2], v00, sw + 7 * TS);   // w21
  fma_tap<CP>(acc[3], v11, sw + 0 * TS);   // w00
  fma_tap<CP>(acc[3], v10, sw + 2 * TS);   // w02
  fma_tap<CP>(acc[3], v01, sw + 6 * TS);   // w20
  fma_tap<CP>(acc[3], v00, sw + 8 * TS);   // w22
  const int OW = 512, OH = 256;
  if (LABELS) {
    uint8_t lab[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float best = acc[q][0];
      int bi = 0;
#pragma unroll
      for (int c = 1; c < CP; ++c)
        if (c < C && acc[q][c] > best) { best = acc[q][c]; bi = c; }
      lab[q] = lut.v[bi];
    }
    uint8_t* o = labels + ((size_t)(n * OH + 2 * i) * OW + 2 * j);
    *reinterpret_cast<uchar2*>(o) = make_uchar2(lab[0], lab[1]);
    *reinterpret_cast<uchar2*>(o + OW) = make_uchar2(lab[2], lab[3]);
  } else {
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      if (c < C) {
        float* o = logits + (((size_t)n * C + c) * OH + 2 * i) * OW + 2 * j;
        *reinterpret_cast<float2*>(o) = make_float2(acc[0][c], acc[1][c]);
        *reinterpret_cast<float2*>(o + OW) = make_float2(acc[2][c], acc[3][c]);
      }
    }
  }
}

template <typename T>
void launch_fullconv(const T* x, int B, int C, const float* w, float* logits, uint8_t* labels,
                     const Lut256* lut, cudaStream_t s) {
  int total = B * 128 * 256;
  int grid = (total + 127) / 128;
  Lut256 l{};
  if (lut) l = *lut;
  int cp = C <= 16 ? 16 : 32;   // weights arrive zero-padded to cp classes (api.cu)
#define BC_FC(CP_)                                                                           \
  if (cp == CP_) {                                                                            \
    if (labels) k_fullconv<T, CP_, true><<<grid, 128, 0, s>>>(x, w, logits, labels, l, C, total);  \
    else k_fullconv<T, CP_, false><<<grid, 128, 0, s>>>(x, w, logits, labels, l, C, total);        \
    return;                                                                                   \
  }
  BC_FC(16) BC_FC(32)
#undef BC_FC
}
template void launch_fullconv<float>(const float*, int, int, const float*, float*, uint8_t*, const Lut256*, cudaStream_t);
template void launch_fullconv<bf16>(const bf16*, int, int, const float*, float*, uint8_t*, const Lut256*, cudaStream_t);
template void launch_fullconv<f16>(const f16*, int, int, const float*, float*, uint8_t*, const Lut256*, cudaStream_t);

}  // namespace bc
