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

// ----------------------------------------------------- downsampling bottleneck, part a
// main: maxpool 2x2 s2 with argmax (2-bit window position, first max wins);
// ext:  conv 2x2 s2 (CIN->CI) + BN + PReLU.  One thread per half-resolution pixel.
// H, W are the OUTPUT (half) resolution.
template <typename T, int CIN, int CI>
__global__ void __launch_bounds__(128)
k_down_a(const T* __restrict__ x, T* __restrict__ pooled, uint8_t* __restrict__ idx,
         T* __restrict__ e1, const float* __restrict__ w, const float* __restrict__ bias,
         const float* __restrict__ alpha, int H, int W, int total) {
  extern __shared__ float sw[];   // [4][CIN][CI]
  for (int i = threadIdx.x; i < 4 * CIN * CI; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int ox = p % W, oy = (p / W) % H, n = p / (W * H);
  const int IW = 2 * W;
  const T* base = x + ((size_t)(n * 2 * H + 2 * oy) * IW + 2 * ox) * CIN;
  float acc[CI];
#pragma unroll
  for (int j = 0; j < CI; ++j) acc[j] = bias[j];
  constexpr int V = CIN >= 8 ? 8 : 4;
  for (int c0 = 0; c0 < CIN; c0 += V) {
    float best[V];
    uint8_t bi[V];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float v[V];
      ld_ch<V>(base + ((size_t)(t >> 1) * IW + (t & 1)) * CIN + c0, v);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        if (t == 0 || v[k] > best[k]) { best[k] = v[k]; bi[k] = (uint8_t)t; }
        const float* wr = sw + ((size_t)t * CIN + c0 + k) * CI;
#pragma unroll
        for (int j = 0; j < CI; ++j) acc[j] = fmaf(v[k], wr[j], acc[j]);
      }
    }
    st_ch<V>(pooled + (size_t)p * CIN + c0, best);
    uint8_t* ip = idx + (size_t)p * CIN + c0;
    if constexpr (V == 8) {
      uint2 t;
      t.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      t.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(ip) = t;
    } else {
      *reinterpret_cast<uint32_t*>(ip) = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    }
  }
#pragma unroll
  for (int j = 0; j < CI; ++j) acc[j] = prelu(acc[j], alpha[j]);
  st_ch<CI>(e1 + (size_t)p * CI, acc);
}

template <typename T>
void launch_down_a(const T* x, int B, int H, int W, int cin, int ci, T* pooled, uint8_t* idx,
                   T* e1, const ConvP& c1, cudaStream_t s) {
  int total = B * H * W;
  int grid = (total + 127) / 128;
  size_t smem = (size_t)4 * cin * ci * sizeof(float);
  if (cin == 16 && ci == 4)
    k_down_a<T, 16, 4><<<grid, 128, smem, s>>>(x, pooled, idx, e1, c1.w, c1.bias, c1.alpha, H, W, total);
  else if (cin == 64 && ci == 16)
    k_down_a<T, 64, 16><<<grid, 128, smem, s>>>(x, pooled, idx, e1, c1.w, c1.bias, c1.alpha, H, W, total);
}
template void launch_down_a<float>(const float*, int, int, int, int, int, float*, uint8_t*, float*, const ConvP&, cudaStream_t);
template void launch_down_a<bf16>(const bf16*, int, int, int, int, int, bf16*, uint8_t*, bf16*, const ConvP&, cudaStream_t);
template void launch_down_a<f16>(const f16*, int, int, int, int, int, f16*, uint8_t*, f16*, const ConvP&, cudaStream_t);

}  // namespace bc
