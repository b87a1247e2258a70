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

// -------------------------------------------------------------------- generic conv
// out[p][g*CPT + j] = act( bias + sum_{tap, c} in[p + tap][c] * w[tap][c][g*CPT+j] )
// RES: out = act_out( act(...) + res[p][c] (c < res_ch) )   -- bottleneck tail
// One thread per pixel and per group of CPT output channels (blockIdx.y = group).
// Zero padding: taps that leave the image are skipped.
template <typename T, int CIN, int COUT, int CPT, int NT, bool RES>
__global__ void __launch_bounds__(128)
k_conv(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ res, int res_ch,
       const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ alpha,
       const float* __restrict__ alpha_out, int H, int W, int total, Taps taps) {
  extern __shared__ float sw[];   // [NT][CIN][CPT] slice of this group
  const int g0 = blockIdx.y * CPT;
  for (int i = threadIdx.x; i < NT * CIN * CPT; i += blockDim.x) {
    int j = i % CPT, tc = i / CPT;
    sw[i] = w[(size_t)tc * COUT + g0 + j];
  }
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int x = p % W, y = (p / W) % H;
  float acc[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = bias[g0 + j];
  constexpr int V = CIN >= 8 ? 8 : 4;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    int yy = y + taps.dy[t], xx = x + taps.dx[t];
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const T* ip = in + ((size_t)p + (size_t)taps.dy[t] * W + taps.dx[t]) * CIN;
#pragma unroll 2
    for (int c0 = 0; c0 < CIN; c0 += V) {
      float v[V];
      ld_ch<V>(ip + c0, v);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float* wr = sw + ((size_t)t * CIN + c0 + k) * CPT;
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[j] = fmaf(v[k], wr[j], acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = prelu(acc[j], alpha[g0 + j]);
  if (RES) {
    constexpr int RV = CPT >= 8 ? 8 : 4;
#pragma unroll
    for (int j0 = 0; j0 < CPT; j0 += RV) {
      if (g0 + j0 < res_ch) {
        float r[RV];
        ld_ch<RV>(res + (size_t)p * res_ch + g0 + j0, r);
#pragma unroll
        for (int k = 0; k < RV; ++k) acc[j0 + k] += r[k];
      }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = prelu(acc[j], alpha_out[g0 + j]);
  }
  st_ch<CPT>(out + (size_t)p * COUT + g0, acc);
}

template <typename T, int CIN, int COUT, int NT>
static void conv_dispatch(const T* in, T* out, const T* res, int res_ch, const ConvP& c,
                          const float* alpha_out, int B, int H, int W, const Taps& taps,
                          cudaStream_t s) {
  constexpr int CPT = COUT > 32 ? 32 : COUT;
  int total = B * H * W;
  dim3 grid((total + 127) / 128, COUT / CPT);
  size_t smem = (size_t)NT * CIN * CPT * sizeof(float);
  if (res) {
    auto k = k_conv<T, CIN, COUT, CPT, NT, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, 128, smem, s>>>(in, out, res, res_ch, c.w, c.bias, c.alpha, alpha_out, H, W, total, taps);
  } else {
    auto k = k_conv<T, CIN, COUT, CPT, NT, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, 128, smem, s>>>(in, out, res, res_ch, c.w, c.bias, c.alpha, alpha_out, H, W, total, taps);
  }
}

template <typename T>
void launch_conv(const T* in, T* out, const T* res, int res_ch, const ConvP& c,
                 const float* alpha_out, int B, int H, int W, const Taps& taps, cudaStream_t s) {
#define BC_CASE(CIN_, COUT_, NT_)                                                          \
  if (c.cin == CIN_ && c.cout == COUT_ && c.ntaps == NT_) {                                 \
    conv_dispatch<T, CIN_, COUT_, NT_>(in, out, res, res_ch, c, alpha_out, B, H, W, taps, s); \
    return;                                                                                 \
  }
  // 1x1 projections / expansions
  BC_CASE(16, 4, 1) BC_CASE(4, 16, 1) BC_CASE(64, 16, 1) BC_CASE(16, 64, 1)
  BC_CASE(128, 32, 1) BC_CASE(32, 128, 1) BC_CASE(4, 64, 1) BC_CASE(16, 128, 1)
  // 3x3 (regular / dilated)
  BC_CASE(4, 4, 9) BC_CASE(16, 16, 9) BC_CASE(32, 32, 9)
  // asymmetric 5x1 / 1x5
  BC_CASE(32, 32, 5)
#undef BC_CASE
}
template void launch_conv<float>(const float*, float*, const float*, int, const ConvP&, const float*, int, int, int, const Taps&, cudaStream_t);
template void launch_conv<bf16>(const bf16*, bf16*, const bf16*, int, const ConvP&, const float*, int, int, int, const Taps&, cudaStream_t);
template void launch_conv<f16>(const f16*, f16*, const f16*, int, const ConvP&, const float*, int, int, int, const Taps&, cudaStream_t);

}  // namespace bc
