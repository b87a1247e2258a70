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

// ------------------------------------------------------- upsampling bottleneck, part b
// Per low-resolution pixel p (one thread per pixel and group of CPT output channels):
//   main[c]  = (1x1 CIN->COUT + BN)(x[p])                       (no activation)
//   for the 4 output pixels q = (2y+dy, 2x+dx), tap = dy*2+dx:
//     e2 = act( tconv2x2 tap (CI->CI) + BN )(e1[p])             (rounded to storage)
//     e3 = (1x1 CI->COUT + BN)(e2)                              (no activation)
//     out[q][c] = act_out( e3[c] + (idx[p][c] == tap ? main[c] : 0) )   (max-unpool)
template <typename T, int CIN, int CI, int COUT, int CPT>
__global__ void __launch_bounds__(128)
k_up_b(const T* __restrict__ x, const T* __restrict__ e1, const uint8_t* __restrict__ idx,
       T* __restrict__ out, const float* __restrict__ wm, const float* __restrict__ bm,
       const float* __restrict__ wt, const float* __restrict__ bt, const float* __restrict__ at,
       const float* __restrict__ w3, const float* __restrict__ b3,
       const float* __restrict__ alpha_out, int H, int W, int total) {
  extern __shared__ float sm[];
  float* swm = sm;                       // [CIN][CPT]
  float* swt = swm + CIN * CPT;          // [4][CI][CI]
  float* sw3 = swt + 4 * CI * CI;        // [CI][CPT]
  const int g0 = blockIdx.y * CPT;
  for (int i = threadIdx.x; i < CIN * CPT; i += blockDim.x) swm[i] = wm[(size_t)(i / CPT) * COUT + g0 + i % CPT];
  for (int i = threadIdx.x; i < 4 * CI * CI; i += blockDim.x) swt[i] = wt[i];
  for (int i = threadIdx.x; i < CI * CPT; i += blockDim.x) sw3[i] = w3[(size_t)(i / CPT) * COUT + g0 + i % CPT];
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int px = p % W, py = (p / W) % H, n = p / (W * H);
  float mainv[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) mainv[j] = bm[g0 + j];
  for (int c0 = 0; c0 < CIN; c0 += 8) {
    float v[8];
    ld_ch<8>(x + (size_t)p * CIN + c0, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float* wr = swm + (size_t)(c0 + k) * CPT;
#pragma unroll
      for (int j = 0; j < CPT; ++j) mainv[j] = fmaf(v[k], wr[j], mainv[j]);
    }
  }
  float ev[CI];
  ld_ch<CI>(e1 + (size_t)p * CI, ev);
  // pool indices of this pixel's CPT channels (COUT channels per low-res pixel)
  uint8_t pi[CPT];
  {
    const uint4* ip = reinterpret_cast<const uint4*>(idx + (size_t)p * COUT + g0);
#pragma unroll
    for (int q = 0; q < CPT / 16; ++q) {
      uint4 t = ip[q];
      const uint8_t* b = reinterpret_cast<const uint8_t*>(&t);
#pragma unroll
      for (int k = 0; k < 16; ++k) pi[q * 16 + k] = b[k];
    }
  }
#pragma unroll 1
  for (int tap = 0; tap < 4; ++tap) {
    float e2[CI];
#pragma unroll
    for (int j = 0; j < CI; ++j) e2[j] = bt[j];
#pragma unroll
    for (int k = 0; k < CI; ++k) {
      const float* wr = swt + ((size_t)tap * CI + k) * CI;
#pragma unroll
      for (int j = 0; j < CI; ++j) e2[j] = fmaf(ev[k], wr[j], e2[j]);
    }
#pragma unroll
    for (int j = 0; j < CI; ++j) e2[j] = rnd<T>(prelu(e2[j], at[j]));
    float o[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) o[j] = b3[g0 + j];
#pragma unroll
    for (int k = 0; k < CI; ++k) {
      const float* wr = sw3 + (size_t)k * CPT;
#pragma unroll
      for (int j = 0; j < CPT; ++j) o[j] = fmaf(e2[k], wr[j], o[j]);
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      float m = (pi[j] == tap) ? mainv[j] : 0.f;
      o[j] = prelu(o[j] + m, alpha_out[g0 + j]);
    }
    int oy = 2 * py + (tap >> 1), ox = 2 * px + (tap & 1);
    st_ch<CPT>(out + ((size_t)(n * 2 * H + oy) * (2 * W) + ox) * COUT + g0, o);
  }
}

template <typename T>
void launch_up_b(const T* x, const T* e1, const uint8_t* idx, T* out, const Bottleneck& bn,
                 int B, int H, int W, cudaStream_t s) {
  int total = B * H * W;
  if (bn.cin == 128 && bn.cout == 64) {
    constexpr int CIN = 128, CI = 32, COUT = 64, CPT = 32;
    size_t smem = (size_t)(CIN * CPT + 4 * CI * CI + CI * CPT) * sizeof(float);
    dim3 grid((total + 127) / 128, COUT / CPT);
    k_up_b<T, CIN, CI, COUT, CPT><<<grid, 128, smem, s>>>(x, e1, idx, out, bn.cm.w, bn.cm.bias,
        bn.c2.w, bn.c2.bias, bn.c2.alpha, bn.c3.w, bn.c3.bias, bn.alpha_out, H, W, total);
  } else if (bn.cin == 64 && bn.cout == 16) {
    constexpr int CIN = 64, CI = 16, COUT = 16, CPT = 16;
    size_t smem = (size_t)(CIN * CPT + 4 * CI * CI + CI * CPT) * sizeof(float);
    dim3 grid((total + 127) / 128, COUT / CPT);
    k_up_b<T, CIN, CI, COUT, CPT><<<grid, 128, smem, s>>>(x, e1, idx, out, bn.cm.w, bn.cm.bias,
        bn.c2.w, bn.c2.bias, bn.c2.alpha, bn.c3.w, bn.c3.bias, bn.alpha_out, H, W, total);
  }
}
template void launch_up_b<float>(const float*, const float*, const uint8_t*, float*, const Bottleneck&, int, int, int, cudaStream_t);
template void launch_up_b<bf16>(const bf16*, const bf16*, const uint8_t*, bf16*, const Bottleneck&, int, int, int, cudaStream_t);
template void launch_up_b<f16>(const f16*, const f16*, const uint8_t*, f16*, const Bottleneck&, int, int, int, cudaStream_t);

}  // namespace bc
