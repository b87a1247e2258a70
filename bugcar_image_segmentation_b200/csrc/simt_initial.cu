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

// ------------------------------------------------------------------- initial block
// conv3x3 s2 p1 (3->13) || maxpool 3x3 s2 p1 (or 2x2 s2: pool2) (3) -> cat -> BN -> PReLU.
// One thread per output pixel (128x256).  Source: uint8 BGR frame through the fp32
// normalisation LUT (models.py:89-91 fused), or the float/double NCHW tensor
// ENET.preprocess returns.  w: [27][13] ((c*3+ky)*3+kx major), g/b: BN scale/shift [16].
template <typename T, int KIND>
__global__ void __launch_bounds__(128)
k_initial(const void* __restrict__ xin, T* __restrict__ out, const float* __restrict__ w,
          const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ alpha,
          const float* __restrict__ lut, int total, int pool2) {
  __shared__ float sw[27 * 13];
  __shared__ float slut[768];
  __shared__ float sg[16], sb[16], sa[16];
  for (int i = threadIdx.x; i < 27 * 13; i += blockDim.x) sw[i] = w[i];
  if (KIND == 0) for (int i = threadIdx.x; i < 768; i += blockDim.x) slut[i] = lut[i];
  if (threadIdx.x < 16) { sg[threadIdx.x] = g[threadIdx.x]; sb[threadIdx.x] = b[threadIdx.x]; sa[threadIdx.x] = alpha[threadIdx.x]; }
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int OW = 256, OH = 128, IW = 512, IH = 256;
  int ox = p % OW, oy = (p / OW) % OH, n = p / (OW * OH);
  float acc[13];
#pragma unroll
  for (int o = 0; o < 13; ++o) acc[o] = 0.f;
  float mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    int iy = 2 * oy - 1 + ky;
    if (iy < 0 || iy >= IH) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      int ix = 2 * ox - 1 + kx;
      if (ix < 0 || ix >= IW) continue;
      float v[3];
      if (KIND == 0) {
        const uint8_t* s = (const uint8_t*)xin + ((size_t)(n * IH + iy) * IW + ix) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = slut[s[2 - c] * 3 + c];   // BGR -> RGB
      } else if (KIND == 1) {
        const float* s = (const float*)xin + (size_t)n * 3 * IH * IW + (size_t)iy * IW + ix;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = s[(size_t)c * IH * IW];
      } else {
        const double* s = (const double*)xin + (size_t)n * 3 * IH * IW + (size_t)iy * IW + ix;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (float)s[(size_t)c * IH * IW];   // TF feed cast
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (!pool2 || (ky > 0 && kx > 0)) mx[c] = fmaxf(mx[c], v[c]);   // 2x2 s2 pool = taps (1..2, 1..2) of the conv window
        const float* wr = sw + ((c * 3 + ky) * 3 + kx) * 13;
#pragma unroll
        for (int o = 0; o < 13; ++o) acc[o] = fmaf(v[c], wr[o], acc[o]);
      }
    }
  }
  float r[16];
#pragma unroll
  for (int o = 0; o < 13; ++o) r[o] = acc[o];
#pragma unroll
  for (int c = 0; c < 3; ++c) r[13 + c] = mx[c];
#pragma unroll
  for (int o = 0; o < 16; ++o) r[o] = prelu(fmaf(r[o], sg[o], sb[o]), sa[o]);
  st_ch<16>(out + (size_t)p * 16, r);
}

template <typename T>
void launch_initial(const void* x, int kind, int B, T* out, int pool_kernel, const float* w, const float* g,
                    const float* b, const float* alpha, const float* lut, cudaStream_t s) {
  const int pool2 = pool_kernel == 2;
  int total = B * 128 * 256;
  int grid = (total + 127) / 128;
  if (kind == 0) k_initial<T, 0><<<grid, 128, 0, s>>>(x, out, w, g, b, alpha, lut, total, pool2);
  else if (kind == 1) k_initial<T, 1><<<grid, 128, 0, s>>>(x, out, w, g, b, alpha, lut, total, pool2);
  else k_initial<T, 2><<<grid, 128, 0, s>>>(x, out, w, g, b, alpha, lut, total, pool2);
}
template void launch_initial<float>(const void*, int, int, float*, int, const float*, const float*, const float*, const float*, const float*, cudaStream_t);
template void launch_initial<bf16>(const void*, int, int, bf16*, int, const float*, const float*, const float*, const float*, const float*, cudaStream_t);
template void launch_initial<f16>(const void*, int, int, f16*, int, const float*, const float*, const float*, const float*, const float*, cudaStream_t);

// ------------------------------------------------------------- debug / parity export
// NHWC activation (storage type T) -> fp32 NCHW, for per-block parity tests.
template <typename T>
__global__ void __launch_bounds__(256)
k_export_nchw(const T* __restrict__ in, float* __restrict__ out, int C, int HW, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n*C*HW, NCHW order
  if (i >= total) return;
  size_t yx = i % HW, c = (i / HW) % C, n = i / ((size_t)HW * C);
  out[i] = (float)in[(n * HW + yx) * C + c];
}
template <typename T>
void launch_export_nchw(const T* in, float* out, int B, int C, int H, int W, cudaStream_t s) {
  size_t total = (size_t)B * C * H * W;
  k_export_nchw<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, out, C, H * W, total);
}
template void launch_export_nchw<float>(const float*, float*, int, int, int, int, cudaStream_t);
template void launch_export_nchw<bf16>(const bf16*, float*, int, int, int, int, cudaStream_t);
template void launch_export_nchw<f16>(const f16*, float*, int, int, int, int, cudaStream_t);

}  // namespace bc
