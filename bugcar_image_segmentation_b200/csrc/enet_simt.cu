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
// value as it will be re-read from storage (bf16 rounding point)
template <typename T> __device__ __forceinline__ float rnd(float v);
template <> __device__ __forceinline__ float rnd<float>(float v) { return v; }
template <> __device__ __forceinline__ float rnd<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : a * v; }

// ------------------------------------------------------------------- initial block
// conv3x3 s2 p1 (3->13, no bias) || maxpool3x3 s2 p1 (3) -> cat -> BN -> PReLU.
// One thread per output pixel (128x256).  Source: uint8 BGR frame through the fp32
// normalisation LUT (models.py:89-91 fused), or the float/double NCHW tensor
// ENET.preprocess returns.  w: [27][13] ((c*3+ky)*3+kx major), g/b: BN scale/shift [16].
template <typename T, int KIND>
__global__ void __launch_bounds__(128)
k_initial(const void* __restrict__ xin, T* __restrict__ out, const float* __restrict__ w,
          const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ alpha,
          const float* __restrict__ lut, int total) {
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
        mx[c] = fmaxf(mx[c], v[c]);
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
void launch_initial(const void* x, int kind, int B, T* out, const float* w, const float* g,
                    const float* b, const float* alpha, const float* lut, cudaStream_t s) {
  int total = B * 128 * 256;
  int grid = (total + 127) / 128;
  if (kind == 0) k_initial<T, 0><<<grid, 128, 0, s>>>(x, out, w, g, b, alpha, lut, total);
  else if (kind == 1) k_initial<T, 1><<<grid, 128, 0, s>>>(x, out, w, g, b, alpha, lut, total);
  else k_initial<T, 2><<<grid, 128, 0, s>>>(x, out, w, g, b, alpha, lut, total);
}
template void launch_initial<float>(const void*, int, int, float*, const float*, const float*, const float*, const float*, const float*, cudaStream_t);
template void launch_initial<bf16>(const void*, int, int, bf16*, const float*, const float*, const float*, const float*, const float*, cudaStream_t);

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

// ------------------------------------------------------------------------- head (K8)
// ConvTranspose2d(16, C, 3, stride 2, padding 1, output_padding 1, no bias) at 128x256 ->
// logits (C,256,512).  One thread per INPUT pixel (i,j) producing the 2x2 output quad:
//   out(2i  ,2j  ) = in(i,j) w11
//   out(2i  ,2j+1) = in(i,j+1) w10 + in(i,j) w12
//   out(2i+1,2j  ) = in(i+1,j) w01 + in(i,j) w21
//   out(2i+1,2j+1) = in(i+1,j+1) w00 + in(i+1,j) w02 + in(i,j+1) w20 + in(i,j) w22
// LABELS = false: fp32 NCHW logits (models.py:52).  LABELS = true: class argmax (first
// max wins, models.py:55) + LUT (models.py:56-58 / 79-80) fused; 1 B/px leaves the SM.
// w: [ky*3+kx][16][CP] fp32 with CP = C rounded up to 4 (zero padded).
template <typename T, int CP, bool LABELS>
__global__ void __launch_bounds__(128)
k_fullconv(const T* __restrict__ x, const float* __restrict__ w, float* __restrict__ logits,
           uint8_t* __restrict__ labels, Lut256 lut, int C, int total) {
  __shared__ float sw[9 * 16 * CP];
  for (int i = threadIdx.x; i < 9 * 16 * CP; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int H = 128, W = 256;
  int j = p % W, i = (p / W) % H, n = p / (W * H);
  float acc[4][CP];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[q][c] = 0.f;
  // neighbour (di,dj) contributes to output q=(qy*2+qx) through tap (ky,kx)
#pragma unroll
  for (int di = 0; di < 2; ++di)
#pragma unroll
    for (int dj = 0; dj < 2; ++dj) {
      if (i + di >= H || j + dj >= W) continue;
      float v[16];
      ld_ch<16>(x + ((size_t)p + (size_t)di * W + dj) * 16, v);
#pragma unroll
      for (int qy = di; qy < 2; ++qy)
#pragma unroll
        for (int qx = dj; qx < 2; ++qx) {
          // rows: qy=0 uses (di=0,ky=1); qy=1 uses (di=1,ky=0) and (di=0,ky=2)
          int ky = (qy == 0) ? 1 : (di ? 0 : 2);
          int kx = (qx == 0) ? 1 : (dj ? 0 : 2);
          const float* wr = sw + (size_t)(ky * 3 + kx) * 16 * CP;
#pragma unroll
          for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[qy * 2 + qx][c] = fmaf(v[k], wr[k * CP + c], acc[qy * 2 + qx][c]);
        }
    }
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
      if (c >= C) break;
      float* o = logits + (((size_t)n * C + c) * OH + 2 * i) * OW + 2 * j;
      *reinterpret_cast<float2*>(o) = make_float2(acc[0][c], acc[1][c]);
      *reinterpret_cast<float2*>(o + OW) = make_float2(acc[2][c], acc[3][c]);
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
  int cp = (C + 3) / 4 * 4;
#define BC_FC(CP_)                                                                           \
  if (cp == CP_) {                                                                            \
    if (labels) k_fullconv<T, CP_, true><<<grid, 128, 0, s>>>(x, w, logits, labels, l, C, total);  \
    else k_fullconv<T, CP_, false><<<grid, 128, 0, s>>>(x, w, logits, labels, l, C, total);        \
    return;                                                                                   \
  }
  BC_FC(4) BC_FC(8) BC_FC(12) BC_FC(16) BC_FC(20) BC_FC(24) BC_FC(32)
#undef BC_FC
}
template void launch_fullconv<float>(const float*, int, int, const float*, float*, uint8_t*, const Lut256*, cudaStream_t);
template void launch_fullconv<bf16>(const bf16*, int, int, const float*, float*, uint8_t*, const Lut256*, cudaStream_t);

}  // namespace bc
