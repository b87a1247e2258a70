// ENet initial block on tcgen05: conv3x3 s2 p1 (3 -> 13, no bias) || maxpool3x3 s2 p1 (3) -> cat
// -> BN -> PReLU -> bf16 NHWC, with the frame normalisation of ENET.preprocess (models.py:89-91)
// fused in for uint8 BGR frames.
//
// The conv is a GEMM with M = 128 output pixels, K = 27 (padded to 32), N = 13 (padded to 16).
// The im2col rows are built by the CTA's threads (one output pixel = one A row = one TMEM lane per
// thread) because the operand is not a plain view of the input: it is LUT-normalised, zero padded
// AFTER normalisation, and stride 2.  To keep fp32 accuracy on a tensor pipe that reads tf32 the
// product is split 3 ways: x = xh + xl, w = wh + wl (h = top 19 bits), acc += xh*wh + xl*wh + xh*wl
// (the dropped xl*wl term is < 2^-20 relative), so results agree with the CUDA-core fp32 kernel to
// fp32 round-off and all three input kinds (uint8 BGR / fp32 / fp64 NCHW) give identical outputs.
//
// A CTA handles one tile at a time; overlap comes from several co-resident CTAs per SM (36 KB of
// shared memory, 32 TMEM columns each).  Warp 0 issues the MMAs, warps 1-4 build A and run the
// epilogue.  Semantics: oracle/enet_oracle.py `initial` (the frozen graph the reference runs,
// models.py:43-44).
#include "umma_common.cuh"

#include <cstring>

namespace bc {
namespace BC_NS {

struct InitParams {
  int num_tiles;             // B * 128 * 256 / 128
  const void* x;             // uint8 (B,256,512,3) BGR | float / double (B,3,256,512)
  act_t* out;                // (B,128,256,16)
  const uint8_t* wblob;      // [hi | lo][16 rows][32 k] tf32 (fp32 bit patterns), 128-byte swizzled rows
  const float* lut;          // [256][3] fp32 normalisation table (RGB order)
  float f[48];               // BN scale g[16], shift b[16], PReLU slope a[16]
  float fu8[16];             // uint8 kernel: g[o] * 2^-e(o), the un-scale of its fp16 weight limbs folded in
  int pool2;                 // 1: the max-pool is 2x2 stride 2 (taps ky, kx in 1..2 of the conv window) instead of 3x3 s2 p1
  float fpool[6];            // uint8 kernel, pooled channels: byte -> BN(normalise(byte)) = byte * fpool[c] + fpool[3 + c]
};

static constexpr int INIT_A = 128 * 128;          // one A tile: 128 rows x 32 tf32
static constexpr int INIT_W = 16 * 128;           // one B tile: 16 rows x 32 tf32
static constexpr int INIT_OFF_A = 0;              // A hi, A lo
static constexpr int INIT_OFF_W = 2 * INIT_A;     // B hi, B lo
static constexpr int INIT_OFF_LUT = INIT_OFF_W + 2 * INIT_W;
static constexpr int INIT_OFF_BAR = INIT_OFF_LUT + 768 * 4;
static constexpr int INIT_SMEM = INIT_OFF_BAR + 64;
static constexpr int INIT_MINB = 5;

// tf32 x tf32 -> fp32, both operands K-major: c_format F32 (bit 4), a/b_format TF32 = 2 (bits 7, 10)
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// executed by a whole (convergent) warp: one elected lane issues
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred pe, p;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(160, INIT_MINB)
k_umma_initial(const __grid_constant__ InitParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + INIT_OFF_BAR);
  enum { A_FULL = 0, D_FULL, W_FULL, NBARS };
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  uint32_t* tmem_slot = (uint32_t*)&bars[NBARS];
  float* slut = (float*)(smem + INIT_OFF_LUT);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index made provably warp-uniform: role branches become uniform branches

  if (tid == 0) {
    mbar_init(bar(A_FULL), 128);
    mbar_init(bar(D_FULL), 1);
    mbar_init(bar(W_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar(W_FULL), 2 * INIT_W);
    bulk_load(sbase + INIT_OFF_W, p.wblob, 2 * INIT_W, bar(W_FULL));
  }
  if (KIND == 0)
    for (int i = tid; i < 768; i += 160) slut[i] = p.lut[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  constexpr uint32_t IDESC = instr_desc_tf32(128, 16);
  const int OW = 256, OH = 128, IW = 512, IH = 256;

  if (warp == 0) {
    mbar_wait(bar(W_FULL), 0);
    for (int k = 0; k < T; ++k) {
      mbar_wait(bar(A_FULL), k & 1);
      tc_fence_after();
      // acc = Ah*Wh + Al*Wh + Ah*Wl, K = 32 as four K = 8 steps each
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const uint32_t a = sbase + INIT_OFF_A + (part == 1 ? INIT_A : 0);
        const uint32_t b = sbase + INIT_OFF_W + (part == 2 ? INIT_W : 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_tf32(tmem, smem_desc<128>(a + kk * 32), smem_desc<128>(b + kk * 32), IDESC, (part | kk) != 0);
      }
      umma_commit_e(bar(D_FULL));
    }
  } else {
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const uint32_t tm_lane = tmem + ((uint32_t)(q4 * 32) << 16);
    for (int k = 0; k < T; ++k) {
      const int tile = blockIdx.x + k * gridDim.x;
      const int pix = tile * 128 + m;
      const int ox = pix % OW, oy = (pix / OW) % OH, n = pix / (OW * OH);
      // ---- im2col row (k = (c*3 + ky)*3 + kx) + the 3-channel max-pool
      float a[32];
#pragma unroll
      for (int i = 27; i < 32; ++i) a[i] = 0.f;
      float mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = 2 * ox - 1 + kx;
          float v[3] = {0.f, 0.f, 0.f};                            // zero padding (of the normalised image)
          if (iy >= 0 && iy < IH && ix >= 0 && ix < IW) {
            if (KIND == 0) {
              const uint8_t* s = (const uint8_t*)p.x + ((size_t)(n * IH + iy) * IW + ix) * 3;
#pragma unroll
              for (int c = 0; c < 3; ++c) v[c] = slut[s[2 - c] * 3 + c];   // BGR -> RGB, (u/256 - mean)/std
            } else if (KIND == 1) {
              const float* s = (const float*)p.x + (size_t)n * 3 * IH * IW + (size_t)iy * IW + ix;
#pragma unroll
              for (int c = 0; c < 3; ++c) v[c] = s[(size_t)c * IH * IW];
            } else {
              const double* s = (const double*)p.x + (size_t)n * 3 * IH * IW + (size_t)iy * IW + ix;
#pragma unroll
              for (int c = 0; c < 3; ++c) v[c] = (float)s[(size_t)c * IH * IW];   // TF feed cast
            }
            if (!p.pool2 || (ky > 0 && kx > 0)) {
#pragma unroll
              for (int c = 0; c < 3; ++c) mx[c] = fmaxf(mx[c], v[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) a[(c * 3 + ky) * 3 + kx] = v[c];
        }
      }
      // the previous tile's MMAs have finished reading A: this thread passed its D_FULL wait below
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(a[4 * j]) & 0xffffe000u);
        hi.y = __uint_as_float(__float_as_uint(a[4 * j + 1]) & 0xffffe000u);
        hi.z = __uint_as_float(__float_as_uint(a[4 * j + 2]) & 0xffffe000u);
        hi.w = __uint_as_float(__float_as_uint(a[4 * j + 3]) & 0xffffe000u);
        lo = make_float4(a[4 * j] - hi.x, a[4 * j + 1] - hi.y, a[4 * j + 2] - hi.z, a[4 * j + 3] - hi.w);
        const uint32_t off = swz<128>((uint32_t)(m * 128 + j * 16));
        *reinterpret_cast<float4*>(smem + INIT_OFF_A + off) = hi;
        *reinterpret_cast<float4*>(smem + INIT_OFF_A + INIT_A + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(bar(A_FULL));
      // ---- epilogue: BN + PReLU on the 13 conv channels and the 3 pooled ones
      mbar_wait(bar(D_FULL), k & 1);
      tc_fence_after();
      float r[16];
      tmem_ld16(tm_lane, r);
      tc_fence_before();
#pragma unroll
      for (int c = 0; c < 3; ++c) r[13 + c] = mx[c];
#pragma unroll
      for (int o = 0; o < 16; ++o) r[o] = prelu_f(fmaf(r[o], p.f[o], p.f[16 + o]), p.f[32 + o]);
      uint4* o = reinterpret_cast<uint4*>(p.out + (size_t)pix * 16);
      o[0] = pack8(r);
      o[1] = pack8(r + 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
  }
}

// ---------------------------------------------------------------------------------------------
// uint8 frames: the normalisation is affine, v = s_c * u + t_c with s_c = 1 / (256 std_c) and
// t_c = -mean_c / std_c (models.py:17-18,91), so
//     sum_k w_k v_k = sum_k (w_k s_c) u_k + sum_taps valid(tap) * (sum_c w_(c,tap) t_c)
// where valid(tap) = 0 for taps in the zero padding (the conv pads the NORMALISED image).  The A
// operand is then the RAW bytes (exact in fp16) plus nine 0/1 validity columns: no lookup table, no
// fp32 operand split on the activation side.  The 27 window bytes of an output pixel arrive as nine
// aligned 32-bit words; a byte pair becomes two fp16 values with ONE byte permute (0x64 above each
// byte: 0x64bb is the fp16 number 1024 + bb) and ONE packed subtraction of 1024 -- no integer-to-float
// conversions, no per-byte extraction; the 3-channel max-pool is a packed byte maximum on the same
// words.  K order (any order is legal as long as B uses the same): per window row ky the bytes
// 0..7 (8 halves), then the three bytes 8, then the validity flags.  Only the first image row / column has taps
// in the padding, so the nine tap flags collapse into four columns -- (oy > 0 and ox > 0), (oy > 0), (ox > 0), 1 --
// whose weights are the sums of the offset terms of the taps they stand for: K = 27 + 4 = 31 -> 32, two K steps.
// The folded weights are U8_LIMBS fp16 limbs of  w_k s_c 2^e(o)  (22 significant bits with two limbs, far below
// the 16-bit rounding of the block's output) with a per-output-channel power of two e(o) that lifts the row into
// fp16's normal range; 2^-e(o) is folded into the batch-norm scale of the epilogue.  The kernel's time is set by
// the number of MMAs per tile (N = 16 MMAs cost their fixed issue interval, not their flops; more resident CTAs
// change nothing: 7, 8 and 9 per SM all measured 148 us with 9 MMAs per tile), hence K = 32 and two limbs: 4 MMAs.
static constexpr int U8_A = 128 * 128;            // A tile: 128 rows x 64 fp16 (48 used)
static constexpr int U8_W = 16 * 128;             // one limb of B: 16 rows x 64 fp16
static constexpr int U8_K = 32;
#ifndef BC_U8_LIMBS
#define BC_U8_LIMBS 2
#endif
static constexpr int U8_LIMBS = BC_U8_LIMBS;
static constexpr int U8_OFF_A = 0;
static constexpr int U8_OFF_W = U8_A;
static constexpr int U8_OFF_BAR = U8_OFF_W + U8_LIMBS * U8_W;
static constexpr int U8_SMEM = U8_OFF_BAR + 64;
#ifndef BC_U8_MINB
#define BC_U8_MINB 7
#endif
static constexpr int U8_MINB = BC_U8_MINB;    // (a second A tile so that tile k+1 is built during tile k's MMAs, 5 CTAs/SM, measured 160 vs 150 us)

// K index of window byte b (0..8) of row ky, and of the validity column of tap (ky, kx):
// 27: first row and first column (needs oy > 0 and ox > 0), 28: first row, 29: first column, 30: always valid
__host__ __device__ constexpr int u8_k_byte(int ky, int b) { return b < 8 ? ky * 8 + b : 24 + ky; }
__host__ __device__ constexpr int u8_k_valid(int ky, int kx) { return ky == 0 ? (kx == 0 ? 27 : 28) : (kx == 0 ? 29 : 30); }

// bytes i and j of `w` as the fp16 pair (i in the low half)
template <int I, int J>
__device__ __forceinline__ uint32_t bytes_to_h2(uint32_t w) {
  const uint32_t biased = __byte_perm(w, 0x64646464u, (4u << 12) | ((uint32_t)J << 8) | (4u << 4) | (uint32_t)I);   // 1024 + byte
  const __half2 r = __hsub2(*reinterpret_cast<const __half2*>(&biased), __half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400)));
  return *reinterpret_cast<const uint32_t*>(&r);
}

__global__ void __launch_bounds__(160, U8_MINB)
k_umma_initial_u8(const __grid_constant__ InitParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + U8_OFF_BAR);
  enum { A_FULL0 = 0, D_FULL0, W_FULL, NBARS };
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  uint32_t* tmem_slot = (uint32_t*)&bars[NBARS];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index made provably warp-uniform: role branches become uniform branches

  if (tid == 0) {
    mbar_init(bar(A_FULL0), 128);
    mbar_init(bar(D_FULL0), 1);
    mbar_init(bar(W_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar(W_FULL), U8_LIMBS * U8_W);
    bulk_load(sbase + U8_OFF_W, p.wblob, U8_LIMBS * U8_W, bar(W_FULL));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // (the MMAs read K columns 0..31 of an A row only, and every tile writes all of them)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int OW = 256, OH = 128, IW = 512, IH = 256;

  if (warp == 0) {
    const uint64_t dA0 = smem_desc<128>(sbase + U8_OFF_A), dB0 = smem_desc<128>(sbase + U8_OFF_W);
    constexpr uint32_t IDESC = instr_desc_fmt(128, 16, 0u);       // fp16 bytes x fp16 weight limbs, whatever act_t is
    mbar_wait(bar(W_FULL), 0);
    for (int k = 0; k < T; ++k) {
      // A_FULL(k) also says that the accumulator is free: every thread read tile k-1's result before it built tile k
      mbar_wait(bar(A_FULL0), k & 1);
      tc_fence_after();
#pragma unroll
      for (int limb = 0; limb < U8_LIMBS; ++limb)
#pragma unroll
        for (int kk = 0; kk < U8_K / 16; ++kk)
          umma_mma_e(tmem, dA0 + (uint64_t)(kk * 2), dB0 + (uint64_t)(limb * (U8_W >> 4) + kk * 2), IDESC, (limb | kk) != 0);
      umma_commit_e(bar(D_FULL0));
    }
  } else {
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const uint32_t tm_lane = tmem + ((uint32_t)(q4 * 32) << 16);
    // The 3 x 3 pixel window of an output pixel is 9 contiguous bytes per input row, starting at byte
    // 6*ox - 3 of the row: fetched as the three aligned 32-bit words that cover them, realigned with
    // funnel shifts.  ox = 0: the first word would lie before the row; its bytes are padding anyway.
    // Words of rows outside the image stay zero, so every padding tap is a zero byte.
    // The words of tile k+1 are requested before this thread waits for tile k's accumulator, so the global-load
    // latency hides behind the MMA round trip (the other resident CTAs hide the round trip itself).
    uint32_t w[9];
    int pix = 0, ox = 0, oy = 0;
    auto fetch = [&](int k) {
      pix = ((int)blockIdx.x + k * (int)gridDim.x) * 128 + m;
      ox = pix % OW;
      oy = (pix / OW) % OH;
      const int n = pix / (OW * OH);
      const int wbase = (6 * ox - 3) & ~3;              // -4 for ox = 0
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy - 1 + ky;
        w[3 * ky] = w[3 * ky + 1] = w[3 * ky + 2] = 0u;
        if (iy >= 0 && iy < IH) {
          const uint8_t* row = (const uint8_t*)p.x + (size_t)(n * IH + iy) * IW * 3;
          if (wbase >= 0) w[3 * ky] = __ldg(reinterpret_cast<const uint32_t*>(row + wbase));
          w[3 * ky + 1] = __ldg(reinterpret_cast<const uint32_t*>(row + wbase + 4));
          w[3 * ky + 2] = __ldg(reinterpret_cast<const uint32_t*>(row + wbase + 8));
        }
      }
    };
    // builds the A row of the tile whose words are in w[]; returns the max-pool of the window as fp16: (B, G) and R.
    // The pool runs on the fp16 pairs the A row is made of (one HMNMX2 per pair; a packed BYTE maximum is seven
    // instructions on this architecture, and the nine of them were a third of this function): first across the
    // window rows (the three rows hold the same pixel / channel at the same position), then across the pixels.
    auto hmax2u = [](uint32_t a, uint32_t b) -> uint32_t {
      const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
      return *reinterpret_cast<const uint32_t*>(&r);
    };
    auto build = [&]() -> uint2 {
      uint8_t* A = smem + U8_OFF_A;
      const int sh = ((6 * ox - 3) & 3) * 8;            // 8 or 24
      uint32_t r2[3], mh[4] = {0u, 0u, 0u, 0u};         // mh: window bytes (0,1) (2,3) (4,5) (6,7), maximum over the rows
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        // window bytes 0-3, 4-7, 8 (B G R of the three pixels, in memory order)
        const uint32_t r0 = __funnelshift_r(w[3 * ky], w[3 * ky + 1], sh), r1 = __funnelshift_r(w[3 * ky + 1], w[3 * ky + 2], sh);
        r2[ky] = (w[3 * ky + 2] >> sh) & 0xffu;
        const uint4 h = make_uint4(bytes_to_h2<0, 1>(r0), bytes_to_h2<2, 3>(r0), bytes_to_h2<0, 1>(r1), bytes_to_h2<2, 3>(r1));
        *reinterpret_cast<uint4*>(A + swz<128>((uint32_t)(m * 128 + ky * 16))) = h;
        if (ky == 0) {
          if (!p.pool2) { mh[0] = h.x; mh[1] = h.y; mh[2] = h.z; mh[3] = h.w; }     // 2x2 s2 pool: window rows 1..2 only
        } else {
          mh[0] = hmax2u(mh[0], h.x); mh[1] = hmax2u(mh[1], h.y); mh[2] = hmax2u(mh[2], h.z); mh[3] = hmax2u(mh[3], h.w);
        }
      }
      // bytes 8 of the three rows, then the validity columns (1.0 = 0x3c00): only the first row / column of the
      // image has taps in the padding (2*oy + 1 <= 255 and 2*ox + 1 <= 511 always are inside)
      const uint32_t vt = oy > 0 ? 0x3c00u : 0u, vl = ox > 0 ? 0x3c00u : 0u, vtl = (oy > 0 && ox > 0) ? 0x3c00u : 0u;
      const uint32_t b8a = bytes_to_h2<0, 2>(r2[0] | (r2[1] << 16)), b8b = bytes_to_h2<0, 1>(r2[2]) & 0xffffu;
      *reinterpret_cast<uint4*>(A + swz<128>((uint32_t)(m * 128 + 3 * 16))) = make_uint4(b8a, b8b | (vtl << 16), vt | (vl << 16), 0x3c00u);
      fence_proxy_async();
      mbar_arrive(bar(A_FULL0));
      // across the pixels: B = max(b0, b3, b6), G = max(b1, b4, b7), R = max(b2, b5, b8); the 2x2 pool leaves out
      // pixel 0 (b0..b2) and row 0 (zero is the identity: the values are bytes)
      const uint32_t px0 = p.pool2 ? 0u : 0xffffffffu;
      const uint32_t bg = hmax2u(hmax2u(mh[0] & px0, __byte_perm(mh[1], mh[2], 0x5432)), mh[3]);      // (b3, b4)
      const uint32_t r8 = hmax2u(p.pool2 ? (b8a & 0xffff0000u) : b8a, __byte_perm(b8b, 0u, 0x1010));  // (b8 rows 0 1) vs (row 2, row 2)
      const uint32_t rr = hmax2u(__byte_perm(mh[1] & px0, mh[2], 0x7610), r8);                        // (b2, b5)
      return make_uint2(bg, hmax2u(rr, rr >> 16));
    };
    if (T > 0) fetch(0);
    for (int k = 0; k < T; ++k) {
      // (the MMAs of tile k-1 have finished reading A: this thread passed its D_FULL wait below)
      const int pix_cur = pix;
      const uint2 mx_cur = build();
      if (k + 1 < T) fetch(k + 1);
      // ---- epilogue of tile k: BN (with the weights' power-of-two un-scale folded in) + PReLU on the 13 conv
      // channels; the 3 pooled channels: max-pool of the normalised image = normalisation of the max byte (the
      // map is increasing), and normalisation + BN are one affine map of the byte (host-folded, fpool)
      mbar_wait(bar(D_FULL0), k & 1);
      tc_fence_after();
      float r[16];
      tmem_ld16(tm_lane, r);
      tc_fence_before();
#pragma unroll
      for (int o = 0; o < 13; ++o) r[o] = prelu_f(fmaf(r[o], p.fu8[o], p.f[16 + o]), p.f[32 + o]);
#pragma unroll
      {                                                // RGB order: R, then G and B of the (B, G) pair
        const float2 bgf = __half22float2(*reinterpret_cast<const __half2*>(&mx_cur.x));
        const float mxf[3] = {__half2float(__ushort_as_half((unsigned short)(mx_cur.y & 0xffffu))), bgf.y, bgf.x};
#pragma unroll
        for (int c = 0; c < 3; ++c)
          r[13 + c] = prelu_f(fmaf(mxf[c], p.fpool[c], p.fpool[3 + c]), p.f[32 + 13 + c]);
      }
      uint4* o = reinterpret_cast<uint4*>(p.out + (size_t)pix_cur * 16);
      o[0] = pack8(r);
      o[1] = pack8(r + 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
  }
}

}  // namespace BC_NS
using namespace BC_NS;

// B image of the uint8 form: U8_LIMBS fp16 limbs of  w_k s_c 2^e(o)  (27 columns) and of
// 2^e(o) sum_(taps of the column) sum_c w_(c,tap) t_c  (4 validity columns), rows = output channels, K order as
// the kernel builds A.
// w: [27][13] ((c*3+ky)*3+kx major, c in RGB order).  unscale[o] = 2^-e(o) (16 floats, 1 for the unused rows).
bool Umma<act_t>::initial_build_u8(uint8_t** out, const float* w, float* unscale) {
  const double mean[3] = {0.485, 0.456, 0.406}, sd[3] = {0.229, 0.224, 0.225};   // models.py:17-18
  std::vector<uint8_t> img(U8_LIMBS * U8_W, 0);
  for (int o = 0; o < 16; ++o) unscale[o] = 1.f;
  for (int o = 0; o < 13; ++o) {
    double col[U8_K] = {0.0};
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int tap = ky * 3 + kx;
        double t = 0.0;
        for (int c = 0; c < 3; ++c) {                        // window byte 3*kx + (2 - c): frames are BGR
          col[u8_k_byte(ky, 3 * kx + 2 - c)] = (double)w[(c * 9 + tap) * 13 + o] / (256.0 * sd[c]);
          t += (double)w[(c * 9 + tap) * 13 + o] * (-mean[c] / sd[c]);
        }
        col[u8_k_valid(ky, kx)] += t;
      }
    double mxv = 0.0;
    for (double v : col) mxv = std::max(mxv, std::fabs(v));
    int e = 0;
    if (mxv > 0.0) { int ex; std::frexp(mxv, &ex); e = 14 - ex; }      // largest entry in [2^13, 2^14)
    e = std::min(std::max(e, -40), 40);
    unscale[o] = (float)std::ldexp(1.0, -e);
    for (int k = 0; k < U8_K; ++k) {
      double v = std::ldexp(col[k], e);
      for (int limb = 0; limb < U8_LIMBS; ++limb) {
        const __half hb = __float2half_rn((float)v);
        v -= (double)__half2float(hb);
        memcpy(img.data() + limb * U8_W + swz<128>((uint32_t)(o * 128 + k * 2)), &hb, 2);
      }
    }
  }
  if (cudaMalloc(out, img.size()) != cudaSuccess) return false;
  return cudaMemcpy(*out, img.data(), img.size(), cudaMemcpyHostToDevice) == cudaSuccess;
}

// w: [27][13] ((c*3+ky)*3+kx major), fp32.  B rows = output channels (13 + 3 zero), K = 32.
bool Umma<act_t>::initial_build(uint8_t** out, const float* w) {
  std::vector<uint8_t> img(2 * INIT_W, 0);
  for (int o = 0; o < 13; ++o)
    for (int k = 0; k < 27; ++k) {
      float v = w[k * 13 + o];
      uint32_t u;
      memcpy(&u, &v, 4);
      u &= 0xffffe000u;
      float hi, lo;
      memcpy(&hi, &u, 4);
      lo = v - hi;
      const uint32_t off = swz<128>((uint32_t)(o * 128 + k * 4));
      memcpy(img.data() + off, &hi, 4);
      memcpy(img.data() + INIT_W + off, &lo, 4);
    }
  if (cudaMalloc(out, img.size()) != cudaSuccess) return false;
  return cudaMemcpy(*out, img.data(), img.size(), cudaMemcpyHostToDevice) == cudaSuccess;
}

cudaError_t Umma<act_t>::prepare_initial() {
  const int smem = INIT_SMEM + 1024;
  cudaError_t e = cudaFuncSetAttribute(k_umma_initial<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_umma_initial<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_umma_initial<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_umma_initial_u8, cudaFuncAttributeMaxDynamicSharedMemorySize, U8_SMEM + 1024);
  return e;
}

cudaError_t Umma<act_t>::launch_initial(const void* x, int kind, int B, act_t* out, int pool_kernel, const uint8_t* wblob,
                                        const uint8_t* wblob_u8,
                                        const float* u8_unscale, const float* lut, const float* g, const float* b, const float* a,
                                        int num_sms, cudaStream_t s) {
  InitParams p{};
  p.num_tiles = B * 256;
  p.x = x;
  p.out = out;
  p.wblob = (kind == 0 && ((uintptr_t)x & 3) == 0) ? wblob_u8 : wblob;
  p.lut = lut;
  p.pool2 = pool_kernel == 2;
  memcpy(p.f, g, 64);
  memcpy(p.f + 16, b, 64);
  memcpy(p.f + 32, a, 64);
  for (int o = 0; o < 16; ++o) p.fu8[o] = g[o] * u8_unscale[o];
  {
    const double mean[3] = {0.485, 0.456, 0.406}, sd[3] = {0.229, 0.224, 0.225};   // models.py:17-18,91
    for (int c = 0; c < 3; ++c) {
      p.fpool[c] = (float)((double)g[13 + c] / (256.0 * sd[c]));
      p.fpool[3 + c] = (float)((double)g[13 + c] * (-mean[c] / sd[c]) + (double)b[13 + c]);
    }
  }
  const int smem = INIT_SMEM + 1024;       // opt-in set per device by prepare_initial()
  if (kind == 0 && ((uintptr_t)x & 3) == 0) {          // the byte-window loads are 32-bit: a frame pointer that is not
    const int ctas8 = num_sms * U8_MINB;               // 4-byte aligned takes the byte-wise tf32 kernel below
    k_umma_initial_u8<<<p.num_tiles < ctas8 ? p.num_tiles : ctas8, 160, U8_SMEM + 1024, s>>>(p);
    return cudaGetLastError();
  }
  const int ctas = num_sms * INIT_MINB;
  const int grid = p.num_tiles < ctas ? p.num_tiles : ctas;
  if (kind == 0) k_umma_initial<0><<<grid, 160, smem, s>>>(p);
  else if (kind == 1) k_umma_initial<1><<<grid, 160, smem, s>>>(p);
  else k_umma_initial<2><<<grid, 160, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace bc
