// First half of an ENet down-sampling bottleneck on tcgen05 (fp16 / bf16 operands, fp32 accumulation):
//
//   x (full resolution, NHWC) --4 strided TMA box loads (the 2x2 window positions)--> smem
//        main: max over the four tap tiles + 2-bit argmax (first maximum wins)  -> pooled, idx (global)
//        ext : conv 2x2 stride 2 = 4 taps x (CIN/16) tcgen05.mma                -> D [128 x 16] (TMEM)
//              epilogue: +bias, PReLU, bf16                                     -> e1 (global, 16 wide,
//                                                                                  zero padded from CI)
//
// One tile = 128 output pixels (whole output rows); tap (dy, dx) of the window is ONE 5-D TMA box
// over the view [n*Ho][dy][Wo][dx][CIN] of the input, so the tile that feeds the tensor core is
// the same shared-memory image the pooling threads read: x leaves HBM exactly once.  The second
// half (3x3 conv + expansion + pooled residual + next projection) is k_umma_bottleneck's narrow
// residual variant (enet_umma.cu).
//
// A CTA works on one tile at a time (one tap set, double-buffered accumulator); overlap comes
// from MINB co-resident CTAs per SM.  Warp roles: 0 TMA producer, 1 MMA issuer, 2-5 pooling +
// epilogue (one output pixel = one TMEM lane per thread).
// Semantics: oracle/enet_oracle.py `down` (the frozen graph the reference runs, models.py:43-44).
#include "umma_common.cuh"

#include <cstring>

namespace bc {
namespace BC_NS {

struct DownParams {
  int num_tiles;          // 128-output-pixel tiles
  int reverse;            // 1: walk the tiles from the last to the first (L2 reuse between consecutive kernels, enet_umma.cu)
  int rows_per_tile;      // output rows per tile (128 / Wo, at least 1)
  act_t* pooled;          // [out px][CIN]
  uint8_t* idx;           // [out px][CIN] window position of the maximum
  act_t* e1;              // [out px][16]
  const uint8_t* wblob;   // [4 taps][16 rows][CIN] bf16, K-major swizzled rows
  float f[32];            // bias[16], slope[16] (by value: constant-bank operands)
};

template <int CIN>
struct DownSmem {
  static constexpr int RB = CIN * 2;                  // bytes per input pixel = swizzle span (32 / 128)
  static constexpr int TAP = 128 * RB;                // one window position of the tile
  static constexpr int WTAP = 16 * RB;                // one tap of the weights
  static constexpr int OFF_TAPS = 0;
  static constexpr int OFF_W = 4 * TAP;
  static constexpr int OFF_BAR = OFF_W + ((4 * WTAP + 1023) / 1024) * 1024;
  static constexpr int TOTAL = OFF_BAR + 128;
  static constexpr int MINB = CIN == 64 ? 3 : 4;
  static_assert((TOTAL + 2048) * MINB <= 233472, "shared memory budget");
};

// packed 16-bit x 2 max with argmax update: where b > a (strictly, per 16-bit lane) take b and tap t
__device__ __forceinline__ void max_arg2(uint32_t& best, uint32_t& bi, uint32_t v, uint32_t t2) {
  const uint32_t m = gt2_mask(v, best);               // 0xFFFF per lane where b > a
  best = (v & m) | (best & ~m);
  bi = (t2 & m) | (bi & ~m);
}

template <int CIN>
__global__ void __launch_bounds__(192, DownSmem<CIN>::MINB)
k_umma_down(const __grid_constant__ CUtensorMap map_x,   // 5D [n*Ho][2][Wo][2][CIN], box [rows][1][boxW][1][CIN]
            const __grid_constant__ DownParams p) {
  using S = DownSmem<CIN>;
  constexpr int RB = S::RB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  enum { TAP_FULL = 0, TAP_EMPTY, D_FULL0, D_FULL1, D_EMPTY0, D_EMPTY1, W_FULL, NBARS };
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  uint32_t* tmem_slot = (uint32_t*)&bars[NBARS];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index made provably warp-uniform: role branches become uniform branches

  if (tid == 0) {
    mbar_init(bar(TAP_FULL), 1);
    mbar_init(bar(TAP_EMPTY), 129);                     // the MMAs' commit + every pooling thread
    mbar_init(bar(D_FULL0), 1);
    mbar_init(bar(D_FULL1), 1);
    mbar_init(bar(D_EMPTY0), 128);
    mbar_init(bar(D_EMPTY1), 128);
    mbar_init(bar(W_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar(W_FULL), 4 * S::WTAP);
    bulk_load(sbase + S::OFF_W, p.wblob, 4 * S::WTAP, bar(W_FULL));
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    for (int k = 0; k < T; ++k) {
      const int tile = p.reverse ? p.num_tiles - 1 - ((int)blockIdx.x + k * (int)gridDim.x) : (int)blockIdx.x + k * (int)gridDim.x;
      if (k >= 1) mbar_wait(bar(TAP_EMPTY), (k - 1) & 1);
      mbar_expect_tx_e(bar(TAP_FULL), 4 * S::TAP);
#pragma unroll
      for (int t = 0; t < 4; ++t)
        tma_load_5d_e(sbase + S::OFF_TAPS + t * S::TAP, &map_x, 0, t & 1, 0, t >> 1, tile * p.rows_per_tile, bar(TAP_FULL));
    }
  } else if (warp == 1) {
    mbar_wait(bar(W_FULL), 0);
    for (int k = 0; k < T; ++k) {
      const int b = k & 1;
      mbar_wait(bar(TAP_FULL), k & 1);
      if (k >= 2) mbar_wait(bar(D_EMPTY0 + b), ((k >> 1) - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int kk = 0; kk < CIN / 16; ++kk)
          umma_mma_e(tmem + b * 16, smem_desc<RB>(sbase + S::OFF_TAPS + t * S::TAP + kk * 32),
                    smem_desc<RB>(sbase + S::OFF_W + t * S::WTAP + kk * 32), instr_desc(128, 16), (t | kk) != 0);
      umma_commit_e(bar(TAP_EMPTY));
      umma_commit_e(bar(D_FULL0 + b));
    }
  } else {
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const uint32_t tm_lane = tmem + ((uint32_t)(q4 * 32) << 16);
    for (int k = 0; k < T; ++k) {
      const int tile = p.reverse ? p.num_tiles - 1 - ((int)blockIdx.x + k * (int)gridDim.x) : (int)blockIdx.x + k * (int)gridDim.x;
      const int b = k & 1;
      const size_t px = (size_t)tile * 128 + m;
      // ---- main branch: max-pool with argmax straight from the tap tiles
      mbar_wait(bar(TAP_FULL), k & 1);
#pragma unroll
      for (int c = 0; c < CIN / 8; ++c) {
        const uint32_t off = swz<RB>((uint32_t)(m * RB + c * 16));
        uint4 best = *reinterpret_cast<const uint4*>(smem + S::OFF_TAPS + off);
        uint32_t i0 = 0, i1 = 0, i2 = 0, i3 = 0;
#pragma unroll
        for (int t = 1; t < 4; ++t) {
          const uint4 v = *reinterpret_cast<const uint4*>(smem + S::OFF_TAPS + t * S::TAP + off);
          const uint32_t t2 = (uint32_t)t * 0x00010001u;
          max_arg2(best.x, i0, v.x, t2);
          max_arg2(best.y, i1, v.y, t2);
          max_arg2(best.z, i2, v.z, t2);
          max_arg2(best.w, i3, v.w, t2);
        }
        *reinterpret_cast<uint4*>(p.pooled + px * CIN + c * 8) = best;
        *reinterpret_cast<uint2*>(p.idx + px * CIN + c * 8) =
            make_uint2(__byte_perm(i0, i1, 0x6420), __byte_perm(i2, i3, 0x6420));
      }
      mbar_arrive(bar(TAP_EMPTY));
      // ---- extension branch: e1 = PReLU(conv2x2 + b), 16 channels (zero beyond CI)
      mbar_wait(bar(D_FULL0 + b), (k >> 1) & 1);
      tc_fence_after();
      float v[16];
      tmem_ld16(tm_lane + b * 16, v);
      tc_fence_before();
      mbar_arrive(bar(D_EMPTY0 + b));
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = prelu_f(v[j] + p.f[j], p.f[16 + j]);
      uint4* o = reinterpret_cast<uint4*>(p.e1 + px * 16);
      o[0] = pack8(v);
      o[1] = pack8(v + 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
  }
}

// ------------------------------------------------------------------------ host side
// w: folded conv [tap = ky*2+kx][cin][ci]; rows beyond ci stay zero (bias 0, slope 1 -> e1 = 0)
}  // namespace BC_NS
using namespace BC_NS;

bool Umma<act_t>::down_build(UmmaPack& out, int cin, int ci, const float* w, const float* bias, const float* alpha) {
  if ((cin != 16 && cin != 64) || ci > 16) return false;
  const int rb = cin * 2, wtap = 16 * rb;
  std::vector<uint8_t> img(4 * wtap, 0);
  for (int t = 0; t < 4; ++t)
    for (int o = 0; o < ci; ++o)
      for (int k = 0; k < cin; ++k) {
        const uint16_t h = host_act_bits(w[((size_t)t * cin + k) * ci + o]);
        uint32_t off = (uint32_t)(o * rb + k * 2);
        memcpy(img.data() + t * wtap + (rb == 128 ? swz<128>(off) : swz<32>(off)), &h, 2);
      }
  out.hf.assign(32, 0.f);
  for (int j = 0; j < 16; ++j) { out.hf[j] = j < ci ? bias[j] : 0.f; out.hf[16 + j] = j < ci ? alpha[j] : 1.f; }
  if (cudaMalloc(&out.wblob, img.size()) != cudaSuccess) return false;
  cudaMemcpy(out.wblob, img.data(), img.size(), cudaMemcpyHostToDevice);
  out.C = cin; out.CI = 16; out.ntaps = 4;
  return true;
}

template <int CIN>
static cudaError_t down_launch_t(const UmmaPack& pk, const act_t* x, act_t* pooled, uint8_t* idx, act_t* e1, int n, int Ho, int Wo,
                                 int num_sms, cudaStream_t s) {
  using S = DownSmem<CIN>;
  const int rows = 128 / Wo > 0 ? 128 / Wo : 1, box_w = 128 / rows;
  if (Wo % box_w != 0 || Wo != box_w || Ho % rows != 0) return cudaErrorInvalidValue;   // a tile = whole output rows
  CUtensorMap mx;
  if (!make_map_window(&mx, x, n, Ho, Wo, CIN, rows, box_w)) return cudaErrorInvalidValue;
  DownParams p{};
  p.num_tiles = n * Ho * Wo / 128;
  p.reverse = g_umma_reverse;
  p.rows_per_tile = rows;
  p.pooled = pooled;
  p.idx = idx;
  p.e1 = e1;
  p.wblob = pk.wblob;
  memcpy(p.f, pk.hf.data(), 32 * sizeof(float));
  const int smem = S::TOTAL + 1024;        // opt-in set per device by prepare_down()
  const int ctas = num_sms * S::MINB;
  k_umma_down<CIN><<<p.num_tiles < ctas ? p.num_tiles : ctas, 192, smem, s>>>(mx, p);
  return cudaGetLastError();
}

cudaError_t Umma<act_t>::prepare_down() {
  cudaError_t e = cudaFuncSetAttribute(k_umma_down<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DownSmem<64>::TOTAL + 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_umma_down<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DownSmem<16>::TOTAL + 1024);
  return e;
}

cudaError_t Umma<act_t>::launch_down(const UmmaPack& pk, const act_t* x, act_t* pooled, uint8_t* idx, act_t* e1, int n, int Ho,
                                     int Wo, int num_sms, cudaStream_t s) {
  if (pk.C == 64) return down_launch_t<64>(pk, x, pooled, idx, e1, n, Ho, Wo, num_sms, s);
  if (pk.C == 16) return down_launch_t<16>(pk, x, pooled, idx, e1, n, Ho, Wo, num_sms, s);
  return cudaErrorInvalidValue;
}

}  // namespace bc
