// ENet upsampling bottleneck on tcgen05 (fp16 / bf16 operands, fp32 accumulation), one kernel:
//
//   x tile (128 low-res pixels, TMA) --G1--> [ main = Wm x | e1 = W1 x ]         D_a [128 x (COUT+CI)]
//        e1: +bias, act, bf16 -> smem             --G2--> 4 taps of the 2x2 stride-2 transposed conv  D_b [128 x 4 CI]
//        e2_t: +bias, act, bf16 -> smem (4 tiles) --G3--> e3_t = W3 e2_t          D_c [4][128 x COUT]
//        out(2y+ky, 2x+kx) = act_out(e3_t + b3 + (pool index == t ? main + bm : 0))   (max-unpool)
//          -> bf16 -> smem, staged as whole high-resolution rows -> TMA store
//        (optional) the next block's 1x1 projection on the staged rows  --G4--> e1' -> global
//
// One TMEM lane = one low-resolution pixel; the four output pixels of a thread differ only in
// which quarter of D_c they read.  Semantics: oracle/enet_oracle.py `up` (the upsampling
// bottlenecks of the frozen graph the reference runs, models.py:43-44).
#include "umma_common.cuh"

#include <cstring>
#include <type_traits>

namespace bc {
namespace BC_NS {

struct UpParams {
  int num_tiles;          // low-resolution 128-pixel tiles
  int reverse;            // 1: walk the tiles from the last to the first (L2 reuse between consecutive kernels, enet_umma.cu)
  int tiles_per_frame;    // Hl * Wl / 128
  int Wl;                 // low-resolution width (64 or 128)
  int has_next;
  const uint8_t* idx;     // [low px][COUT] pool window position (2 bits)
  act_t* e1_next;         // [high px][16]
  const uint8_t* wblob;
  // bm[COUT] b1[CI] a1[CI] bt[CI] at[CI] b3[COUT] aout[COUT] b1n[16] a1n[16], by value: constant-bank
  // operands of the epilogue arithmetic (compile-time channel indices), no shared-memory traffic
  float f[384];
};

#ifndef BC_UP5_MINB
#define BC_UP5_MINB 4
#endif
template <int CIN, int CI, int COUT>
struct UpSmem {
  static constexpr int RB = CI * 2;                    // row bytes of CI-wide operands
  static constexpr int ORB = COUT * 2;                 // bytes per output pixel (128 or 32)
  static constexpr int NSUB = CIN / 64;
  static constexpr int XSUB = 128 * 128;
  static constexpr int XBUF = NSUB * XSUB;
  static constexpr int N1 = COUT + CI;
  static constexpr int E_TILE = 128 * RB;
  static constexpr int OUT_BYTES = 512 * ORB;          // 512 high-res pixels per tile
  static constexpr int OUT_ROW = COUT == 64 ? 128 * ORB : 256 * ORB;   // one staged high-res row
  static constexpr int NROWS = OUT_BYTES / OUT_ROW;    // 4 (up4) or 2 (up5)
  static constexpr int B1_SUB = N1 * 128;
  static constexpr int WT_BYTES = 4 * CI * RB;
  static constexpr int W3_BYTES = COUT * RB;
  static constexpr int W1N_BYTES = 16 * 128;           // next projection [16][64] (COUT == 64 only)
  // upsample5_0 (COUT = 16) is small enough for four CTAs per SM with one x buffer, D_c aliased onto D_b (dead
  // once epilogue B has read it) and the staged output rows aliased onto the e2 tiles (dead once G3 has read them;
  // epilogue B of the next tile waits for the TMA store to have read the rows); upsample4_0 fills the SM with one CTA
  static constexpr int MINB = COUT == 16 ? BC_UP5_MINB : 1;
  // upsample4_0 keeps ONE tile in flight per SM, so its four epilogues are the critical path: two warps per TMEM
  // lane quarter, each taking half of the columns / taps / rows of every epilogue
  static constexpr int EPW = COUT == 64 ? 2 : 1;
  static constexpr int THREADS = 64 + 128 * EPW;
  static constexpr int NXB = COUT == 16 ? 1 : 2;       // x buffers
  static constexpr int OFF_X = 0;
  static constexpr int OFF_E1 = OFF_X + NXB * XBUF;
  static constexpr int OFF_E2 = OFF_E1 + ((E_TILE + 1023) / 1024) * 1024;
  static constexpr bool OUT_ON_E2 = COUT == 16 && MINB >= 4;
  static_assert(!OUT_ON_E2 || OUT_BYTES <= 4 * E_TILE, "staged rows must fit over the e2 tiles");
  static constexpr int OFF_OUT = OUT_ON_E2 ? OFF_E2 : OFF_E2 + 4 * E_TILE;
  static constexpr int OFF_W = OUT_ON_E2 ? OFF_E2 + 4 * E_TILE : OFF_OUT + OUT_BYTES;
  static constexpr int OFF_B1 = OFF_W;
  static constexpr int OFF_WT = OFF_B1 + ((NSUB * B1_SUB + 1023) / 1024) * 1024;
  static constexpr int OFF_W3 = OFF_WT + ((WT_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_W1N = OFF_W3 + ((W3_BYTES + 1023) / 1024) * 1024;
  static constexpr int W_BYTES = OFF_W1N - OFF_W + W1N_BYTES;
  static constexpr int OFF_F = OFF_W + ((W_BYTES + 1023) / 1024) * 1024;
  static constexpr int NF = 3 * COUT + 4 * CI + 32;
  static constexpr int OFF_BAR = OFF_F + ((NF * 4 + 63) / 64) * 64;
  static constexpr int TOTAL = OFF_BAR + 192;
  static constexpr uint32_t COL_A = 0, COL_B = N1, COL_C = COUT == 16 ? COL_B : N1 + 4 * CI, COL_D = COL_B;
  static constexpr uint32_t COLS_USED = COL_C + 4 * COUT;
  static constexpr uint32_t TMEM_COLS = COLS_USED <= 128 ? 128 : COLS_USED <= 256 ? 256 : 512;
  static_assert(COUT != 16 || 4 * COUT <= 4 * CI, "D_c must fit over D_b");
};

template <int CIN, int CI, int COUT>
__global__ void __launch_bounds__((UpSmem<CIN, CI, COUT>::THREADS), (UpSmem<CIN, CI, COUT>::MINB))
k_umma_up(const __grid_constant__ CUtensorMap map_x,   // 2D [low px][CIN], box [128][64], 128-byte swizzle
          const __grid_constant__ CUtensorMap map_y,   // 2D [high px][COUT], box = one staged row
          const __grid_constant__ UpParams p) {
  using S = UpSmem<CIN, CI, COUT>;
  constexpr int RB = S::RB, ORB = S::ORB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer: LDS/STS, not generic LD/ST
  const uint32_t sbase = smem_u32(smem);
  constexpr int F_BM = 0, F_B1 = COUT, F_A1 = F_B1 + CI, F_BT = F_A1 + CI, F_AT = F_BT + CI, F_B3 = F_AT + CI,
                F_AOUT = F_B3 + COUT, F_B1N = F_AOUT + COUT, F_A1N = F_B1N + 16;
  static_assert(F_A1N + 16 <= 384, "parameter block");
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  enum { X_FULL0 = 0, X_FULL1, X_EMPTY0, X_EMPTY1, DA_FULL, E1_FULL, DB_FULL, E2_FULL, DC_FULL, OUT_FULL, OUT_EMPTY,
         DD_FULL, W_FULL, NBARS };
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  uint32_t* tmem_slot = (uint32_t*)&bars[NBARS];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index made provably warp-uniform: role branches become uniform branches

  if (tid == 0) {
    const int one[] = {X_FULL0, X_FULL1, X_EMPTY0, X_EMPTY1, DA_FULL, DB_FULL, DC_FULL, OUT_EMPTY, DD_FULL, W_FULL};
    for (int b : one) mbar_init(bar(b), 1);
    const int all[] = {E1_FULL, E2_FULL, OUT_FULL};
    for (int b : all) mbar_init(bar(b), 128 * S::EPW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar(W_FULL), S::W_BYTES);
    bulk_load(sbase + S::OFF_W, p.wblob, S::W_BYTES, bar(W_FULL));
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(S::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int rows_lo = 128 / p.Wl;                       // low-res rows per tile (2 or 1)
  const int Wh = 2 * p.Wl;                              // high-res width

  if (warp == 0) {
    for (int k = 0; k < T; ++k) {
      const int tile = p.reverse ? p.num_tiles - 1 - ((int)blockIdx.x + k * (int)gridDim.x) : (int)blockIdx.x + k * (int)gridDim.x;
      const int b = k % S::NXB;
      if (k >= S::NXB) mbar_wait(bar(X_EMPTY0 + b), ((k / S::NXB) - 1) & 1);
      mbar_expect_tx_e(bar(X_FULL0 + b), S::XBUF);
      for (int s = 0; s < S::NSUB; ++s)
        tma_load_2d_e(sbase + S::OFF_X + b * S::XBUF + s * S::XSUB, &map_x, s * 64, tile * 128, bar(X_FULL0 + b));
    }
  } else if (warp == 1) {
    mbar_wait(bar(W_FULL), 0);
    for (int k = 0; k < T; ++k) {
      const int b = k % S::NXB;
      // G1: [main | e1] = x * B1^T
      mbar_wait(bar(X_FULL0 + b), (k / S::NXB) & 1);
      if (k >= 1) mbar_wait(bar(OUT_FULL), (k - 1) & 1);       // epilogue done reading main of tile k-1
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < CIN / 16; ++kk)
        umma_mma_e(tmem + S::COL_A, smem_desc<128>(sbase + S::OFF_X + b * S::XBUF + (kk / 4) * S::XSUB + (kk % 4) * 32),
                  smem_desc<128>(sbase + S::OFF_B1 + (kk / 4) * S::B1_SUB + (kk % 4) * 32), instr_desc(128, S::N1), kk != 0);
      umma_commit_e(bar(X_EMPTY0 + b));
      umma_commit_e(bar(DA_FULL));
      // G2: the four transposed-conv taps
      mbar_wait(bar(E1_FULL), k & 1);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < CI / 16; ++kk)
        umma_mma_e(tmem + S::COL_B, smem_desc<RB>(sbase + S::OFF_E1 + kk * 32), smem_desc<RB>(sbase + S::OFF_WT + kk * 32),
                  instr_desc(128, 4 * CI), kk != 0);
      umma_commit_e(bar(DB_FULL));
      // G3: expansion of every tap
      mbar_wait(bar(E2_FULL), k & 1);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int kk = 0; kk < CI / 16; ++kk)
          umma_mma_e(tmem + S::COL_C + t * COUT, smem_desc<RB>(sbase + S::OFF_E2 + t * S::E_TILE + kk * 32),
                    smem_desc<RB>(sbase + S::OFF_W3 + kk * 32), instr_desc(128, COUT), kk != 0);
      umma_commit_e(bar(DC_FULL));
      // G4: next block's projection on the staged high-res rows (COUT == 64: one row = one M tile)
      if constexpr (COUT == 64) {
        if (p.has_next) {
          mbar_wait(bar(OUT_FULL), k & 1);
          tc_fence_after();
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_mma_e(tmem + S::COL_D + r * 16, smem_desc<128>(sbase + S::OFF_OUT + r * S::OUT_ROW + kk * 32),
                        smem_desc<128>(sbase + S::OFF_W1N + kk * 32), instr_desc(128, 16), kk != 0);
          umma_commit_e(bar(DD_FULL));
        }
      }
    }
  } else {
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const uint32_t tm_lane = tmem + ((uint32_t)(q4 * 32) << 16);
    const bool storer = (warp == 2 && lane == 0);
    const int eh = (warp - 2) >> 2;                       // which half of every epilogue this warp takes (EPW == 2)
    const int lr = m / p.Wl, lx = m % p.Wl;             // position of my low-res pixel inside the tile
    mbar_wait(bar(W_FULL), 0);
    for (int k = 0; k < T; ++k) {
      const int tile = p.reverse ? p.num_tiles - 1 - ((int)blockIdx.x + k * (int)gridDim.x) : (int)blockIdx.x + k * (int)gridDim.x;
      // ---- E_A: e1 = act(proj + b1) -> smem
      mbar_wait(bar(DA_FULL), k & 1);
      tc_fence_after();
      if constexpr (S::EPW == 2) {
        static_assert(S::EPW == 1 || CI == 32, "split epilogue A assumes 32 internal channels");
        auto ep_a = [&](auto H) {                       // compile-time half: the parameters stay constant-bank operands
          constexpr int h = decltype(H)::value;
          float v[16];
          tmem_ld16(tm_lane + S::COL_A + COUT + 16 * h, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = prelu_f(v[j] + p.f[F_B1 + 16 * h + j], p.f[F_A1 + 16 * h + j]);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            *reinterpret_cast<uint4*>(smem + S::OFF_E1 + swz<RB>(m * RB + (2 * h + c) * 16)) =
                pack8(v + 8 * c);
        };
        if (eh == 0) ep_a(std::integral_constant<int, 0>{}); else ep_a(std::integral_constant<int, 1>{});
      } else {
        float v[CI];
        if constexpr (CI == 32) tmem_ld32(tm_lane + S::COL_A + COUT, v); else tmem_ld16(tm_lane + S::COL_A + COUT, v);
#pragma unroll
        for (int j = 0; j < CI; ++j) v[j] = prelu_f(v[j] + p.f[F_B1 + j], p.f[F_A1 + j]);
#pragma unroll
        for (int c = 0; c < CI / 8; ++c)
          *reinterpret_cast<uint4*>(smem + S::OFF_E1 + swz<RB>(m * RB + c * 16)) =
              pack8(v + 8 * c);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(E1_FULL));
      // ---- E_B: e2_t = act(tconv tap + bt) -> 4 smem tiles
      mbar_wait(bar(DB_FULL), k & 1);
      if constexpr (S::OUT_ON_E2) {                     // the previous tile's staged rows lie here until the TMA store has read them
        if (k >= 1) mbar_wait(bar(OUT_EMPTY), (k - 1) & 1);
      }
      tc_fence_after();
#pragma unroll 1
      for (int t = (S::EPW == 2 ? 2 * eh : 0); t < (S::EPW == 2 ? 2 * eh + 2 : 4); ++t) {
        float v[CI];
        if constexpr (CI == 32) tmem_ld32(tm_lane + S::COL_B + t * CI, v); else tmem_ld16(tm_lane + S::COL_B + t * CI, v);
#pragma unroll
        for (int j = 0; j < CI; ++j) v[j] = prelu_f(v[j] + p.f[F_BT + j], p.f[F_AT + j]);
#pragma unroll
        for (int c = 0; c < CI / 8; ++c)
          *reinterpret_cast<uint4*>(smem + S::OFF_E2 + t * S::E_TILE + swz<RB>(m * RB + c * 16)) =
              pack8(v + 8 * c);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(E2_FULL));
      // ---- E_C: out = act_out(e3_t + b3 + unpooled main) -> staged high-res rows
      mbar_wait(bar(DC_FULL), k & 1);
      if (k >= 1) mbar_wait(bar(OUT_EMPTY), (k - 1) & 1);
      tc_fence_after();
      const uint8_t* ip = p.idx + ((size_t)tile * 128 + m) * COUT;
#pragma unroll(COUT == 64 ? 1 : 4)   // 64 channels: a rolled loop (dynamic constant-bank index) beats 4x the code
      for (int c0 = (S::EPW == 2 ? eh * (COUT / 2) : 0); c0 < (S::EPW == 2 ? (eh + 1) * (COUT / 2) : COUT); c0 += 16) {
        float mainv[16];
        tmem_ld16(tm_lane + S::COL_A + c0, mainv);
        const uint4 iv = *reinterpret_cast<const uint4*>(ip + c0);
        const uint8_t* ib = reinterpret_cast<const uint8_t*>(&iv);
#pragma unroll
        for (int j = 0; j < 16; ++j) mainv[j] += p.f[F_BM + c0 + j];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float v[16];
          tmem_ld16(tm_lane + S::COL_C + t * COUT + c0, v);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float o0 = prelu_f(v[j] + p.f[F_B3 + c0 + j] + (ib[j] == t ? mainv[j] : 0.f), p.f[F_AOUT + c0 + j]);
            float o1 = prelu_f(v[j + 1] + p.f[F_B3 + c0 + j + 1] + (ib[j + 1] == t ? mainv[j + 1] : 0.f), p.f[F_AOUT + c0 + j + 1]);
            pk[j / 2] = pack_act(o0, o1);
          }
          // high-res pixel (2*lr + ky, 2*lx + kx) of this tile
          const int hr = 2 * lr + (t >> 1), hx = 2 * lx + (t & 1);
          uint8_t* orow = smem + S::OFF_OUT + (COUT == 64 ? hr * S::OUT_ROW : (t >> 1) * S::OUT_ROW);
          const uint32_t off = (uint32_t)(hx * ORB + c0 * 2);
          *reinterpret_cast<uint4*>(orow + swz<(COUT == 64 ? 128 : 32)>(off)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(orow + swz<(COUT == 64 ? 128 : 32)>(off + 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(OUT_FULL));
      if (storer) {
        mbar_wait(bar(OUT_FULL), k & 1);
        const int n = tile / p.tiles_per_frame;
        const int y0 = (tile % p.tiles_per_frame) * rows_lo;          // first low-res row of the tile
        const int hrow0 = n * (p.tiles_per_frame * rows_lo * 2) + 2 * y0;   // global high-res row index
        for (int r = 0; r < S::NROWS; ++r)
          tma_store_2d(&map_y, sbase + S::OFF_OUT + r * S::OUT_ROW, 0, (hrow0 + r) * Wh);
        tma_store_commit();
      }
      // ---- E_D: next block's projection
      if constexpr (COUT == 64) {
        if (p.has_next) {
          mbar_wait(bar(DD_FULL), k & 1);
          tc_fence_after();
          const int n = tile / p.tiles_per_frame;
          const int y0 = (tile % p.tiles_per_frame) * rows_lo;
          const size_t hrow0 = (size_t)n * (p.tiles_per_frame * rows_lo * 2) + 2 * y0;
#pragma unroll 1
          for (int r = (S::EPW == 2 ? 2 * eh : 0); r < (S::EPW == 2 ? 2 * eh + 2 : 4); ++r) {
            float v[16];
            tmem_ld16(tm_lane + S::COL_D + r * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = prelu_f(v[j] + p.f[F_B1N + j], p.f[F_A1N + j]);
            uint4* o = reinterpret_cast<uint4*>(p.e1_next + ((hrow0 + r) * Wh + m) * 16);
            o[0] = pack8(v);
            o[1] = pack8(v + 8);
          }
          tc_fence_before();
        }
      }
      if (storer) {
        tma_store_wait_read();
        mbar_arrive(bar(OUT_EMPTY));
      }
    }
    if (storer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(S::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------ host side
static void put_rows(uint8_t* dst, int rows, int row_bytes, int sw, const float* w, size_t stride_row, size_t stride_k,
                     int k0) {
  // element (row r, k) = w[r * stride_row + (k0 + k) * stride_k]
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < row_bytes / 2; ++k) {
      const uint16_t h = host_act_bits(w[(size_t)r * stride_row + (size_t)(k0 + k) * stride_k]);
      uint32_t off = (uint32_t)(r * row_bytes + k * 2);
      uint32_t so = sw == 128 ? swz<128>(off) : sw == 64 ? swz<64>(off) : swz<32>(off);
      memcpy(dst + so, &h, 2);
    }
}

template <int CIN, int CI, int COUT>
static bool up_build_t(UmmaPack& out, const float* wm, const float* bmv, const float* w1, const float* b1v, const float* a1v,
                       const float* wt, const float* btv, const float* atv, const float* w3, const float* b3v,
                       const float* aoutv, const float* w1n, const float* b1nv, const float* a1nv) {
  using S = UpSmem<CIN, CI, COUT>;
  std::vector<uint8_t> img(S::W_BYTES, 0);
  // folded layouts are [tap][cin][cout]: element (out o, in k) = w[k * cout + o]
  for (int s = 0; s < S::NSUB; ++s) {
    uint8_t* d = img.data() + (S::OFF_B1 - S::OFF_W) + s * S::B1_SUB;
    put_rows(d, COUT, 128, 128, wm, 1, COUT, s * 64);                       // rows 0..COUT-1: main conv
    put_rows(d + COUT * 128, CI, 128, 128, w1, 1, CI, s * 64);              // rows COUT..: projection
  }
  for (int t = 0; t < 4; ++t)                                                // rows t*CI + j
    put_rows(img.data() + (S::OFF_WT - S::OFF_W) + t * CI * S::RB, CI, S::RB, S::RB, wt + (size_t)t * CI * CI, 1, CI, 0);
  put_rows(img.data() + (S::OFF_W3 - S::OFF_W), COUT, S::RB, S::RB, w3, 1, COUT, 0);
  if (w1n) put_rows(img.data() + (S::OFF_W1N - S::OFF_W), 16, 128, 128, w1n, 1, 16, 0);
  std::vector<float> f(S::NF, 0.f);
  float* o = f.data();
  memcpy(o, bmv, COUT * 4); o += COUT;
  memcpy(o, b1v, CI * 4); o += CI;
  memcpy(o, a1v, CI * 4); o += CI;
  memcpy(o, btv, CI * 4); o += CI;
  memcpy(o, atv, CI * 4); o += CI;
  memcpy(o, b3v, COUT * 4); o += COUT;
  memcpy(o, aoutv, COUT * 4); o += COUT;
  if (w1n) { memcpy(o, b1nv, 16 * 4); memcpy(o + 16, a1nv, 16 * 4); }
  if (cudaMalloc(&out.wblob, S::W_BYTES) != cudaSuccess) return false;
  cudaMemcpy(out.wblob, img.data(), S::W_BYTES, cudaMemcpyHostToDevice);
  out.hf = f;
  out.C = CIN; out.CI = CI; out.ntaps = 4; out.has_exp = true; out.has_next = w1n != nullptr;
  return true;
}

template <int CIN, int CI, int COUT>
static cudaError_t up_launch_t(const UmmaPack& pk, const act_t* x, const uint8_t* idx, act_t* y, act_t* e1_next, int n, int Hl,
                               int Wl, int has_next, int num_sms, cudaStream_t s);
}  // namespace BC_NS
using namespace BC_NS;

// NB: a row of put_rows' source walks `stride_row` = 1 (out channel contiguous in [cin][cout]).
bool Umma<act_t>::up_build(UmmaPack& out, int cin, int ci, int cout, const float* wm, const float* bm, const float* w1, const float* b1,
              const float* a1, const float* wt, const float* bt, const float* at, const float* w3, const float* b3,
              const float* aout, const float* w1n, const float* b1n, const float* a1n) {
  if (cin == 128 && ci == 32 && cout == 64)
    return up_build_t<128, 32, 64>(out, wm, bm, w1, b1, a1, wt, bt, at, w3, b3, aout, w1n, b1n, a1n);
  if (cin == 64 && ci == 16 && cout == 16)
    return up_build_t<64, 16, 16>(out, wm, bm, w1, b1, a1, wt, bt, at, w3, b3, aout, nullptr, nullptr, nullptr);
  return false;
}

namespace BC_NS {
template <int CIN, int CI, int COUT>
static cudaError_t up_launch_t(const UmmaPack& pk, const act_t* x, const uint8_t* idx, act_t* y, act_t* e1_next, int n, int Hl,
                               int Wl, int has_next, int num_sms, cudaStream_t s) {
  using S = UpSmem<CIN, CI, COUT>;
  CUtensorMap mx, my;
  const size_t lpx = (size_t)n * Hl * Wl;
  if (!make_map_x(&mx, x, lpx, CIN)) return cudaErrorInvalidValue;
  if (COUT == 64) { if (!make_map_x(&my, y, lpx * 4, 64)) return cudaErrorInvalidValue; }
  else if (!make_map_rows(&my, y, lpx * 4, COUT, 256, 32)) return cudaErrorInvalidValue;
  UpParams p{};
  p.num_tiles = (int)(lpx / 128);
  p.reverse = g_umma_reverse;
  p.tiles_per_frame = Hl * Wl / 128;
  p.Wl = Wl;
  p.has_next = has_next;
  p.idx = idx;
  p.e1_next = e1_next;
  p.wblob = pk.wblob;
  memcpy(p.f, pk.hf.data(), pk.hf.size() * sizeof(float));
  const int smem = S::TOTAL + 1024;        // opt-in set per device by prepare_up()
  static_assert((S::TOTAL + 2048) * S::MINB <= 233472 && S::TMEM_COLS * S::MINB <= 512, "CTAs per SM");
  const int ctas = num_sms * S::MINB;                    // upsample5_0 fits three times per SM: three tiles in flight
  int grid = p.num_tiles < ctas ? p.num_tiles : ctas;
  k_umma_up<CIN, CI, COUT><<<grid, S::THREADS, smem, s>>>(mx, my, p);
  return cudaGetLastError();
}

}  // namespace BC_NS

cudaError_t Umma<act_t>::prepare_up() {
  cudaError_t e = cudaFuncSetAttribute(k_umma_up<128, 32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       UpSmem<128, 32, 64>::TOTAL + 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_umma_up<64, 16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpSmem<64, 16, 16>::TOTAL + 1024);
  return e;
}

cudaError_t Umma<act_t>::launch_up(const UmmaPack& pk, int cin, int cout, const act_t* x, const uint8_t* idx, act_t* y,
                                   act_t* e1_next, int n, int Hl, int Wl, int has_next, int num_sms, cudaStream_t s) {
  if (cin == 128 && cout == 64 && Wl == 64) return up_launch_t<128, 32, 64>(pk, x, idx, y, e1_next, n, Hl, Wl, has_next, num_sms, s);
  if (cin == 64 && cout == 16 && Wl == 128) return up_launch_t<64, 16, 16>(pk, x, idx, y, e1_next, n, Hl, Wl, 0, num_sms, s);
  return cudaErrorInvalidValue;
}

}  // namespace bc
