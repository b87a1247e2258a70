// ENet bottlenecks on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), 16-bit storage
// (fp16 or bf16, see umma_common.cuh), fp32 accumulation.  One kernel per bottleneck:
//
//   e1 (CI ch, produced by the previous kernel) --TMA tap boxes--> smem
//        conv KxK / dilated / 5x1 / 1x5  : NT taps x (CI/16) tcgen05.mma   -> D1 [128 x CI]  (TMEM)
//        epilogue 1: +bias, PReLU, bf16                                   -> e2 tile (smem, A operand)
//        expand 1x1 : (CI/16) tcgen05.mma                                 -> D2 [128 x C]   (TMEM)
//        epilogue 2: +bias, PReLU, + x (residual tile, TMA), PReLU, bf16  -> y tile (smem) -> TMA store
//        next block's projection 1x1 : (C/16) tcgen05.mma on the y tile   -> D3 [128 x CI]  (TMEM)
//        epilogue 3: +bias, PReLU, bf16                                   -> e1' (global)
//
// so a bottleneck moves x in, y out and the two quarter-width tensors e1 / e1' through HBM/L2
// and nothing else.  M = 128 output pixels per tile = one TMEM lane per pixel; a tile is a
// run of whole image rows (2 rows of 64 at 32x64, 1 row of 128 at 64x128), so every conv tap
// is ONE TMA box load whose out-of-bounds zero fill implements padding and dilation.
// Operands are K-major in shared memory in the canonical swizzled layouts (32/64/128-byte
// rows), the same layouts TMA writes.
//
// Down-sampling bottlenecks reuse the kernel ("narrow residual", CRES < C): e1 is then the output
// of the strided 2x2 conv (simt_down.cu), the residual is the max-pooled input with CRES channels
// (the remaining C - CRES output channels get no residual: the zero padding of the main branch),
// y has its own buffer per group, and the next block's projection may be wider than CI (CN).
//
// Reference: the regular / dilated / asymmetric bottlenecks of the frozen ENet graph the
// reference executes (models.py:43-44); semantics in oracle/enet_oracle.py `regular`.
#include "umma_common.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

namespace bc {
namespace BC_NS {

// =================================================================== the kernel
struct UmmaParams {
  uint64_t pol_x, pol_y, pol_e1, pol_e1n;   // L2 eviction-priority hints: x loads, y stores, e1 tap loads, e1' stores
  int num_tiles;        // 128-pixel tiles in this launch
  int tiles_per_frame;  // H*W/128
  int rows_per_tile;    // image rows a tile spans (128 / W)
  int ntaps;
  int8_t dy[9], dx[9];
  int reverse;          // 1: walk the tiles from the last to the first (see launch_one: L2 reuse between consecutive kernels)
  int vslab;            // 1: 5x1 conv on a full-width two-row tile: ONE 6-row box per tile, ky taps = descriptors 64 pixels apart
  int rowslab;          // 1: 3x3 dilation-1 conv on a one-row tile: 3 row loads of 130 pixels, kx through shifted descriptors
  int pairslab;         // 1: 3x3 (any dilation) / 1x5 conv on a two-row tile of a 64-wide map: one pixel-interleaved slab per kernel row (see below)
  int pair_step;        // pair-slab: pixels between two kernel columns (the dilation; 1 for 1x5)
  int pair_halo;        // pair-slab: pixels left of the tile in the slab (dilation, or 2 for 1x5)
  int has_next;         // 1: compute the next block's projection from the y tile
  act_t* out_small;     // e2 (conv-only specialisation) or e1' (has_next): [pixels][CI]
  const uint8_t* wblob; // packed weights, exact shared-memory image (see UmmaSmem)
  // fp32 bias / PReLU-slope block b2[CI] a2[CI] b3[C] a3[C] aout[C] b1n[CN] a1n[CN], by value: the
  // epilogues index it with compile-time channel numbers, so every use is a constant-bank operand
  // of the FADD / FMUL itself (no load instruction, no shared-memory traffic)
  // fused asymmetric block: + b0[CI] a0[CI] of the 5x1 conv at the end
  float f[576];
};

// weight image (identical in global and shared memory): offsets relative to its start
template <int C, int CI, int CN, int NT = 9>
struct UmmaWeights {
  static constexpr int RB = CI * 2;                 // row bytes of the CI-wide operands (64 / 32)
  static constexpr int W2_TAP = CI * RB;            // one tap of W2: [CI out][CI in]
  static constexpr int W3_BYTES = C * RB;           // [C out][CI in]
  static constexpr int NSUB = C / 64;
  static constexpr int W1_SUB = CN * 128;           // [CN out][64 in] sub-tile of the next projection
  static constexpr int OFF_W2 = 0;
  static constexpr int OFF_W3 = NT * W2_TAP;
  static constexpr int OFF_W1 = ((NT * W2_TAP + W3_BYTES + 1023) / 1024) * 1024;   // 128-byte swizzle: 1 KB aligned
  static constexpr int W_BYTES = OFF_W1 + NSUB * W1_SUB;
  static constexpr int NF = 2 * CI + 3 * C + 2 * CN + (NT == 10 ? 2 * CI : 0);   // fp32 parameter block
};

// CN = width of the next block's projection, CRES = residual channels (C: regular bottleneck,
// < C: down-sampling bottleneck), NG = epilogue groups per CTA (tiles in flight), MINB = CTAs per SM
// CONV: conv-only specialisation (first half of an asymmetric bottleneck: taps -> e2 in global memory);
// no residual / y tiles, no e2 tile, only the conv weights: small enough for two CTAs per SM
// ASYM: the whole asymmetric block in one kernel: 5x1 conv (vertical slab from the ring) -> D0 -> epilogue 0 writes the
// result as a PIXEL-INTERLEAVED slab into shared memory -> 1x5 conv through shifted descriptors on that slab -> D1 ->
// the usual chain.  The intermediate tensor never leaves the SM and one launch per block disappears.
template <int C, int CI, int CN, int CRES, int NG_, int MINB_, bool CONV = false, int EPW_ = 1, bool ASYM_ = false>
struct UmmaSmem {
  static constexpr bool ASYM = ASYM_;
  using Wt = UmmaWeights<C, CI, CN, ASYM_ ? 10 : 9>;
  static constexpr int SERVICE = ASYM ? 8 : 4;      // service warps (ASYM: a second conv issuer; padded to a multiple of four so that
                                                    // an epilogue warp's index mod 4 stays its TMEM lane quarter)
  static constexpr int NG = NG_, MINB = MINB_;
  static constexpr bool NARROW = CRES < C;
  static constexpr int EPW = EPW_;                  // warps per TMEM lane quarter of an epilogue group
  static constexpr int THREADS = 32 * SERVICE + 128 * NG * EPW;
  static constexpr int RB = CI * 2;
  static constexpr int TAP_BYTES = 128 * RB;        // one A tap tile
  // ring slot: a tap tile, or (row-slab mode, CI = 16) one image row of 128 + 2 pixels, padded to the 32-byte-swizzle repeat
  static constexpr int ROW_BYTES = 130 * RB;
  static constexpr int SLOT_BYTES = CI == 16 ? 17 * 256 : TAP_BYTES;
  static constexpr int XSUB = 128 * 128;            // one 64-channel sub-tile of x / y
  static constexpr int NSUB = C / 64;
  static constexpr int XBUF = NSUB * XSUB;          // one x / y tile
  // Three groups on the 128-channel block: the only way to a third tile in flight is to give up the spare
  // residual buffer (each group then owns one x / y buffer) and to let D3 (next projection) reuse D1's TMEM
  // columns (D1 is dead once epilogue 1 has read it): 3 x (32 + 128) = 480 of 512 columns
  static constexpr bool ALIAS13 = !CONV && !NARROW && NG >= 3 && C == 128;
  // (a second spare buffer paid for with the tap ring -- four x buffers, one tile of taps -- measured 68.5 vs 65.6 us)
  static constexpr int NX = ALIAS13 ? NG : NG + 1;  // residual tile ring (one tile of prefetch)
  static constexpr int NY = CONV ? 0 : NARROW ? NG : NX;   // C-wide tiles: x/y in place, or one y per group
  static constexpr int RES_RB = CRES * 2 >= 128 ? 128 : CRES * 2;   // row bytes / swizzle of a narrow residual tile
  static constexpr int RBUF = (NARROW && !CONV) ? 128 * CRES * 2 : 0;
  static constexpr int NRING = (MINB == 1 && CI == 16) ? 18 : ALIAS13 ? 7 : 9;   // conv-tap ring slots
  // pair-slab mode (CI = 32, two-row tiles): a slot holds one kernel row of the tile, [64 + 2 halo px][2 rows][CI]
  // with halo <= 16: 12 KB; the same shared memory then holds NRING_PAIR slots = NRING_PAIR / 3 tiles of a 3x3 conv
  static constexpr int PAIR_SLOT = 96 * 2 * RB;
  static constexpr int NRING_PAIR = ALIAS13 ? 5 : NRING * SLOT_BYTES / PAIR_SLOT;
  static constexpr int RING_BYTES = (CI == 32 && !CONV && NRING_PAIR * PAIR_SLOT > NRING * SLOT_BYTES) ? NRING_PAIR * PAIR_SLOT
                                                                                                    : NRING * SLOT_BYTES;
  // offsets (all multiples of 1024)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_R = OFF_X + NY * XBUF;
  static constexpr int OFF_TAPS = OFF_R + NX * RBUF;
  static constexpr int OFF_E2 = ((OFF_TAPS + RING_BYTES + 1023) / 1024) * 1024;   // one e2 tile per group
  // ASYM: the group's slab of the 5x1 result, [64 + 4 px][2 rows][CI] = 8 704 B, shares its memory with the e2 tile (the
  // 1x5 MMAs have consumed the slab before epilogue 1 writes e2)
  static constexpr int PS_BYTES = 68 * 2 * RB;
  static constexpr int E2_STRIDE = ASYM ? ((PS_BYTES + 1023) / 1024) * 1024 : TAP_BYTES;
  static constexpr int OFF_W = OFF_E2 + (CONV ? 0 : NG) * E2_STRIDE;  // weight image starts here
  static constexpr int OFF_W2 = OFF_W + Wt::OFF_W2, OFF_W3 = OFF_W + Wt::OFF_W3, OFF_W1 = OFF_W + Wt::OFF_W1;
  static constexpr int W_LOAD = CONV ? 9 * Wt::W2_TAP : Wt::W_BYTES;  // bytes of the weight image this kernel needs
  static_assert(!ASYM || (!CONV && !NARROW && CI == 32), "fused asymmetric block: full 128-channel form only");
  static constexpr int OFF_BAR = OFF_W + ((W_LOAD + 1023) / 1024) * 1024;
  static constexpr int TOTAL = OFF_BAR + 1024;
  // barriers
  static constexpr int X_FULL = 0, D1_FULL = X_FULL + NX, D1_EMPTY = D1_FULL + NG,
                       E2_FULL = D1_EMPTY + NG, D2_FULL = E2_FULL + NG, Y_FULL = D2_FULL + NG, D3_FULL = Y_FULL + NG,
                       D3_EMPTY = D3_FULL + NG, W_FULL = D3_EMPTY + NG, TAP_FULL = W_FULL + 1, TAP_EMPTY = TAP_FULL + NRING,
                       D0_FULL = TAP_EMPTY + NRING, D0_EMPTY = D0_FULL + NG, PS_FULL = D0_EMPTY + NG,
                       NBARS = PS_FULL + NG;
  static_assert(NBARS * 8 + 8 + 4 * NX <= 1024, "barrier block");
  // 227 KB per CTA, 228 KB per SM with 1 KB reserved per resident CTA; + 1 KB alignment slack
  static_assert(TOTAL + 1024 <= 232448 && (TOTAL + 2048) * MINB <= 233472, "shared memory budget");
  // TMEM columns: every group owns a D1 / D2 / D3 accumulator
  static constexpr uint32_t COL_D1 = 0, COL_D2 = NG * CI, COL_D3 = ALIAS13 ? COL_D1 : NG * (CI + C);
  static constexpr uint32_t COL_D0 = NG * (CI + C + CN);              // ASYM: accumulator of the 5x1 conv
  static constexpr uint32_t COLS_USED = CONV ? NG * CI : ALIAS13 ? NG * (CI + C) : NG * (CI + C + CN) + (ASYM ? NG * CI : 0);
  static_assert(!ALIAS13 || CN == CI, "D3 reuses D1's columns");
  static constexpr uint32_t TMEM_COLS = COLS_USED <= 32 ? 32 : COLS_USED <= 64 ? 64 : COLS_USED <= 128 ? 128
                                        : COLS_USED <= 256 ? 256 : 512;
  static_assert(COLS_USED <= 512 && TMEM_COLS * MINB <= 512, "TMEM budget");
};

// Warp-specialised persistent kernel, 128 + 128 * NG threads.  A CTA keeps NG tiles in flight:
// tile k of the CTA belongs to epilogue group k % NG, which owns a D1/D2/D3 accumulator and an
// e2 tile, so the serial chain of one tile (conv MMA -> epilogue 1 -> expansion MMA -> epilogue
// 2 -> projection MMA -> epilogue 3) overlaps with the chains of the other groups.  Every
// service thread blocks on exactly one barrier sequence (a thread that polls several barriers
// turned out to be the bottleneck: ~200 cycles per probe).
//   warp 0    TMA producer   conv taps through a ring of NRING slots (+ the first residual tiles)
//   warp 1    MMA issuer     conv taps      -> D1[group]
//   warp 2    MMA issuer     expansion      -> D2[group]
//   warp 3    MMA issuer     next projection-> D3[group]
//   warps 4.. epilogue       group g = (warp - 4) / 4; one TMEM lane (= pixel) per thread:
//                            D1 -> e2 (smem), D2 + x -> y (smem, TMA store), D3 -> e1' (global);
//                            its first thread stores y and requests the x tile that reuses the buffer
template <int C, int CI, int CN, int CRES, int NG, int MINB, bool CONV = false, int EPW = 1, bool ASYM = false>
__global__ void __launch_bounds__((ASYM ? 256 : 128) + 128 * NG * EPW, MINB)
k_umma_bottleneck(const __grid_constant__ CUtensorMap map_e1,   // 4D [N][H][W][CI], box = one tile, swizzle RB
                  const __grid_constant__ CUtensorMap map_x,    // 2D [pixels][C], box [128 px][64 ch], swizzle 128
                                                                // (narrow: [pixels][CRES], box [128 px][CRES])
                  const __grid_constant__ CUtensorMap map_y,    // same shape, the output
                  const __grid_constant__ UmmaParams p) {
  using S = UmmaSmem<C, CI, CN, CRES, NG, MINB, CONV, EPW, ASYM>;
  using Wt = typename S::Wt;
  constexpr int RB = S::RB;
  constexpr int NX = S::NX;
  constexpr bool NARROW = S::NARROW;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic shared memory is only guaranteed 16-byte aligned: align by hand (1 KB slack requested)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer: LDS/STS, not generic LD/ST
  const uint32_t sbase = smem_u32(smem);
  constexpr int F_B2 = 0, F_A2 = CI, F_B3 = 2 * CI, F_A3 = F_B3 + C, F_AOUT = F_A3 + C, F_B1N = F_AOUT + C, F_A1N = F_B1N + CN;
  constexpr int F_B0 = F_A1N + CN, F_A0 = F_B0 + CI;        // ASYM: bias / slope of the 5x1 conv
  static_assert(F_A1N + CN + (ASYM ? 2 * CI : 0) <= 576, "parameter block");
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  uint32_t* tmem_slot = (uint32_t*)&bars[S::NBARS];
  uint32_t* xfills = tmem_slot + 2;                    // [NX] refills issued per x buffer (see epilogue 2)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index made provably warp-uniform: role branches become uniform branches

  // ---- one-time setup: barriers, weights + parameters to smem (two bulk copies), TMEM
  if (tid == 0) {
    for (int i = 0; i < S::NBARS; ++i) {
      // D1_EMPTY / E2_FULL / Y_FULL: every thread of the group arrives; the rest: one arrival
      const bool by_group = (i >= S::D1_EMPTY && i < S::D1_EMPTY + NG) || (i >= S::E2_FULL && i < S::E2_FULL + NG) ||
                            (i >= S::Y_FULL && i < S::Y_FULL + NG) || (i >= S::D3_EMPTY && i < S::D3_EMPTY + NG) ||
                            (i >= S::D0_EMPTY && i < S::D0_EMPTY + NG) || (i >= S::PS_FULL && i < S::PS_FULL + NG);
      mbar_init(bar(i), by_group ? 128 * EPW : 1);
    }
    for (int i = 0; i < NX; ++i) xfills[i] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar(S::W_FULL), S::W_LOAD);
    bulk_load(sbase + S::OFF_W, p.wblob, S::W_LOAD, bar(S::W_FULL));
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(S::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if constexpr (ASYM) {
    // zero columns of the 5x1-result slabs (2 pixels left and right of the 64: slab rows 0-3 and 132-135); the right
    // ones are never written again, the left ones are re-zeroed by epilogue 0 (the e2 tile shares their memory)
    for (int i = tid; i < NG * 32; i += S::THREADS) {
      const int g = i >> 5, c = i & 31;
      *reinterpret_cast<uint4*>(smem + S::OFF_E2 + g * S::E2_STRIDE + (c < 16 ? c * 16 : 132 * RB + (c - 16) * 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t IDESC_CONV = instr_desc(128, CI), IDESC_EXP = instr_desc(128, C), IDESC_PROJ = instr_desc(128, CN);
  const int T = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // my tiles
  constexpr bool full = !CONV;
  // k-th tile of this CTA; consecutive kernels walk in opposite directions so that a kernel starts on the
  // tiles its predecessor wrote last (still in L2)
  auto tile_of = [&](int k) {
    const int t = (int)blockIdx.x + k * (int)gridDim.x;
    return p.reverse ? p.num_tiles - 1 - t : t;
  };

  // Service warps run their loops convergently (all 32 lanes) and elect one lane per instruction
  // inside the asm (umma_common.cuh): no divergence, no per-lane operand recomputation.
  // Descriptors advance by plain adds: the address field (bits 0-13, 16-byte units) never carries.
  // Programmatic dependent launch: let the next kernel's CTAs start (barrier init, TMEM allocation, weight copy) as
  // soon as SMs drain, and do not touch activations before the previous kernel has completed.  Every global
  // access of this kernel follows from a TMA load of the producer warp, so that warp alone waits.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ============================================================ TMA producer (conv taps)
    if (full)                                            // first NX residual tiles; the rest are
      for (int k = 0; k < T && k < NX; ++k) {            // requested by the thread that frees a buffer
        const int tile = tile_of(k);
        if constexpr (NARROW) {
          mbar_expect_tx_e(bar(S::X_FULL + k), S::RBUF);
          tma_load_2d_e(sbase + S::OFF_R + k * S::RBUF, &map_x, 0, tile * 128, bar(S::X_FULL + k), p.pol_x);
        } else {
          mbar_expect_tx_e(bar(S::X_FULL + k), S::XBUF);
          for (int s = 0; s < S::NSUB; ++s)
            tma_load_2d_e(sbase + S::OFF_X + k * S::XBUF + s * S::XSUB, &map_x, s * 64, tile * 128, bar(S::X_FULL + k), p.pol_x);
        }
      }
    int slot = 0, round = 0;
    // (frame, tile-in-frame) of the current tile, advanced without divisions
    int n = tile_of(0) / p.tiles_per_frame, ty = tile_of(0) % p.tiles_per_frame;
    const int dn = (int)gridDim.x / p.tiles_per_frame, dty = (int)gridDim.x % p.tiles_per_frame;
    for (int k = 0; k < T; ++k) {
      const int y0 = ty * p.rows_per_tile;
      if (CI == 32 && p.vslab) {
        // vertical taps of a full-width tile: rows y0-2 .. y0+3 arrive as one box (three ring slots), tap ky starts ky rows in
        if (round >= 1) mbar_wait(bar(S::TAP_EMPTY + slot), (round - 1) & 1);
        mbar_expect_tx_e(bar(S::TAP_FULL + slot), 3 * S::TAP_BYTES);
        tma_load_4d_e(sbase + S::OFF_TAPS + slot * S::SLOT_BYTES, &map_e1, 0, 0, y0 - 2, n, bar(S::TAP_FULL + slot), p.pol_e1);
        slot += 3;
        if (slot == S::NRING) { slot = 0; ++round; }
      } else if (CI == 32 && !CONV && p.pairslab) {
        // two-row tiles of a 64-wide map: TMEM lane m is pixel (row m % 2, column m / 2) of the tile, so the A
        // operand of kernel row ky is ONE slab [64 + 2 halo pixels][2 image rows][CI] (map_e1 lists the image row
        // BEFORE the column, so TMA writes the two rows pixel-interleaved) and kernel column kx is the same slab
        // read through a descriptor that starts kx * step pixels (128 bytes each) in: 3 loads of 8-12 KB per tile
        // instead of 9 x 8 KB (1x5: one load instead of five), and the ring holds two tiles
        const int nky = p.ntaps == 9 ? 3 : 1;
        const uint32_t bytes = (uint32_t)(64 + 2 * p.pair_halo) * 2 * RB;
        for (int ky = 0; ky < nky; ++ky) {
          if (round >= 1) mbar_wait(bar(S::TAP_EMPTY + slot), (round - 1) & 1);
          mbar_expect_tx_e(bar(S::TAP_FULL + slot), bytes);
          tma_load_4d_e(sbase + S::OFF_TAPS + slot * S::PAIR_SLOT, &map_e1, 0, y0 + (nky == 3 ? (ky - 1) * p.pair_step : 0),
                        -p.pair_halo, n, bar(S::TAP_FULL + slot), p.pol_e1);
          if (++slot == S::NRING_PAIR) { slot = 0; ++round; }
        }
      } else if (CI == 16 && p.rowslab) {
        // one-row tiles: the three taps of a kernel row are the same 130-pixel row slab read at 0 / 1 / 2 pixels
        // offset (map_e1's box is 130 pixels wide here): 3 loads and a third of the L2 -> SM bytes per tile
        for (int ky = 0; ky < 3; ++ky) {
          if (round >= 1) mbar_wait(bar(S::TAP_EMPTY + slot), (round - 1) & 1);
          mbar_expect_tx_e(bar(S::TAP_FULL + slot), S::ROW_BYTES);
          tma_load_4d_e(sbase + S::OFF_TAPS + slot * S::SLOT_BYTES, &map_e1, 0, -1, y0 + ky - 1, n, bar(S::TAP_FULL + slot), p.pol_e1);
          if (++slot == S::NRING) { slot = 0; ++round; }
        }
      } else
      for (int t = 0; t < p.ntaps; ++t) {
        if (round >= 1) mbar_wait(bar(S::TAP_EMPTY + slot), (round - 1) & 1);
        mbar_expect_tx_e(bar(S::TAP_FULL + slot), S::TAP_BYTES);
        tma_load_4d_e(sbase + S::OFF_TAPS + slot * S::SLOT_BYTES, &map_e1, 0, p.dx[t], y0 + p.dy[t], n, bar(S::TAP_FULL + slot), p.pol_e1);
        if (++slot == S::NRING) { slot = 0; ++round; }
      }
      if (p.reverse) {
        n -= dn; ty -= dty;
        if (ty < 0) { ty += p.tiles_per_frame; --n; }
      } else {
        n += dn; ty += dty;
        if (ty >= p.tiles_per_frame) { ty -= p.tiles_per_frame; ++n; }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer: conv taps -> D1[group]
    const uint64_t dA0 = smem_desc<RB>(sbase + S::OFF_TAPS), dB0 = smem_desc<RB>(sbase + S::OFF_W2);
    // the conv fed from the tap ring accumulates into D1 -- or, in the fused asymmetric block, into D0 (5x1 half)
    constexpr uint32_t COL_C = ASYM ? S::COL_D0 : S::COL_D1;
    constexpr int C_FULL = ASYM ? S::D0_FULL : S::D1_FULL, C_EMPTY = ASYM ? S::D0_EMPTY : S::D1_EMPTY;
    int slot = 0, round = 0;
    mbar_wait(bar(S::W_FULL), 0);
    for (int k = 0; k < T; ++k) {
      const int g = k % NG;
      if (k >= NG) {
        mbar_wait(bar(C_EMPTY + g), ((k / NG) - 1) & 1);
        if (S::ALIAS13 && p.has_next) mbar_wait(bar(S::D3_EMPTY + g), ((k / NG) - 1) & 1);   // D3 of the group's previous tile lives in D1's columns
      }
      if (CI == 32 && p.vslab) {
        mbar_wait(bar(S::TAP_FULL + slot), round & 1);
        tc_fence_after();
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)                    // one image row = 64 pixels x 64 bytes = 8 swizzle atoms
#pragma unroll
          for (int kk = 0; kk < CI / 16; ++kk)
            umma_mma_e(tmem + COL_C + g * CI, dA0 + (uint64_t)(slot * (S::SLOT_BYTES >> 4) + ky * 256 + kk * 2),
                        dB0 + (uint64_t)(ky * (Wt::W2_TAP >> 4) + kk * 2), IDESC_CONV, (ky | kk) != 0);
        umma_commit_e(bar(S::TAP_EMPTY + slot));
        slot += 3;
        if (slot == S::NRING) { slot = 0; ++round; }
      } else if (CI == 32 && !CONV && p.pairslab) {
        const int nky = p.ntaps == 9 ? 3 : 1, nkx = p.ntaps == 9 ? 3 : 5;
        const uint32_t dstep = (uint32_t)p.pair_step * (2 * RB >> 4);      // one pixel of the slab = two 64-byte rows
        for (int ky = 0; ky < nky; ++ky) {
          mbar_wait(bar(S::TAP_FULL + slot), round & 1);
          tc_fence_after();
          for (int kx = 0; kx < nkx; ++kx)                  // (the swizzle follows the absolute address: a whole-row shift is legal)
#pragma unroll
            for (int kk = 0; kk < CI / 16; ++kk)
              umma_mma_e(tmem + S::COL_D1 + g * CI, dA0 + (uint64_t)(slot * (S::PAIR_SLOT >> 4) + kx * dstep + kk * 2),
                         dB0 + (uint64_t)((ky * nkx + kx) * (Wt::W2_TAP >> 4) + kk * 2), IDESC_CONV, (ky | kx | kk) != 0);
          umma_commit_e(bar(S::TAP_EMPTY + slot));
          if (++slot == S::NRING_PAIR) { slot = 0; ++round; }
        }
      } else if (CI == 16 && p.rowslab) {
        for (int ky = 0; ky < 3; ++ky) {
          mbar_wait(bar(S::TAP_FULL + slot), round & 1);
          tc_fence_after();
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)                  // the swizzle follows the absolute address: a 32-byte shift is legal
            umma_mma_e(tmem + S::COL_D1 + g * CI, dA0 + (uint64_t)(slot * (S::SLOT_BYTES >> 4) + kx * 2),
                        dB0 + (uint64_t)((ky * 3 + kx) * (Wt::W2_TAP >> 4)), IDESC_CONV, (ky | kx) != 0);
          umma_commit_e(bar(S::TAP_EMPTY + slot));
          if (++slot == S::NRING) { slot = 0; ++round; }
        }
      } else
      for (int t = 0; t < p.ntaps; ++t) {
        mbar_wait(bar(S::TAP_FULL + slot), round & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < CI / 16; ++kk)
          umma_mma_e(tmem + S::COL_D1 + g * CI, dA0 + (uint64_t)(slot * (S::SLOT_BYTES >> 4) + kk * 2),
                      dB0 + (uint64_t)(t * (Wt::W2_TAP >> 4) + kk * 2), IDESC_CONV, (t | kk) != 0);
        umma_commit_e(bar(S::TAP_EMPTY + slot));          // slot reusable once these MMAs retire
        if (++slot == S::NRING) { slot = 0; ++round; }
      }
      umma_commit_e(bar(C_FULL + g));
    }
  } else if (ASYM && warp == 4) {
    // ============================================================ MMA issuer (fused asymmetric block): 1x5 conv on the
    // group's slab of the 5x1 result -> D1[group]; kernel column kx = the slab read kx pixels (128 bytes) in
    const uint64_t dA0 = smem_desc<RB>(sbase + S::OFF_E2), dB0 = smem_desc<RB>(sbase + S::OFF_W2 + 5 * Wt::W2_TAP);
    mbar_wait(bar(S::W_FULL), 0);
    for (int k = 0; k < T; ++k) {
      const int g = k % NG;
      mbar_wait(bar(S::PS_FULL + g), (k / NG) & 1);
      if (k >= NG) mbar_wait(bar(S::D1_EMPTY + g), ((k / NG) - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int kx = 0; kx < 5; ++kx)
#pragma unroll
        for (int kk = 0; kk < CI / 16; ++kk)
          umma_mma_e(tmem + S::COL_D1 + g * CI, dA0 + (uint64_t)(g * (S::E2_STRIDE >> 4) + kx * (2 * RB >> 4) + kk * 2),
                     dB0 + (uint64_t)(kx * (Wt::W2_TAP >> 4) + kk * 2), IDESC_CONV, (kx | kk) != 0);
      umma_commit_e(bar(S::D1_FULL + g));
    }
  } else if (warp == 2) {
    // ============================================================ MMA issuer: expansion e2 -> D2[group]
    if (full) {
      const uint64_t dA0 = smem_desc<RB>(sbase + S::OFF_E2), dB0 = smem_desc<RB>(sbase + S::OFF_W3);
      mbar_wait(bar(S::W_FULL), 0);
      for (int k = 0; k < T; ++k) {
        const int g = k % NG;
        mbar_wait(bar(S::E2_FULL + g), (k / NG) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < CI / 16; ++kk)
          umma_mma_e(tmem + S::COL_D2 + g * C, dA0 + (uint64_t)(g * (S::E2_STRIDE >> 4) + kk * 2), dB0 + (uint64_t)(kk * 2),
                      IDESC_EXP, kk != 0);
        umma_commit_e(bar(S::D2_FULL + g));
      }
    }
  } else if (warp == 3) {
    // ============================================================ MMA issuer: next projection y -> D3[group]
    if (full && p.has_next) {
      const uint64_t dA0 = smem_desc<128>(sbase + S::OFF_X), dB0 = smem_desc<128>(sbase + S::OFF_W1);
      mbar_wait(bar(S::W_FULL), 0);
      for (int k = 0; k < T; ++k) {
        const int g = k % NG;
        const uint32_t xo = (uint32_t)(NARROW ? g : k % NX) * (S::XBUF >> 4);
        mbar_wait(bar(S::Y_FULL + g), (k / NG) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < C / 16; ++kk)
          umma_mma_e(tmem + S::COL_D3 + g * CN, dA0 + (uint64_t)(xo + (kk / 4) * (S::XSUB >> 4) + (kk % 4) * 2),
                      dB0 + (uint64_t)((kk / 4) * (Wt::W1_SUB >> 4) + (kk % 4) * 2), IDESC_PROJ, kk != 0);
        umma_commit_e(bar(S::D3_FULL + g));
      }
    }
  } else if (warp >= S::SERVICE) {
    // ============================================================ epilogue (warps 4.., or 8.. in the fused asymmetric block)
    // EPW warps share a TMEM lane quarter of a group: warp-half eh takes 1 / EPW of the columns of every
    // epilogue (compile-time halves, so the biases and slopes stay constant-bank operands)
    const int grp = (warp - S::SERVICE) / (4 * EPW);  // epilogue group = tile slot
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;             // row of the tile = TMEM lane
    // the tile's pixel (in memory order) this lane holds: itself, or in pair-slab mode (row m % 2, column m / 2)
    const int mp = (CI == 32 && !CONV && p.pairslab) ? ((m & 1) << 6) | (m >> 1) : m;
    const uint32_t tm_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* e2buf = smem + S::OFF_E2 + grp * S::E2_STRIDE;
    mbar_wait(bar(S::W_FULL), 0);
    auto run = [&](auto HH) {
    constexpr int EH = decltype(HH)::value;
    constexpr int CI1 = CI / EPW, C1 = C / EPW, CN1 = CN / EPW;     // this warp's share of the columns
    static_assert(CI1 % 16 == 0 && CN1 % 16 == 0 && C1 % 32 == 0, "column split");
    const bool storer = (q == 0 && EH == 0 && lane == 0);
    // ---- epilogue 0 (fused asymmetric block): 5x1 result + bias, PReLU, 16-bit -> the group's slab.  D0's lanes
    // are in image order (lane m = row m / 64, column m % 64: the vertical-slab conv); the slab is pixel-interleaved
    // [column + 2][row], which is what gives D1 the pair order the later epilogues assume
    auto ep0 = [&](int k) {
      if constexpr (ASYM) {
        static_assert(!ASYM || EPW == 1, "fused asymmetric block: one warp per lane quarter");
        float v[CI];
        mbar_wait(bar(S::D0_FULL + grp), (uint32_t)(k / NG) & 1);
        tc_fence_after();
        tmem_ld32(tm_lane + S::COL_D0 + grp * CI, v);
        tc_fence_before();
        mbar_arrive(bar(S::D0_EMPTY + grp));
#pragma unroll
        for (int j = 0; j < CI; ++j) v[j] = prelu_f(v[j] + p.f[F_B0 + j], p.f[F_A0 + j]);
        const int pos = (((m & 63) + 2) << 1) | (m >> 6);          // slab row of this pixel
#pragma unroll
        for (int c = 0; c < CI / 8; ++c)
          *reinterpret_cast<uint4*>(e2buf + swz<RB>(pos * RB + c * 16)) = pack8(v + 8 * c);
        if (m < 16) *reinterpret_cast<uint4*>(e2buf + m * 16) = make_uint4(0u, 0u, 0u, 0u);   // left zero columns (the previous e2 tile lay over them)
        fence_proxy_async();
        mbar_arrive(bar(S::PS_FULL + grp));
      }
    };
    // ---- epilogue 1: +bias, PReLU, 16-bit -> e2 tile (A operand of the expansion) or global
    auto ep1 = [&](int k) {
      float v[CI1];
      mbar_wait(bar(S::D1_FULL + grp), (uint32_t)(k / NG) & 1);
      tc_fence_after();
      if constexpr (CI1 == 32) tmem_ld32(tm_lane + S::COL_D1 + grp * CI + EH * CI1, v);
      else tmem_ld16(tm_lane + S::COL_D1 + grp * CI + EH * CI1, v);
      tc_fence_before();
      mbar_arrive(bar(S::D1_EMPTY + grp));
#pragma unroll
      for (int j = 0; j < CI1; ++j) v[j] = prelu_f(v[j] + p.f[F_B2 + EH * CI1 + j], p.f[F_A2 + EH * CI1 + j]);
      if constexpr (!full) {
        uint4* o = reinterpret_cast<uint4*>(p.out_small + ((size_t)tile_of(k) * 128 + m) * CI + EH * CI1);
#pragma unroll
        for (int c = 0; c < CI1 / 8; ++c)
          st_global_hint(o + c, pack8(v + 8 * c), p.pol_e1n);
      } else {
#pragma unroll
        for (int c = 0; c < CI1 / 8; ++c)
          *reinterpret_cast<uint4*>(e2buf + swz<RB>(m * RB + (EH * CI1 / 8 + c) * 16)) = pack8(v + 8 * c);
        fence_proxy_async();
        mbar_arrive(bar(S::E2_FULL + grp));
      }
    };
    // The chain of a tile is  [ep0 | 1x5 MMAs] ep1 | expansion MMAs | ep2 | projection MMAs | ep3, and the group idles
    // through every MMA round trip (arrive -> issuer wakes -> MMAs -> commit -> waiters wake, ~600 cycles each).  In the
    // fused asymmetric block the extra stage is hidden by software pipelining: ep0(k + NG) runs right after ep2(k),
    // while tile k's projection MMAs run, and the next tile's 1x5 MMAs run under ep3(k) (82 -> 77 us per launch; the
    // slab is free: the MMAs that read it finished before the D1 / D2 barriers this group has passed).  Pulling ep1
    // forward the same way in the ordinary blocks was measured SLOWER (65 -> 69 us, also when done only if the
    // accumulator is already there: the later ep3 delays the residual buffer's refill), so they keep program order.
    if constexpr (ASYM) {
      if (grp < T) ep0(grp);
    }
    for (int k = grp; k < T; k += NG) {
      const int tile = tile_of(k);
      const uint32_t par = (uint32_t)(k / NG) & 1;
      ep1(k);
      if constexpr (!full) continue;
      // ---- epilogue 2: +bias, PReLU, + residual, PReLU, 16-bit -> y tile (in place over x; narrow: own buffer)
      const int xb = k % NX;
      uint8_t* yt = smem + S::OFF_X + (NARROW ? grp : xb) * S::XBUF;
      const uint8_t* rt = smem + S::OFF_R + xb * S::RBUF;         // narrow residual tile
      // The x ring is shared by the groups, and a parity wait cannot tell "fill j done" from "fill j-1 still
      // pending": a group that runs two tiles ahead of the other one (nothing orders them when there is no
      // projection stage) would sail through, read the previous tile's x and then re-arm a barrier whose
      // phase is still open (a trap).  So the storer publishes how many refills it has issued per buffer and
      // the consumer checks that its fill exists before trusting the parity (almost always true at once).
      if (k >= NX) {
        const uint32_t need = (uint32_t)(k / NX);
        uint32_t have;
        do {
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(have) : "r"(smem_u32(&xfills[xb])) : "memory");
        } while (have < need);
      }
      mbar_wait(bar(S::X_FULL + xb), (uint32_t)(k / NX) & 1);
      mbar_wait(bar(S::D2_FULL + grp), par);
      tc_fence_after();
#pragma unroll
      for (int c0 = EH * C1; c0 < (EH + 1) * C1; c0 += 32) {
        float v[32];
        tmem_ld32(tm_lane + S::COL_D2 + grp * C + c0, v);
        uint8_t* yrow = yt + (c0 / 64) * S::XSUB;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                 // 4 chunks of 8 channels
          const int ch = c0 + 8 * c;
          uint4* py = reinterpret_cast<uint4*>(yrow + swz<128>(mp * 128 + ((ch % 64) / 8) * 16));
          uint4 xr = make_uint4(0u, 0u, 0u, 0u);      // channels beyond CRES: zero padding of the main branch
          if constexpr (!NARROW) xr = *py;
          else if (ch < CRES) xr = *reinterpret_cast<const uint4*>(rt + swz<S::RES_RB>(mp * S::RES_RB + (ch / 8) * 16));
          const uint32_t* xh = reinterpret_cast<const uint32_t*>(&xr);
          float o[8];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            float2 xf = unpack_act(xh[qq]);
            int j = 8 * c + 2 * qq;
            o[2 * qq] = prelu_f(prelu_f(v[j] + p.f[F_B3 + c0 + j], p.f[F_A3 + c0 + j]) + xf.x, p.f[F_AOUT + c0 + j]);
            o[2 * qq + 1] = prelu_f(prelu_f(v[j + 1] + p.f[F_B3 + c0 + j + 1], p.f[F_A3 + c0 + j + 1]) + xf.y,
                                    p.f[F_AOUT + c0 + j + 1]);
          }
          *py = pack8(o);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(S::Y_FULL + grp));
      // ---- y tile -> global (TMA store by one thread once every row is written)
      if (storer) {
        mbar_wait(bar(S::Y_FULL + grp), par);
        for (int s = 0; s < S::NSUB; ++s)
          tma_store_2d(&map_y, smem_u32(yt) + s * S::XSUB, s * 64, tile * 128, p.pol_y);
        tma_store_commit();
        if constexpr (NARROW) {                       // every thread has read the residual tile: refill it
          if (k + NX < T) {
            const int nt = tile_of(k + NX);
            mbar_expect_tx(bar(S::X_FULL + xb), S::RBUF);
            tma_load_2d(sbase + S::OFF_R + xb * S::RBUF, &map_x, 0, nt * 128, bar(S::X_FULL + xb), p.pol_x);
            asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&xfills[xb])), "r"((uint32_t)((k + NX) / NX)) : "memory");
          }
        }
      }
      // ---- pulled forward: the epilogue that feeds the group's next tile (see above)
      if constexpr (ASYM) {
        if (k + NG < T) ep0(k + NG);
      }
      // ---- epilogue 3: the next block's projection: +bias, PReLU, 16-bit -> e1' (global)
      if (p.has_next) {
        mbar_wait(bar(S::D3_FULL + grp), par);
        tc_fence_after();
        float v[CN1];
        if constexpr (CN1 == 32) tmem_ld32(tm_lane + S::COL_D3 + grp * CN + EH * CN1, v);
        else tmem_ld16(tm_lane + S::COL_D3 + grp * CN + EH * CN1, v);
        tc_fence_before();
        if constexpr (S::ALIAS13) mbar_arrive(bar(S::D3_EMPTY + grp));
#pragma unroll
        for (int j = 0; j < CN1; ++j) v[j] = prelu_f(v[j] + p.f[F_B1N + EH * CN1 + j], p.f[F_A1N + EH * CN1 + j]);
        uint4* o = reinterpret_cast<uint4*>(p.out_small + ((size_t)tile * 128 + m) * CN + EH * CN1);
#pragma unroll
        for (int c = 0; c < CN1 / 8; ++c)
          st_global_hint(o + c, pack8(v + 8 * c), p.pol_e1n);
      }
      // The y buffer may be overwritten once the store has read it (and the projection MMAs, which
      // the D3_FULL wait above covers, have consumed it).  In place: the thread that knows first
      // requests the x tile that reuses the buffer.  Narrow: the group owns the buffer, so its
      // threads meet on a named barrier before the next tile's epilogue 2 writes it.
      if (storer) {
        tma_store_wait_read();
        if constexpr (!NARROW) {
          if (k + NX < T) {
            const int nt = tile_of(k + NX);
            mbar_expect_tx(bar(S::X_FULL + xb), S::XBUF);
            for (int s = 0; s < S::NSUB; ++s)
              tma_load_2d(sbase + S::OFF_X + xb * S::XBUF + s * S::XSUB, &map_x, s * 64, nt * 128, bar(S::X_FULL + xb), p.pol_x);
            asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&xfills[xb])), "r"((uint32_t)((k + NX) / NX)) : "memory");
          }
        }
      }
      if constexpr (NARROW) asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(128 * EPW) : "memory");
    }
    if (storer) tma_store_wait_all();
    };
    if constexpr (EPW == 1) {
      run(std::integral_constant<int, 0>{});
    } else {
      if (((warp - S::SERVICE) >> 2) % EPW == 0) run(std::integral_constant<int, 0>{}); else run(std::integral_constant<int, 1>{});
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(S::TMEM_COLS));
  }
}

// =================================================================== host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// 4D map over a [N][H][W][CI] bf16 tensor; box = [1][rows][cols][CI] (128 pixels)
bool make_map_e1(CUtensorMap* m, const act_t* base, int N, int H, int W, int CI) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return false;
  int rows = 128 / W > 0 ? 128 / W : 1;
  cuuint64_t dims[4] = {(cuuint64_t)CI, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)CI * 2, (cuuint64_t)W * CI * 2, (cuuint64_t)H * W * CI * 2};
  cuuint32_t box[4] = {(cuuint32_t)CI, (cuuint32_t)(128 / rows), (cuuint32_t)rows, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, ACT_TMAP, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle_for(CI * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// same tensor view, explicit box [1][box_rows][box_w][CI]
bool make_map_box(CUtensorMap* m, const act_t* base, int N, int H, int W, int CI, int box_w, int box_rows) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)CI, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)CI * 2, (cuuint64_t)W * CI * 2, (cuuint64_t)H * W * CI * 2};
  cuuint32_t box[4] = {(cuuint32_t)CI, (cuuint32_t)box_w, (cuuint32_t)box_rows, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, ACT_TMAP, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle_for(CI * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// pair-slab view of a [N][H][W][CI] tensor: dims (CI, y, x, n) -- the image row BEFORE the column -- and box
// [1][box_x][2][CI], so that TMA writes two image rows pixel-interleaved: smem [x][row][CI]
bool make_map_pair(CUtensorMap* m, const act_t* base, int N, int H, int W, int CI, int box_x) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)CI, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)W * CI * 2, (cuuint64_t)CI * 2, (cuuint64_t)H * W * CI * 2};
  cuuint32_t box[4] = {(cuuint32_t)CI, 2, (cuuint32_t)box_x, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, ACT_TMAP, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle_for(CI * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// 2D map over [pixels][C] bf16; box = [128 px][64 ch], 128-byte swizzle
bool make_map_x(CUtensorMap* m, const act_t* base, size_t pixels, int C) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)pixels};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  return enc(m, ACT_TMAP, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2D map over [pixels][C] bf16; box = [box_px][C] (whole pixel rows), swizzle = C * 2 bytes
bool make_map_rows(CUtensorMap* m, const act_t* base, size_t pixels, int C, int box_px, int sw) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)pixels};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)box_px};
  cuuint32_t es[2] = {1, 1};
  return enc(m, ACT_TMAP, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle_for(sw), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 5D window view of a [N][2Ho][2Wo][C] bf16 tensor: dims (C, dx, ox, dy, n*Ho + oy); one box =
// one window position of `rows` output rows x `box_w` output pixels
bool make_map_window(CUtensorMap* m, const act_t* base, int N, int Ho, int Wo, int C, int rows, int box_w) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)Wo, 2, (cuuint64_t)N * Ho};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)2 * C * 2, (cuuint64_t)2 * Wo * C * 2, (cuuint64_t)4 * Wo * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)C, 1, (cuuint32_t)box_w, 1, (cuuint32_t)rows};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(m, ACT_TMAP, 5, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle_for(C * 2 >= 128 ? 128 : C * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// K-major operand image: rows of `row_bytes` (one swizzle row each), 16-bit activation type, swizzled
static void pack_rows(uint8_t* dst, int rows, int row_bytes, const std::vector<float>& w, int cin_total, int cout_total,
                      int tap, int k0) {
  // element (row = out channel o, k = in channel k0 + kk) of folded [tap][cin][cout]
  int kper = row_bytes / 2;
  for (int o = 0; o < rows; ++o)
    for (int kk = 0; kk < kper; ++kk) {
      const uint16_t h = host_act_bits(w[((size_t)tap * cin_total + (k0 + kk)) * cout_total + o]);
      uint32_t off = (uint32_t)(o * row_bytes + kk * 2);
      uint32_t so = row_bytes == 128 ? swz<128>(off) : row_bytes == 64 ? swz<64>(off) : swz<32>(off);
      memcpy(dst + so, &h, 2);
    }
}

// Builds the device weight image + fp32 parameter block of one bottleneck.
//   conv: folded [ntaps][CI][CI] (+bias/alpha); expand: [1][CI][C]; next: [1][C][CN] of the NEXT block (may be null)
template <int C, int CI, int CN>
static bool build_t(UmmaPack& out, const float* conv_w, int ntaps, const float* conv_b, const float* conv_a,
                    const float* exp_w, const float* exp_b, const float* exp_a, const float* alpha_out,
                    const float* next_w, const float* next_b, const float* next_a) {
  using S = UmmaWeights<C, CI, CN>;
  std::vector<uint8_t> img(S::W_BYTES, 0);
  std::vector<float> cw(conv_w, conv_w + (size_t)ntaps * CI * CI);
  for (int t = 0; t < ntaps; ++t) pack_rows(img.data() + S::OFF_W2 + t * S::W2_TAP, CI, S::RB, cw, CI, CI, t, 0);
  if (exp_w) {
    std::vector<float> ew(exp_w, exp_w + (size_t)CI * C);
    pack_rows(img.data() + S::OFF_W3, C, S::RB, ew, CI, C, 0, 0);
  }
  if (next_w) {
    std::vector<float> nw(next_w, next_w + (size_t)C * CN);
    for (int s = 0; s < S::NSUB; ++s)
      pack_rows(img.data() + S::OFF_W1 + s * S::W1_SUB, CN, 128, nw, C, CN, 0, s * 64);
  }
  std::vector<float> f(S::NF, 0.f);
  float* b2 = f.data(); float* a2 = b2 + CI; float* b3 = a2 + CI; float* a3 = b3 + C; float* ao = a3 + C;
  float* b1n = ao + C; float* a1n = b1n + CN;
  for (int j = 0; j < CI; ++j) { b2[j] = conv_b[j]; a2[j] = conv_a[j]; }
  if (exp_w) for (int j = 0; j < C; ++j) { b3[j] = exp_b[j]; a3[j] = exp_a[j]; ao[j] = alpha_out[j]; }
  if (next_w) for (int j = 0; j < CN; ++j) { b1n[j] = next_b[j]; a1n[j] = next_a[j]; }
  if (cudaMalloc(&out.wblob, S::W_BYTES) != cudaSuccess) return false;
  cudaMemcpy(out.wblob, img.data(), S::W_BYTES, cudaMemcpyHostToDevice);
  out.hf = f;
  out.C = C; out.CI = CI; out.CN = CN; out.ntaps = ntaps; out.has_exp = exp_w != nullptr; out.has_next = next_w != nullptr;
  return true;
}

template <int C, int CI, int CN, int CRES, int NG, int MINB, bool CONV = false, int EPW = 1, bool ASYM = false>
static cudaError_t launch_one(const UmmaPack& pk, const act_t* e1, const act_t* x, act_t* y, act_t* out_small, int n, int H,
                              int W, const Taps& taps, int ntaps, int conv_only, int has_next, int num_sms, cudaStream_t s) {
  using S = UmmaSmem<C, CI, CN, CRES, NG, MINB, CONV, EPW, ASYM>;
  CUtensorMap me1, mx, my;
  // row-slab mode: a plain 3x3 conv (dilation 1, taps in row-major order) whose 128-pixel tile is one image row
  static const bool no_rowslab = getenv("BC_NO_ROWSLAB") != nullptr, no_vslab = getenv("BC_NO_VSLAB") != nullptr,
                    no_pairslab = getenv("BC_NO_PAIRSLAB") != nullptr;   // read once
  // pair-slab mode: a 3x3 conv (one dilation for both axes, taps in row-major order) or a 1x5 conv on two-row tiles
  int pair_step = 0, pair_halo = 0;
  bool pairslab = CI == 32 && !CONV && W == 64 && H % 2 == 0 && !no_pairslab && (ntaps == 9 || ntaps == 5);
  if (pairslab && ntaps == 9) {
    pair_step = pair_halo = taps.dx[8];
    pairslab = pair_step >= 1 && pair_step <= 16;
    for (int t = 0; pairslab && t < 9; ++t) pairslab = taps.dy[t] == (t / 3 - 1) * pair_step && taps.dx[t] == (t % 3 - 1) * pair_step;
  } else if (pairslab) {
    pair_step = 1; pair_halo = 2;
    for (int t = 0; pairslab && t < 5; ++t) pairslab = taps.dy[t] == 0 && taps.dx[t] == t - 2;
  }
  bool rowslab = CI == 16 && !CONV && ntaps == 9 && W == 128 && !no_rowslab;
  for (int t = 0; rowslab && t < 9; ++t) rowslab = taps.dy[t] == t / 3 - 1 && taps.dx[t] == t % 3 - 1;
  // vertical-slab mode: a 5x1 conv (taps dy = -2..2, dx = 0) whose tile is two full-width rows of 64 pixels
  bool vslab = CI == 32 && S::NRING == 9 && ntaps == 5 && W == 64 && !no_vslab;
  for (int t = 0; vslab && t < 5; ++t) vslab = taps.dy[t] == t - 2 && taps.dx[t] == 0;
  if (ASYM) {        // fused asymmetric block: `taps` are the 5x1 taps (vertical slab from the ring); the 1x5 half runs on-chip
    if (!vslab || W != 64 || H % 2 != 0) return cudaErrorInvalidValue;
    pairslab = true; pair_step = 1; pair_halo = 2;      // lane order of D1 and of the later epilogues
  }
  if (ASYM ? !make_map_box(&me1, e1, n, H, W, CI, 64, 6)
      : pairslab ? !make_map_pair(&me1, e1, n, H, W, CI, 64 + 2 * pair_halo)
      : rowslab ? !make_map_box(&me1, e1, n, H, W, CI, 130, 1)
      : vslab ? !make_map_box(&me1, e1, n, H, W, CI, 64, 6) : !make_map_e1(&me1, e1, n, H, W, CI))
    return cudaErrorInvalidValue;
  size_t px = (size_t)n * H * W;
  if (S::NARROW) {
    if (!make_map_rows(&mx, x, px, CRES, 128, S::RES_RB)) return cudaErrorInvalidValue;
    if (!make_map_x(&my, y, px, C)) return cudaErrorInvalidValue;
  } else {
    // conv_only never touches x / y: reuse the e1 tensor as a valid placeholder address
    if (!make_map_x(&mx, conv_only ? e1 : x, conv_only ? px * CI / C : px, C)) return cudaErrorInvalidValue;
    if (!make_map_x(&my, conv_only ? e1 : y, conv_only ? px * CI / C : px, C)) return cudaErrorInvalidValue;
  }
  UmmaParams p{};
  p.num_tiles = (int)(px / 128);
  p.tiles_per_frame = H * W / 128;
  p.rows_per_tile = 128 / W > 0 ? 128 / W : 1;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) { p.dy[t] = taps.dy[t]; p.dx[t] = taps.dx[t]; }
  if ((conv_only != 0) != CONV) return cudaErrorInvalidValue;
  p.rowslab = rowslab ? 1 : 0;
  p.vslab = vslab ? 1 : 0;
  p.pairslab = pairslab ? 1 : 0;
  p.pair_step = pair_step;
  p.pair_halo = pair_halo;
  p.reverse = g_umma_reverse;
  {
    // BC_L2_HINTS = four letters (n normal, f evict-first, l evict-last) for: x loads, y stores, e1 tap loads, e1' stores
    static const uint64_t* pol = [] {
      static uint64_t v[4] = {L2_EVICT_NORMAL, L2_EVICT_NORMAL, L2_EVICT_NORMAL, L2_EVICT_NORMAL};
      const char* e = getenv("BC_L2_HINTS");
      for (int i = 0; e && i < 4 && e[i]; ++i) v[i] = e[i] == 'f' ? L2_EVICT_FIRST : e[i] == 'l' ? L2_EVICT_LAST : L2_EVICT_NORMAL;
      return v;
    }();
    p.pol_x = pol[0]; p.pol_y = pol[1]; p.pol_e1 = pol[2]; p.pol_e1n = pol[3];
  }
  p.has_next = has_next;
  p.out_small = out_small;
  p.wblob = pk.wblob;
  memcpy(p.f, pk.hf.data(), pk.hf.size() * sizeof(float));
  const int smem = S::TOTAL + 1024;        // opt-in set per device by prepare_bottleneck()
  const int ctas = num_sms * MINB;
  int grid = p.num_tiles < ctas ? p.num_tiles : ctas;
  static const bool pdl = getenv("BC_PDL") != nullptr && atoi(getenv("BC_PDL")) != 0;   // A/B knob, read once
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(S::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_umma_bottleneck<C, CI, CN, CRES, NG, MINB, CONV, EPW, ASYM>, me1, mx, my, p);
  }
  k_umma_bottleneck<C, CI, CN, CRES, NG, MINB, CONV, EPW, ASYM><<<grid, S::THREADS, smem, s>>>(me1, mx, my, p);
  return cudaGetLastError();
}

// Tiles in flight per SM = epilogue groups per CTA x CTAs per SM.  Tuning knob (measured on
// B200): BC_UMMA_CFG64 = "<groups><ctas>" for the 64-channel regular bottleneck.
static int cfg_from_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && v[0] >= '1' && v[0] <= '4' && v[1] >= '1' && v[1] <= '2' && !v[2]) ? (v[0] - '0') * 10 + (v[1] - '0') : dflt;
}

}  // namespace BC_NS
using namespace BC_NS;

bool Umma<act_t>::available() { return encode_fn() != nullptr; }

// the kernel configurations launch() can select (one opt-in per instantiation and device)
#define BC_UMMA_CONFIGS(X) \
  X(128, 32, 32, 128, 2, 2, true, 1) X(128, 32, 32, 128, 2, 1, false, 1) X(128, 32, 32, 128, 3, 1, false, 1) \
  X(128, 32, 32, 128, 2, 1, false, 2) X(64, 16, 16, 64, 1, 2, false, 1) X(64, 16, 16, 64, 2, 2, false, 1) \
  X(64, 16, 16, 64, 4, 1, false, 1) X(64, 16, 16, 16, 2, 2, false, 1) X(128, 16, 32, 64, 2, 1, false, 1)

cudaError_t Umma<act_t>::prepare_bottleneck() {
  cudaError_t e = cudaSuccess;
#define BC_SET(C_, CI_, CN_, CR_, NG_, MB_, CV_, EP_)                                                         \
  if (e == cudaSuccess)                                                                                        \
    e = cudaFuncSetAttribute(k_umma_bottleneck<C_, CI_, CN_, CR_, NG_, MB_, CV_, EP_>,                         \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaSmem<C_, CI_, CN_, CR_, NG_, MB_, CV_, EP_>::TOTAL + 1024);
  BC_UMMA_CONFIGS(BC_SET)
#undef BC_SET
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_umma_bottleneck<128, 32, 32, 128, 2, 1, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             UmmaSmem<128, 32, 32, 128, 2, 1, false, 1, true>::TOTAL + 1024);
  return e;
}

// (C, CI, CN, CRES) combinations of the ENet graph: regular stage-2/3 and stage-1/4 bottlenecks,
// downsample1_0 (internal width 4 zero-padded to 16) and downsample2_0
bool Umma<act_t>::build(UmmaPack& out, int C, int CI, int CN, int CRES, const float* conv_w, int ntaps, const float* conv_b,
                const float* conv_a, const float* exp_w, const float* exp_b, const float* exp_a, const float* alpha_out,
                const float* next_w, const float* next_b, const float* next_a) {
  bool ok = false;
  if (C == 128 && CI == 32 && CN == 32 && CRES == 128)
    ok = build_t<128, 32, 32>(out, conv_w, ntaps, conv_b, conv_a, exp_w, exp_b, exp_a, alpha_out, next_w, next_b, next_a);
  else if (C == 64 && CI == 16 && CN == 16 && (CRES == 64 || CRES == 16))
    ok = build_t<64, 16, 16>(out, conv_w, ntaps, conv_b, conv_a, exp_w, exp_b, exp_a, alpha_out, next_w, next_b, next_a);
  else if (C == 128 && CI == 16 && CN == 32 && CRES == 64)
    ok = build_t<128, 16, 32>(out, conv_w, ntaps, conv_b, conv_a, exp_w, exp_b, exp_a, alpha_out, next_w, next_b, next_a);
  out.CRES = CRES;
  return ok;
}

// Fused asymmetric block (128 channels): W2 area = 5 taps of the 5x1 conv then 5 taps of the 1x5 conv; the fp32 block
// ends with the 5x1 conv's bias / slope.
bool Umma<act_t>::build_asym(UmmaPack& out, const float* c0_w, const float* c0_b, const float* c0_a, const float* c1_w,
                             const float* c1_b, const float* c1_a, const float* exp_w, const float* exp_b, const float* exp_a,
                             const float* alpha_out, const float* next_w, const float* next_b, const float* next_a) {
  constexpr int C = 128, CI = 32, CN = 32;
  using Wt = UmmaWeights<C, CI, CN, 10>;
  std::vector<uint8_t> img(Wt::W_BYTES, 0);
  std::vector<float> w0(c0_w, c0_w + (size_t)5 * CI * CI), w1(c1_w, c1_w + (size_t)5 * CI * CI);
  for (int t = 0; t < 5; ++t) {
    pack_rows(img.data() + Wt::OFF_W2 + t * Wt::W2_TAP, CI, Wt::RB, w0, CI, CI, t, 0);
    pack_rows(img.data() + Wt::OFF_W2 + (5 + t) * Wt::W2_TAP, CI, Wt::RB, w1, CI, CI, t, 0);
  }
  std::vector<float> ew(exp_w, exp_w + (size_t)CI * C);
  pack_rows(img.data() + Wt::OFF_W3, C, Wt::RB, ew, CI, C, 0, 0);
  if (next_w) {
    std::vector<float> nw(next_w, next_w + (size_t)C * CN);
    for (int sidx = 0; sidx < Wt::NSUB; ++sidx) pack_rows(img.data() + Wt::OFF_W1 + sidx * Wt::W1_SUB, CN, 128, nw, C, CN, 0, sidx * 64);
  }
  std::vector<float> f(Wt::NF, 0.f);
  float* b2 = f.data(); float* a2 = b2 + CI; float* b3 = a2 + CI; float* a3 = b3 + C; float* ao = a3 + C;
  float* b1n = ao + C; float* a1n = b1n + CN; float* b0 = a1n + CN; float* a0 = b0 + CI;
  for (int j = 0; j < CI; ++j) { b2[j] = c1_b[j]; a2[j] = c1_a[j]; b0[j] = c0_b[j]; a0[j] = c0_a[j]; }
  for (int j = 0; j < C; ++j) { b3[j] = exp_b[j]; a3[j] = exp_a[j]; ao[j] = alpha_out[j]; }
  if (next_w) for (int j = 0; j < CN; ++j) { b1n[j] = next_b[j]; a1n[j] = next_a[j]; }
  if (cudaMalloc(&out.wblob, Wt::W_BYTES) != cudaSuccess) return false;
  cudaMemcpy(out.wblob, img.data(), Wt::W_BYTES, cudaMemcpyHostToDevice);
  out.hf = f;
  out.C = C; out.CI = CI; out.CN = CN; out.CRES = C; out.ntaps = 5; out.has_exp = true; out.has_next = next_w != nullptr;
  return true;
}

// e1: the block's projection; taps5x1: (dy = -2..2, dx = 0)
cudaError_t Umma<act_t>::launch_asym(const UmmaPack& pk, const act_t* e1, const act_t* x, act_t* y, act_t* out_small, int n, int H,
                                     int W, const Taps& taps5x1, int has_next, int num_sms, cudaStream_t s) {
  if (pk.C != 128 || pk.CI != 32 || W != 64) return cudaErrorInvalidValue;
  return launch_one<128, 32, 32, 128, 2, 1, false, 1, true>(pk, e1, x, y, out_small, n, H, W, taps5x1, 5, 0, has_next, num_sms, s);
}

cudaError_t Umma<act_t>::launch(const UmmaPack& pk, const act_t* e1, const act_t* x, act_t* y, act_t* out_small, int n, int H,
                                int W, const Taps& taps, int conv_only, int has_next, int num_sms, cudaStream_t s) {
  if (128 % W != 0 && W % 128 != 0) return cudaErrorInvalidValue;
  static const int cfg64 = cfg_from_env("BC_UMMA_CFG64", 41), cfg128 = cfg_from_env("BC_UMMA_CFG128", 21);
#define BC_LAUNCH(C_, CI_, CN_, CR_, NG_, MB_) \
  return launch_one<C_, CI_, CN_, CR_, NG_, MB_>(pk, e1, x, y, out_small, n, H, W, taps, pk.ntaps, conv_only, has_next, num_sms, s)
  if (pk.C == 128 && pk.CI == 32 && pk.CRES == 128) {
    if (conv_only)
      return launch_one<128, 32, 32, 128, 2, 2, true>(pk, e1, x, y, out_small, n, H, W, taps, pk.ntaps, 1, 0, num_sms, s);
    // three groups (D3 over D1's TMEM columns, one x / y buffer per group, no spare) measured SLOWER than two
    // groups with a spare residual buffer: 70.8 vs 65.8 us per launch; kept selectable for A/B runs
    if (cfg128 == 31) BC_LAUNCH(128, 32, 32, 128, 3, 1);
    if (cfg128 == 22)       // two warps per TMEM lane quarter: every epilogue of a tile is split over 8 warps
      return launch_one<128, 32, 32, 128, 2, 1, false, 2>(pk, e1, x, y, out_small, n, H, W, taps, pk.ntaps, 0, has_next, num_sms, s);
    BC_LAUNCH(128, 32, 32, 128, 2, 1);
  }
  if (pk.C == 64 && pk.CI == 16 && pk.CRES == 64) {
    // row-slab mode changed the balance: one CTA per SM with four groups (and an 18-slot ring = six tiles of
    // row slabs) now beats two CTAs with two groups each (3.27 vs 3.44 ms per 30 launches)
    if (cfg64 == 12) BC_LAUNCH(64, 16, 16, 64, 1, 2);
    if (cfg64 == 22) BC_LAUNCH(64, 16, 16, 64, 2, 2);
    BC_LAUNCH(64, 16, 16, 64, 4, 1);
  }
  if (pk.C == 64 && pk.CI == 16 && pk.CRES == 16) BC_LAUNCH(64, 16, 16, 16, 2, 2);     // downsample1_0
  if (pk.C == 128 && pk.CI == 16 && pk.CRES == 64) BC_LAUNCH(128, 16, 32, 64, 2, 1);   // downsample2_0
#undef BC_LAUNCH
  return cudaErrorInvalidValue;
}

}  // namespace bc
