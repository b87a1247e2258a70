// Device helpers shared by the tcgen05 kernels: mbarrier, TMA (tensor + bulk), tcgen05
// MMA / commit / TMEM load, shared-memory and instruction descriptors, swizzle.
//
// The translation units that include this header are compiled TWICE (Makefile): with
// -DBC_ACT_F16=0 the 16-bit activation / operand type is bf16 (namespace bc::as_bf16), with
// -DBC_ACT_F16=1 it is IEEE fp16 (namespace bc::as_f16; same bytes, same TMA boxes, same
// tcgen05.mma.kind::f16 with another format field, 8x less rounding error).  Everything that
// depends on the type lives in namespace bc::BC_NS; the entry points are the members of
// Umma<act_t> (internal.h).
#pragma once
#include "internal.h"

#include <cuda.h>
#include <cstring>

#ifndef BC_ACT_F16
#define BC_ACT_F16 0
#endif
#if BC_ACT_F16
#define BC_NS as_f16
#else
#define BC_NS as_bf16
#endif

namespace bc {

// =================================================================== device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// non-blocking probe (event loops that watch several barriers)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 or bf16 operands per the instruction descriptor) -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of warp w reads lane 32*(w%4)+t, columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- "elected" variants for warp-uniform service loops -----------------------------------------
// The whole warp runs the loop convergently and ONE elected lane issues the instruction (elect.sync
// inside the asm).  Under `if (lane == 0)` the compiler treats the code as divergent and wraps
// every uniform-datapath instruction (UTCHMMA, UTCBAR, UTMALDG) in an ELECT / BRA.U.ANY loop and
// recomputes its operands per lane; a service thread is on the critical path of every tile, so
// its instruction count matters.
__device__ __forceinline__ void umma_mma_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_e(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_e(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                              uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_e(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                              uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_e(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// ---- the same copies with an L2 eviction-priority hint (createpolicy encodings, fraction 1.0)
static constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull, L2_EVICT_FIRST = 0x12F0000000000000ull,
                          L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_e(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint64_t pol) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_e(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                              uint32_t bar, uint64_t pol) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_global_hint(uint4* p, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset (ignored for swizzled K-major; 1 like CUTLASS), stride byte offset =
// 8 rows x row pitch, version 1 (Blackwell), layout type 2/4/6 = 128/64/32-byte swizzle.
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  constexpr uint64_t layout = ROW_BYTES == 128 ? 2 : ROW_BYTES == 64 ? 4 : 6;
  constexpr uint64_t sbo = (8 * ROW_BYTES) >> 4;
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor) of kind::f16: fp32 accumulate (bit 4), A / B
// format at bits 7 / 10 (0 = fp16, 1 = bf16), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc_fmt(int M, int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Swizzle<B,4,3>: 16-byte chunk index (address bits [4,4+B)) ^= address bits [7,7+B)
template <int ROW_BYTES>
__device__ __host__ __forceinline__ uint32_t swz(uint32_t off) {
  constexpr uint32_t mask = ROW_BYTES == 128 ? 7 : ROW_BYTES == 64 ? 3 : 1;
  return off ^ (((off >> 7) & mask) << 4);
}

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }
// two fp32 -> packed bf16x2 (a in the low half)
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- everything that depends on the 16-bit activation type ------------------------------------
namespace BC_NS {
#if BC_ACT_F16
typedef f16 act_t;
typedef __half2 act2_t;
constexpr uint32_t ACT_FMT = 0u;
constexpr CUtensorMapDataType ACT_TMAP = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
// round to nearest even, saturating at +-65504 instead of overflowing to infinity
__device__ __forceinline__ uint32_t pack_act(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack_act(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
inline uint16_t host_act_bits(float v) {
  const float lim = 65504.f;
  __half h = __float2half_rn(v > lim ? lim : v < -lim ? -lim : v);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
#else
typedef bf16 act_t;
typedef __nv_bfloat162 act2_t;
constexpr uint32_t ACT_FMT = 1u;
constexpr CUtensorMapDataType ACT_TMAP = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
__device__ __forceinline__ uint32_t pack_act(float a, float b) { return pack_bf16(a, b); }
__device__ __forceinline__ float2 unpack_act(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); }
inline uint16_t host_act_bits(float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
#endif
// 0xFFFF per 16-bit lane where b > a (ordered compare)
__device__ __forceinline__ uint32_t gt2_mask(uint32_t b, uint32_t a) {
  return __hgt2_mask(*reinterpret_cast<const act2_t*>(&b), *reinterpret_cast<const act2_t*>(&a));
}
// A and B operands in the activation type
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) { return instr_desc_fmt(M, N, ACT_FMT); }
// eight fp32 -> one 16-byte chunk of packed activations
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_act(v[0], v[1]), pack_act(v[2], v[3]), pack_act(v[4], v[5]), pack_act(v[6], v[7]));
}

// host: TMA tensor maps (enet_umma.cu)
bool make_map_e1(CUtensorMap* m, const act_t* base, int N, int H, int W, int CI);   // 4D [N][H][W][CI], box = 128 px
bool make_map_box(CUtensorMap* m, const act_t* base, int N, int H, int W, int CI, int box_w, int box_rows);
bool make_map_x(CUtensorMap* m, const act_t* base, size_t pixels, int C);           // 2D [px][C], box [128][64]
bool make_map_rows(CUtensorMap* m, const act_t* base, size_t pixels, int C, int box_px, int sw);   // 2D [px][C], box [box_px][C]
// 5D view [N*Ho][2][Wo][2][C] of a [N][2Ho][2Wo][C] tensor (2x2 stride-2 windows), box [rows][1][box_w][1][C]
bool make_map_window(CUtensorMap* m, const act_t* base, int N, int Ho, int Wo, int C, int rows, int box_w);

}  // namespace BC_NS
}  // namespace bc
