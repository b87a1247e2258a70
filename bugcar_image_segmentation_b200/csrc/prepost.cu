// Pre/post-processing kernels of the hot path (CUDA cores; integer / fp64 work).
//
//   K1  k_resize        cv2.resize(bgr,(512,256)) bit-exact            models.py:87
//       k_preprocess    BGR->RGB, (u/256-mean)/std, HWC->CHW           models.py:89-94
//   K8' k_argmax_lut    tf.math.argmax(axis=1) + class LUT             models.py:55-58,67 / 78-81
//   K9  k_occ_table     per calibration: nearest-resize + crop/paste + warpPerspective
//                       coordinates of every grid cell's 5x5 template neighbourhood
//       k_occgrid       per frame: bilinear label blend + 3x3 open + int8 map,
//                       one kernel, no intermediate image              bev.py:166-246 / 97-144
//
// All of it is byte/integer arithmetic with fp64 coordinates (K9: 131 072 B labels in, Hc*Wc B out per
// frame; a gather of up to 25 x 4 label bytes per cell, bound by load latency and issue slots, not by HBM).  Semantics follow OpenCV 4.x as
// restated (and pinned against cv2) in oracle/cv_ops.py.
#include "internal.h"

#include <algorithm>
#include <cstdlib>

namespace bc {

static constexpr int NET_W = 512, NET_H = 256;

// ------------------------------------------------------------------------------ K1
// mode 0 identity copy; mode 1 area 2x2: (a+b+c+d+2)>>2 (cv::resize swaps INTER_LINEAR
// for INTER_AREA when both scales are exactly 2); mode 2 the 11-bit fixed-point separable
// bilinear kernel (HResizeLinear / VResizeLinear, INTER_RESIZE_COEF_BITS = 11).
__global__ void __launch_bounds__(256)
k_resize(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst,
         ResizeTab t, int total_px) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;   // destination pixel over B*256*512
  if (p >= total_px) return;
  int x = p % NET_W;
  int y = (p / NET_W) % NET_H;
  int n = p / (NET_W * NET_H);
  const uint8_t* s = src + (size_t)n * sh * sw * 3;
  uint8_t* d = dst + (size_t)p * 3;
  if (t.mode == 1) {
    const uint8_t* r0 = s + ((size_t)(2 * y) * sw + 2 * x) * 3;
    const uint8_t* r1 = r0 + (size_t)sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      d[c] = (uint8_t)((r0[c] + r0[3 + c] + r1[c] + r1[3 + c] + 2) >> 2);
    return;
  }
  int x0 = t.x0[x], x1 = t.x1[x], a0 = t.a0[x], a1 = t.a1[x];
  int y0 = t.y0[y], y1 = t.y1[y], b0 = t.b0[y], b1 = t.b1[y];
  const uint8_t* r0 = s + (size_t)y0 * sw * 3;
  const uint8_t* r1 = s + (size_t)y1 * sw * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int S0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
    int S1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
    int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
    d[c] = (uint8_t)min(max(v, 0), 255);
  }
}

void launch_resize(const uint8_t* src, int h, int w, int B, uint8_t* dst, const ResizeTab& t,
                   cudaStream_t s) {
  if (t.mode == 0) {
    cudaMemcpyAsync(dst, src, (size_t)B * NET_H * NET_W * 3, cudaMemcpyDeviceToDevice, s);
    return;
  }
  int total = B * NET_H * NET_W;
  k_resize<<<(total + 255) / 256, 256, 0, s>>>(src, h, w, dst, t, total);
}

// ----------------------------------------------------------------- preprocess (API)
// out[n][c][y][x] = lut64[bgr[n][y][x][2-c]][c]; lut64 is the fp64 table
// ((u/256.0)-mean[c])/std[c] computed on the host exactly as numpy does (models.py:91).
template <typename OUT>
__global__ void __launch_bounds__(256)
k_preprocess(const uint8_t* __restrict__ bgr, OUT* __restrict__ out,
             const double* __restrict__ lut64, int total_px) {
  __shared__ double slut[256 * 3];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) slut[i] = lut64[i];
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_px) return;
  int n = p / (NET_W * NET_H);
  int yx = p % (NET_W * NET_H);
  const uint8_t* s = bgr + (size_t)p * 3;
  OUT* o = out + (size_t)n * 3 * NET_W * NET_H + yx;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[(size_t)c * NET_W * NET_H] = (OUT)slut[s[2 - c] * 3 + c];
}

void launch_preprocess(const uint8_t* bgr256, int B, void* out, int out_f64, const double* lut64,
                       cudaStream_t s) {
  int total = B * NET_H * NET_W;
  if (out_f64)
    k_preprocess<double><<<(total + 255) / 256, 256, 0, s>>>(bgr256, (double*)out, lut64, total);
  else
    k_preprocess<float><<<(total + 255) / 256, 256, 0, s>>>(bgr256, (float*)out, lut64, total);
}

// ------------------------------------------------------------------- argmax + LUT
// logits NCHW fp32; first maximum wins (tf.math.argmax / np.argmax tie-break).
__global__ void __launch_bounds__(256)
k_argmax_lut(const float* __restrict__ logits, int C, int HW, Lut256 lut,
             uint8_t* __restrict__ labels, int total_px) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_px) return;
  int n = p / HW, yx = p % HW;
  const float* l = logits + (size_t)n * C * HW + yx;
  float best = l[0];
  int bi = 0;
  for (int c = 1; c < C; ++c) {
    float v = l[(size_t)c * HW];
    if (v > best) { best = v; bi = c; }
  }
  labels[p] = lut.v[bi];
}

void launch_argmax_lut(const float* logits, int B, int C, int H, int W, const Lut256& lut,
                       uint8_t* labels, cudaStream_t s) {
  int total = B * H * W;
  k_argmax_lut<<<(total + 255) / 256, 256, 0, s>>>(logits, C, H * W, lut, labels, total);
}

// ------------------------------------------------------------------------------ K9
// The geometry of K9 does not depend on the frame: which template pixel a grid cell takes,
// which 5x5 template pixels the 3x3 opening looks at, and where each of those falls in the
// label map are functions of the calibration only.  k_occ_table evaluates them ONCE per
// (calibration, grid request) into a table; the per-frame kernel is then a pure integer
// gather + blend + bit logic on the label bytes.
//
// Table entry (16 bytes) for cell `c`, neighbourhood position `pos = j*5+i` (template pixel
// (ty-2+j, tx-2+i)), stored [pos][cell] so that a warp reads it coalesced:
//   .x = offset of the (clamped) top-left source pixel of the bilinear footprint in the label map
//   .y = w00 | w01 << 16, .z = w10 | w11 << 16     tap weights wy*wx (0..1024); a tap outside the
//                                                  label map (BORDER_CONSTANT 0), a pixel outside the
//                                                  pasted crop or outside the template: weight 0
//   .w = bit 0: +1 reaches the right tap, bit 1: +cols reaches the lower tap (0 when clamped),
//        bit 2: inside the template
// Fixed-point source coordinates of warped pixel (x, y): OpenCV's warpPerspective,
// INTER_LINEAR, 5-bit fractions.  The fp64 association (block origin xb) and the absence of
// FMA contraction reproduce WarpPerspectiveInvoker exactly (oracle/cv_ops.py warp_coords_fixed).
__device__ __forceinline__ void warp_coords(const BevGeom& g, int x, int y, int& X, int& Y) {
  int xb = (x / g.bw0) * g.bw0;
  double dxb = (double)xb, dx1 = (double)(x - xb), dy = (double)y;
  double X0 = __dadd_rn(__dadd_rn(__dmul_rn(g.Mi[0], dxb), __dmul_rn(g.Mi[1], dy)), g.Mi[2]);
  double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(g.Mi[3], dxb), __dmul_rn(g.Mi[4], dy)), g.Mi[5]);
  double W0 = __dadd_rn(__dadd_rn(__dmul_rn(g.Mi[6], dxb), __dmul_rn(g.Mi[7], dy)), g.Mi[8]);
  double W = __dadd_rn(W0, __dmul_rn(g.Mi[6], dx1));
  W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
  double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(g.Mi[0], dx1)), W);
  double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(g.Mi[3], dx1)), W);
  fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
  fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
  X = __double2int_rn(fX);      // cvRound: round half to even
  Y = __double2int_rn(fY);
}

// Table entry (8 bytes) of one template pixel: its bilinear footprint in the label map as a fixed 2 x 2 block of
// label pixels (rows r, r + 1, columns b, b + 1, always inside the image) and the separable weights of those
// four pixels -- zero for taps outside the image (cv::remap BORDER_CONSTANT 0), moved to the other column / row
// where the clamped block starts one pixel off the footprint (first and last column / row):
//   .x = r * cols + b
//   .y = wx0 | wx1 << 8 | wy0 << 16 | wy1 << 24      (weights 0..32)
// The 25 entries of a cell are followed by one word per cell whose bit `pos` says that position lies in the
// template (layout: uint2 [25][cells], uint32 [cells]).
__device__ __forceinline__ void occ_axis(int s, int a, int n, int& base, unsigned& w0, unsigned& w1) {
  base = 0; w0 = 0; w1 = 0;
  if (s >= 0 && s + 1 < n) { base = s; w0 = 32 - a; w1 = a; }
  else if (s == -1) { w0 = a; }                               // only pixel 0 (the footprint's second tap) is inside
  else if (s == n - 1) { base = n - 2; w1 = 32 - a; }         // only pixel n-1 (the footprint's first tap) is inside
}

__global__ void __launch_bounds__(128)
k_occ_table(const BevGeom g, uint2* __restrict__ table, unsigned* __restrict__ inside) {
  const int cells = g.Hc * g.Wc;
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  int pos = blockIdx.y;
  if (cell >= cells) return;
  int cx = cell % g.Wc, cy = cell / g.Wc;
  // cv::resize INTER_NEAREST (bev.py:209-212): min(floor(d * ifx), src - 1) in fp64
  int tx = min((int)floor(__dmul_rn((double)cx, g.ifx)), g.occ_w_px - 1);
  int ty = min((int)floor(__dmul_rn((double)cy, g.ify)), g.occ_h_px - 1);
  int x = tx - 2 + pos % 5, y = ty - 2 + pos / 5;
  uint2 e = make_uint2(0u, 0u);
  if (x >= 0 && x < g.occ_w_px && y >= 0 && y < g.occ_h_px) {
    atomicOr(&inside[cell], 1u << pos);
    // crop of the warped image pasted into a zero template (bev.py:183-195)
    int px = x - g.gl, py = y - g.gt;
    if (px >= 0 && py >= 0 && px < g.crop_w && py < g.crop_h) {
      int X, Y;
      warp_coords(g, px + g.wl, py + g.wt, X, Y);
      const int sx = min(max(X >> 5, -32768), 32767);   // saturate_cast<short>
      const int sy = min(max(Y >> 5, -32768), 32767);
      int bx, by;
      unsigned wx0, wx1, wy0, wy1;
      occ_axis(sx, X & 31, g.in_cols, bx, wx0, wx1);
      occ_axis(sy, Y & 31, g.in_rows, by, wy0, wy1);
      e.x = (unsigned)(by * g.in_cols + bx);
      e.y = wx0 | (wx1 << 8) | (wy0 << 16) | (wy1 << 24);
    }
  }
  table[(size_t)pos * cells + cell] = e;
}

size_t occ_table_bytes(int cells) { return (size_t)cells * (25 * sizeof(uint2) + sizeof(unsigned)); }

void launch_occ_table(const BevGeom& g, uint2* table, cudaStream_t s) {
  const int cells = g.Hc * g.Wc;
  unsigned* inside = reinterpret_cast<unsigned*>(table + (size_t)25 * cells);
  cudaMemsetAsync(inside, 0, (size_t)cells * sizeof(unsigned), s);
  dim3 grid((cells + 127) / 128, 25);
  k_occ_table<<<grid, 128, 0, s>>>(g, table, inside);
}

// (labels + 1) blended at the table entry's footprint: (sum p * wy * wx + 512) >> 10, evaluated as
// wy0 * (wx0 p00 + wx1 p01) + wy1 * (wx0 p10 + wx1 p11) -- the same integer -- in three dp2a;
// np.add(segmap, 1) wraps in uint8 (bev.py:177).  `nb` = frame index * rows * cols (32 bits: launch_occgrid
// splits batches whose label maps exceed 4 GB).
__device__ __forceinline__ unsigned occ_sample(const uint8_t* __restrict__ labels, unsigned nb, uint2 e, int cols) {
  const uint8_t* p0 = labels + (nb + e.x);
  const uint8_t* p1 = p0 + cols;
  // a row's two pixels as the 16-bit halves of one word, + 1 each with uint8 wrap-around
  const unsigned top = (((unsigned)p0[1] << 16 | (unsigned)p0[0]) + 0x00010001u) & 0x00ff00ffu;
  const unsigned bot = (((unsigned)p1[1] << 16 | (unsigned)p1[0]) + 0x00010001u) & 0x00ff00ffu;
  const unsigned t = __dp2a_lo(top, e.y, 0u);          // p00 * wx0 + p01 * wx1   (<= 255 * 32)
  const unsigned b = __dp2a_lo(bot, e.y, 0u);
  return __dp2a_hi(b << 16 | t, e.y, 512u) >> 10;      // t * wy0 + b * wy1 + 512
}

__device__ __forceinline__ bool is_occ(unsigned v, int binary) {
  return binary ? (v == 1u) : ((v | 2u) == 3u);     // 1, or 1 / 3: bev.py:128 / bev.py:196
}

// One thread per grid cell, `fpb` frames per block.  The cell takes template pixel (ty, tx) (nearest
// resize); if that pixel is "occupied" the 3x3 opening decides whether it is a speck:
//   opened(p) = OR_{q in N3(p)} AND_{r in N3(q)} occ(r)   (erode ignores pixels outside
//   the template, dilate treats them as 0 -- OpenCV default border values)
// evaluated lazily: q = p needs only the inner 3x3 ring; the outer ring of the 5x5 block is
// sampled only when that fails (the interior of an occupied region never gets there).
//
// The kernel is a chain of dependent gathers (centre -> inner ring -> outer ring), so what it needs is
// loads in flight: the nine inner table entries of the cell live in registers for all frames of the block
// (no shared memory: 8 blocks of 128 threads per SM), the centre samples of OCC_CH frames are issued
// together before any of them is consumed, and the rarely needed outer-ring entries come from the table in
// global memory (L2 resident, 2 MB).  Blocks stay small: the cost of a cell depends on the scene (occupied
// cells sample up to 25 template pixels, free ones a single one), so many short blocks balance themselves
// over the SMs; one resident wave of 37-frame blocks measured 129 us against 105 us.
static constexpr unsigned OCC_INNER = (7u << 6) | (7u << 11) | (7u << 16);     // 3x3 block around bit 12
static constexpr int OCC_CH = 4;             // frames whose centre samples are in flight together
static constexpr int OCC_MINB = 8;           // resident blocks per SM the register budget is set for

// outer ring of the 5x5 block, k = 0..15 -> bit position
__device__ __forceinline__ int occ_outer_pos(int k) {
  return k < 5 ? k : (k < 11 ? 5 * (((k - 5) >> 1) + 1) + ((k - 5) & 1) * 4 : k + 9);
}

__global__ void __launch_bounds__(128, OCC_MINB)
k_occgrid(const uint8_t* __restrict__ labels, const BevGeom g, int B, int fpb, int8_t* __restrict__ grids) {
  const uint2* __restrict__ table = g.table;
  const int cells = g.Hc * g.Wc;
  const int cell = blockIdx.x * 128 + threadIdx.x;
  const bool live = cell < cells;
  const int lane = threadIdx.x & 31, half = lane >> 4, k16 = lane & 15;
  const unsigned half_mask = half ? 0xffff0000u : 0x0000ffffu;
  uint2 ein[9];                    // inner 3x3 block, [j * 3 + i] <-> position (1 + j) * 5 + (1 + i)
#pragma unroll
  for (int k = 0; k < 9; ++k)
    ein[k] = live ? table[(size_t)((1 + k / 3) * 5 + 1 + k % 3) * cells + cell] : make_uint2(0u, 0u);
  const unsigned inside = live ? reinterpret_cast<const unsigned*>(table + (size_t)25 * cells)[cell] : 0u;
  const int cx = cell % g.Wc, cy = cell / g.Wc;
  const size_t o = g.ros_layout ? (size_t)(g.Wc - 1 - cx) * g.Hc + (g.Hc - 1 - cy)   // occgrid_to_ros.py:18-21
                                : (size_t)cell;
  const int cols = g.in_cols;
  const unsigned frame = (unsigned)g.in_rows * (unsigned)cols;
  const int n1 = min(B, (int)(blockIdx.y + 1) * fpb);
  // what this lane does when it helps with another lane's outer ring / opening test
  const int my_outer = occ_outer_pos(k16);
  const uint2* __restrict__ my_outer_row = table + (size_t)my_outer * cells + (cell - lane);
  unsigned my_q = 0, my_nb = 0;                    // lanes 0..8 of each half: one erosion centre q each
  if (k16 < 9) {
    const int qj = 1 + k16 / 3, qi = 1 + k16 % 3;
    my_q = 1u << (qj * 5 + qi);
    for (int rj = -1; rj <= 1; ++rj)
      for (int ri = -1; ri <= 1; ++ri) my_nb |= 1u << ((qj + rj) * 5 + (qi + ri));
  }
  for (int n0 = blockIdx.y * fpb; n0 < n1; n0 += OCC_CH) {
    unsigned vcp = 0;                              // the centre values (<= 255) of OCC_CH frames, one byte each
#pragma unroll
    for (int c = 0; c < OCC_CH; ++c)               // a frame past the end repeats the last one (loads stay in bounds)
      vcp |= occ_sample(labels, (unsigned)min(n0 + c, n1 - 1) * frame, ein[4], cols) << (8 * c);
#pragma unroll 1
    for (int c = 0; c < OCC_CH; ++c) {
      const int n = n0 + c;
      if (n >= n1) break;                          // block-uniform
      const unsigned nb = (unsigned)n * frame;
      unsigned v = (vcp >> (8 * c)) & 255u;
      const bool occupied = live && is_occ(v, g.binary);
      unsigned occ = 1u << 12;     // 5x5 occupancy mask around (ty, tx); bit (j*5+i) <-> (ty-2+j, tx-2+i)
      bool opened = true;
      if (occupied) {
        // positions outside the template have all-zero entries: they sample 0, "not occupied", and are
        // masked out by `inside` below
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          if (k == 4) continue;
          if (is_occ(occ_sample(labels, nb, ein[k], cols), g.binary)) occ |= 1u << ((1 + k / 3) * 5 + 1 + k % 3);
        }
        opened = ((~occ) & inside & OCC_INNER) == 0;            // q = p: every in-template neighbour occupied
      }
      // The undecided cells (occupied, but not the centre of a full 3x3 block) need the outer ring.
      // They are few, so the warp serves them two at a time: 16 lanes sample one outer pixel each,
      // 9 lanes test one erosion centre each.
      unsigned todo = __ballot_sync(0xffffffffu, occupied && !opened);
      while (todo) {
        const int s0 = __ffs(todo) - 1;
        todo &= todo - 1;
        const int s1 = todo ? __ffs(todo) - 1 : -1;
        if (s1 >= 0) todo &= todo - 1;
        const int src = half ? s1 : s0;
        const bool helping = src >= 0;
        const int srcl = helping ? src : s0;
        const unsigned s_inside = __shfl_sync(0xffffffffu, inside, srcl);
        const unsigned s_occ = __shfl_sync(0xffffffffu, occ, srcl);
        unsigned bit = 0;
        if (helping && ((s_inside >> my_outer) & 1u))
          if (is_occ(occ_sample(labels, nb, my_outer_row[srcl], cols), g.binary)) bit = 1u << my_outer;
        const unsigned all = s_occ | __reduce_or_sync(half_mask, bit);
        // opened(p) = OR_q AND_{r in N3(q)} occ(r): dilate ignores q outside the template, erode ignores r outside
        const bool ok = helping && my_q && (s_inside & my_q) && (((~all) & s_inside & my_nb) == 0);
        const unsigned okb = __ballot_sync(0xffffffffu, ok);     // the owner may sit in the other half-warp
        if (lane == s0) opened = (okb & 0xffffu) != 0;
        if (lane == s1) opened = (okb >> 16) != 0;
      }
      if (occupied && !opened) v = 2;                   // bev.py:203-205
      unsigned out;
      if (g.raw_template) {
        out = v;                                        // laserscan branch: the template itself (bev.py:209-212)
      } else if (g.binary) {
        const unsigned m = (v * 100u) & 255u;           // uint8 * 100 (bev.py:139-142)
        out = (m == 0u) ? 255u : ((200u - m) & 255u);   // bev.py:143-144
      } else {
        if (v == 3u) v = 1u;                            // bev.py:242
        out = (v == 0u) ? 255u : ((200u - ((v * 100u) & 255u)) & 255u);   // bev.py:244-245
      }
      if (live) grids[(size_t)n * cells + o] = (int8_t)out;
    }
  }
}

// kept for bc_create's per-device set-up list; the kernel no longer needs a shared-memory opt-in
cudaError_t prepare_occgrid() { return cudaSuccess; }

void launch_occgrid(const uint8_t* labels, int B, const BevGeom& g, int8_t* grids, cudaStream_t s) {
  static const int fpb_env = getenv("BC_OCC_FPB") ? atoi(getenv("BC_OCC_FPB")) : 0;     // tuning knob, read once
  const int cell_blocks = (g.Hc * g.Wc + 127) / 128;
  // Frames per block: a small multiple of OCC_CH, chosen so that the grid is close to a whole number of waves
  // of resident blocks (OCC_MINB per SM)
  int fpb = fpb_env;
  if (fpb <= 0) {
    static int slots = 0;
    if (!slots) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      slots = OCC_MINB * sms;
    }
    double best = -1.0;
    for (int f = OCC_CH; f <= 4 * OCC_CH; f += OCC_CH) {
      const long long blocks = (long long)cell_blocks * ((B + f - 1) / f);
      const long long waves = (blocks + slots - 1) / slots;
      const double eff = (double)blocks / (double)(waves * slots) * ((double)cell_blocks * B / (blocks * (double)f));
      if (eff > best + 1e-9) { best = eff; fpb = f; }
    }
  }
  // the kernel indexes the label maps with 32-bit offsets: at most 4 GB of them per launch
  const size_t frame = (size_t)g.in_rows * g.in_cols;
  const int per_launch = (int)std::min<size_t>((size_t)B, 0xffffffffull / frame);
  for (int b0 = 0; b0 < B; b0 += per_launch) {
    const int nb = std::min(per_launch, B - b0);
    dim3 grid(cell_blocks, (nb + fpb - 1) / fpb);
    k_occgrid<<<grid, 128, 0, s>>>(labels + (size_t)b0 * frame, g, nb, fpb, grids + (size_t)b0 * g.Hc * g.Wc);
  }
}

// ------------------------------------------------------------------ streaming gather flags
// One thread per flag.  The store is a system-scope release: everything this GPU wrote before the kernel
// (stream order), K9's peer stores included, is visible to whoever acquires the flag.
__global__ void k_flag_store(uint32_t* const* ptrs, uint32_t* single, int n, uint32_t value) {
  const int i = threadIdx.x;
  if (i >= n) return;
  uint32_t* p = single ? single : ptrs[i];
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(value) : "memory");
}
// Waits until every flag has reached `value` (flags only grow).  Bounded: after ~10 s the kernel gives up and
// raises *d_err, so a dead peer cannot hang the GPU.
__global__ void k_flag_wait(const uint32_t* flags, int n, uint32_t value, int* d_err) {
  const int i = threadIdx.x;
  if (i >= n) return;
  const long long t0 = clock64();
  uint32_t v;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
    if ((int)(v - value) >= 0) break;
    __nanosleep(200);
    if (clock64() - t0 > 20000000000LL) {            // ~10 s at 1.9 GHz
      *(volatile int*)d_err = 1;
      __threadfence_system();
      break;
    }
  }
}
void launch_flag_store(uint32_t* const* ptrs, int n, uint32_t value, cudaStream_t s) {
  k_flag_store<<<1, 32, 0, s>>>(ptrs, nullptr, n, value);
}
void launch_flag_store1(uint32_t* ptr, uint32_t value, cudaStream_t s) { k_flag_store<<<1, 32, 0, s>>>(nullptr, ptr, 1, value); }
void launch_flag_wait(const uint32_t* flags, int n, uint32_t value, int* d_err, cudaStream_t s) {
  k_flag_wait<<<1, 32, 0, s>>>(flags, n, value, d_err);
}

}  // namespace bc
