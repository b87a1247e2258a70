// contour_noise_removal (image_processing_utils.py:4-44) as connected-component labelling.
//
// The reference closes the mask with a k x k box, extracts every contour (findContours,
// RETR_LIST), rasterises each one with fillPoly to measure its overlap with the bottom strip,
// and finally fills all kept contours in one even-odd fillPoly call.  None of that needs
// contours: a contour is the border between an 8-connected foreground component and one
// 4-connected background region, fillPoly of it covers exactly the pixels that region
// relation encloses (plus the border chain itself), so the whole function is
//
//   bit-pack -> binary close -> union-find labelling of BOTH classes in one label image
//   -> nesting tree (component -> enclosing region -> enclosing component ...)
//   -> strip counts accumulated up the tree -> per-pixel even-odd walk up the tree.
//
// oracle/contour_oracle.py states the same thing on the CPU and is pinned against the
// reference function; tests/test_gpu_contour.py compares this file with both.
//
// Data layout: masks as bit rows (ceil(W/32) words per row, 16 KB for a 256x512 frame, so
// the morphology is word-parallel and L2-resident), one int32 label per pixel (frame-relative
// index of the component's raster-first pixel, -1 = background connected to the image frame)
// and one 16-byte node record at the same index, touched only for tile-local roots.
// Labelling is two-level: 32 x 256 pixel tiles are resolved in shared memory (find chains and
// atomics never leave the SM), then only the links that cross a tile border touch global memory.
// Grids are (x blocks, rows, frames): index arithmetic stays in 32 bits.
#include "internal.h"
#include <algorithm>
#include <cmath>

namespace bc {
namespace {

struct __align__(16) Node { int up; int own; int enc; int ring; };
// up: enclosing node; own: strip pixels of the node itself; enc: strip pixels of everything it
// encloses (itself included); ring: strip pixels of the component ring around a hole

constexpr int TH = 32, TW = 256;          // labelling tile (rows x pixels); TW / 32 warps per CTA

// ------------------------------------------------------------------------ bit rows
// one thread packs 16 pixels, a lane pair one word
template <bool VEC>
__global__ void k_cn_pack(const uint8_t* __restrict__ seg, uint32_t* __restrict__ bits, int W, int WW, unsigned halves) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;       // half-word index over all rows
  unsigned row = i / (2u * WW);
  unsigned hx = i - row * (2u * WW);
  uint32_t m = 0;
  if (i < halves) {
    int x0 = (int)hx * 16;
    const uint8_t* src = seg + (size_t)row * W + x0;
    if (VEC) {
      if (x0 < W) {
        uint4 v = *reinterpret_cast<const uint4*>(src);
        uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j) m |= (uint32_t)(((w4[q] >> (8 * j)) & 0xffu) != 0u) << (4 * q + j);
      }
    } else {
      for (int j = 0; j < 16; ++j)
        if (x0 + j < W) m |= (uint32_t)(src[j] != 0) << j;
    }
  }
  uint32_t other = __shfl_xor_sync(0xffffffffu, m, 1);
  if (i < halves && !(hx & 1u)) bits[(size_t)row * WW + (hx >> 1)] = m | (other << 16);
}

// horizontal k-window of one bit row word: OR (dilate) or AND (erode) of the row shifted by
// -a .. k-1-a pixels.  Pixels outside the image never win (0 for OR, 1 for AND).
template <bool ERODE>
__device__ __forceinline__ uint32_t hwindow(const uint32_t* __restrict__ row, int wx, int WW, int W, int k, int a) {
  const uint32_t fill = ERODE ? 0xffffffffu : 0u;
  uint32_t lo = wx > 0 ? row[wx - 1] : fill;
  uint32_t mid = row[wx];
  uint32_t hi = wx + 1 < WW ? row[wx + 1] : fill;
  int tail = W & 31;
  if (ERODE && tail) {                       // bits past the last column count as "outside"
    uint32_t pad = ~((1u << tail) - 1u);
    if (wx == WW - 1) mid |= pad;
    if (wx + 1 == WW - 1) hi |= pad;
  }
  uint32_t acc = mid;
  // out(x) = op over d in [-a, k-1-a] of in(x + d)
  for (int d = 1; d <= a; ++d) {             // in(x - d): shift towards higher bit positions
    uint32_t v = __funnelshift_l(lo, mid, d);
    acc = ERODE ? (acc & v) : (acc | v);
  }
  for (int d = 1; d <= k - 1 - a; ++d) {     // in(x + d)
    uint32_t v = __funnelshift_r(mid, hi, d);
    acc = ERODE ? (acc & v) : (acc | v);
  }
  return acc;
}

// grid (ceil(H * WW / 256), 1, B)
template <bool ERODE>
__global__ void k_cn_morph(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int H, int W, int WW, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * WW) return;
  int y = i / WW;
  int wx = i - y * WW;
  const uint32_t* frame = in + (size_t)blockIdx.z * H * WW;
  int a = k / 2;
  uint32_t acc = ERODE ? 0xffffffffu : 0u;
  for (int dy = -a; dy <= k - 1 - a; ++dy) {
    int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    uint32_t v = hwindow<ERODE>(frame + (size_t)yy * WW, wx, WW, W, k, a);
    acc = ERODE ? (acc & v) : (acc | v);
  }
  int tail = W & 31;
  if (tail && wx == WW - 1) acc &= (1u << tail) - 1u;
  out[((size_t)blockIdx.z * H + y) * WW + wx] = acc;
}

// ------------------------------------------------------------------------ union-find
// parent pointers always point to a smaller index of the same set; -1 = exterior
__device__ __forceinline__ int uf_find(const int* L, int x) {
  while (true) {
    int l = L[x];
    if (l < 0) return -1;
    if (l == x) return x;
    x = l;
  }
}

__device__ __forceinline__ void uf_merge(int* L, int a, int b) {
  while (true) {
    a = a < 0 ? -1 : uf_find(L, a);
    b = b < 0 ? -1 : uf_find(L, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }      // a > b >= -1: hang a under b
    int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;                                      // somebody re-parented a meanwhile: carry on from there
  }
}

// the same on the tile's shared-memory labels; one pad word per 32 keeps the word-parallel
// phases (stride-32 accesses) off a single bank
__device__ __forceinline__ int& sref(int* sl, int i) { return sl[i + (i >> 5)]; }

__device__ __forceinline__ int sfind(int* sl, int x) {
  while (true) {
    int l = sref(sl, x);
    if (l < 0) return -1;
    if (l == x) return x;
    x = l;
  }
}

__device__ __forceinline__ void smerge(int* sl, int a, int b) {
  while (true) {
    a = a < 0 ? -1 : sfind(sl, a);
    b = b < 0 ? -1 : sfind(sl, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    int old = atomicMin(&sref(sl, a), b);
    if (old == a) return;
    a = old;
  }
}

// 34-bit window of a bit row around word wx: bit (lane + 1) = pixel x, bit lane = x - 1, bit (lane + 2) = x + 1
__device__ __forceinline__ uint64_t window(const uint32_t* __restrict__ row, int wx, int WW) {
  uint32_t lo = wx > 0 ? row[wx - 1] : 0u;
  uint32_t hi = wx + 1 < WW ? row[wx + 1] : 0u;
  return ((uint64_t)(hi & 1u) << 33) | ((uint64_t)row[wx] << 1) | (uint64_t)(lo >> 31);
}

__device__ __forceinline__ uint64_t window_valid(int wx, int W) {
  int lo = wx * 32 - 1;                        // pixel index of window bit 0
  int hi_px = W - lo;                          // window bits below hi_px are pixels
  uint64_t valid = hi_px >= 34 ? 0x3ffffffffull : ((1ull << hi_px) - 1ull);
  if (lo < 0) valid &= ~1ull;
  return valid;
}

// The links a pixel owns (to its upper / left neighbours), thinned so that every adjacency is
// still implied: one vertical link per run of vertical overlaps, the left link only where a run
// crosses a word boundary (inside a word the initial label is the run head), diagonal links
// (foreground only, 8-connectivity) only where no 4-connected detour exists.
struct Links { bool fg, vert, left, dul, dur; };

__device__ __forceinline__ Links links_of(uint64_t cur, uint64_t up, uint64_t valid, bool has_up, int lane) {
  Links l;
  int b = lane + 1;
  l.fg = (cur >> b) & 1ull;
  uint64_t curc = (l.fg ? cur : ~cur) & valid;
  uint64_t upc = has_up ? ((l.fg ? up : ~up) & valid) : 0ull;
  uint64_t v = curc & upc;
  l.vert = ((v >> b) & 1ull) && !(lane > 0 && ((v >> (b - 1)) & 1ull));
  l.left = lane == 0 && ((curc >> (b - 1)) & 1ull);
  bool up_bg = !((up >> b) & 1ull);
  l.dul = l.fg && has_up && up_bg && ((up >> (b - 1)) & 1ull) && !((cur >> (b - 1)) & 1ull);
  l.dur = l.fg && has_up && up_bg && ((up >> (b + 1)) & 1ull) && !((cur >> (b + 1)) & 1ull);
  return l;
}

// head (bit index) of the run of equal bits that contains bit b of word m
__device__ __forceinline__ int run_head(uint32_t m, int b) {
  uint32_t cls = ((m >> b) & 1u) ? m : ~m;
  uint32_t other_below = ~cls & ((1u << b) - 1u);
  return other_below ? 32 - __clz(other_below) : 0;
}

// Tile pass: grid (ceil(W / TW), ceil(H / TH), B), TW threads.  The per-pixel label array lives in
// shared memory; union-find runs on RUN HEADS only, one thread per 32-pixel word (TH * TW / 32
// = TW words per tile), with the same thinned link set as links_of() restricted to links that stay
// inside the tile.  Pixel-parallel phases (initial heads, final expansion) use thread = column.
__global__ void __launch_bounds__(TW) k_cn_tile(const uint32_t* __restrict__ bits, int* __restrict__ Lall,
                                                Node* __restrict__ nall, int H, int W, int WW) {
  constexpr int TWW = TW / 32;                      // words per tile row
  static_assert(TH * TWW == TW, "one thread per tile word");
  __shared__ int sl[TH * TW + TH * TW / 32];
  __shared__ uint32_t sb[TH + 1][TWW + 2];          // bit rows y0-1 .. y0+TH-1, words wx0-1 .. wx0+TWW (0 outside)
  const int tid = threadIdx.x, lane = tid & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, wx0 = x0 >> 5;
  const uint32_t* frame = bits + (size_t)blockIdx.z * H * WW;
  const int rows = min(TH, H - y0);

  for (int i = tid; i < (TH + 1) * (TWW + 2); i += TW) {
    int r = i / (TWW + 2), c = i - r * (TWW + 2);
    int y = y0 - 1 + r, wx = wx0 - 1 + c;
    sb[r][c] = (y >= 0 && y < H && wx >= 0 && wx < WW) ? frame[(size_t)y * WW + wx] : 0u;
  }
  __syncthreads();
  // every pixel points at the head of its run inside its word
  {
    const int wcol = tid >> 5;
    for (int ly = 0; ly < rows; ++ly)
      sref(sl, ly * TW + tid) = ly * TW + tid - lane + run_head(sb[ly + 1][wcol + 1], lane);
  }
  __syncthreads();
  // one thread per word: links of the runs that start in this word
  const int ly = tid / TWW, wcol = tid - ly * TWW;
  const int y = y0 + ly, wxg = wx0 + wcol;
  const bool word_live = ly < rows && wxg < WW;
  uint32_t vm = 0;
  if (word_live) {
    int npx = W - wxg * 32;
    vm = npx >= 32 ? 0xffffffffu : ((1u << npx) - 1u);
    const uint32_t m = sb[ly + 1][wcol + 1], lw = sb[ly + 1][wcol], rw = sb[ly + 1][wcol + 2];
    const uint32_t um = sb[ly][wcol + 1], ulw = sb[ly][wcol], urw = sb[ly][wcol + 2];
    const int base = ly * TW + wcol * 32, ubase = base - TW;
    const bool frame_row = y == 0 || y == H - 1;
    const int last_bit = W - 1 - wxg * 32;                 // bit of pixel x = W - 1 if it is in this word
#pragma unroll
    for (int c = 1; c >= 0; --c) {                         // c = 1 foreground, 0 background
      uint32_t rem = (c ? m : ~m) & vm;
      const uint32_t ucm = ly > 0 ? ((c ? um : ~um) & vm) : 0u;
      while (rem) {
        uint32_t low = rem & (0u - rem);
        uint32_t run = rem & ~(rem + low);                 // the lowest run of ones
        rem &= ~run;
        int s = __ffs(run) - 1, e = 31 - __clz(run);
        int h = base + s;
        // the run continues from the previous word of the tile
        if (s == 0 && wcol > 0 && ((lw >> 31) & 1u) == (uint32_t)c) smerge(sl, h, base - 32 + run_head(lw, 31));
        // one vertical link per run of vertical overlaps
        uint32_t v = run & ucm;
        while (v) {
          int b = __ffs(v) - 1;
          smerge(sl, h, ubase + run_head(um, b));
          v &= v + (v & (0u - v));
        }
        if (c && ly > 0) {
          // diagonal links only where no 4-connected detour exists: at the run's ends, under background
          if (!((um >> s) & 1u)) {
            if (s > 0) {
              if ((um >> (s - 1)) & 1u) smerge(sl, h, ubase + run_head(um, s - 1));
            } else if (wcol > 0 && (ulw >> 31) && !(lw >> 31)) {
              smerge(sl, h, ubase - 32 + run_head(ulw, 31));
            }
          }
          if (!((um >> e) & 1u)) {
            if (e < 31) {
              if ((um >> (e + 1)) & 1u) smerge(sl, h, ubase + e + 1);      // that bit starts a run
            } else if (wcol < TWW - 1 && (urw & 1u) && !(rw & 1u)) {
              smerge(sl, h, ubase + 32);
            }
          }
        }
        // background touching the frame = exterior
        if (!c && (frame_row || (wxg == 0 && s == 0) || e == last_bit)) smerge(sl, h, -1);
      }
    }
  }
  __syncthreads();
  // compress: every run head points at its root
  if (word_live) {
    const uint32_t m = sb[ly + 1][wcol + 1];
    uint32_t heads = (m ^ (m << 1) ^ (~m & 1u)) & vm;          // bit b set <=> bit b starts a run (bit 0 always)
    heads |= 1u;
    heads &= vm;
    const int base = ly * TW + wcol * 32;
    while (heads) {
      int b = __ffs(heads) - 1;
      heads &= heads - 1u;
      sref(sl, base + b) = sfind(sl, base + b);
    }
  }
  __syncthreads();
  const int x = x0 + tid;
  if (x >= W) return;
  int* L = Lall + (size_t)blockIdx.z * H * W;
  Node* nodes = nall + (size_t)blockIdx.z * H * W;
  for (int r_ = 0; r_ < rows; ++r_) {
    int li = r_ * TW + tid;
    int r = sref(sl, li);                            // run head (or root, or -1)
    if (r >= 0) r = sref(sl, r);                     // root of the head
    int p = (y0 + r_) * W + x;
    L[p] = r < 0 ? -1 : (y0 + (r / TW)) * W + x0 + (r % TW);
    if (r == li) nodes[p] = Node{-1, 0, 0, 0};
  }
}

// Border pass: the links that cross a tile border, on the global labels.
// grid.x covers nby * W row-border pixels, then nbx * 2 * H column-border pixels; grid.y = frame
__global__ void k_cn_border(const uint32_t* __restrict__ bits, int* __restrict__ Lall, int H, int W, int WW, int nby,
                            int nbx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int x, y;
  if (i < nby * W) {
    int j = i / W;
    x = i - j * W;
    y = (j + 1) * TH;
  } else {
    i -= nby * W;
    if (i >= nbx * 2 * H) return;
    int j = i / (2 * H);
    int rem = i - j * 2 * H;
    int side = rem >= H;
    y = rem - side * H;
    x = (j + 1) * TW - 1 + side;
    if (x >= W) return;
  }
  const uint32_t* frame = bits + (size_t)blockIdx.y * H * WW;
  int* L = Lall + (size_t)blockIdx.y * H * W;
  int wx = x >> 5, lane = x & 31;
  uint64_t cur = window(frame + (size_t)y * WW, wx, WW);
  uint64_t up = y > 0 ? window(frame + (size_t)(y - 1) * WW, wx, WW) : 0ull;
  Links k = links_of(cur, up, window_valid(wx, W), y > 0, lane);
  bool top = (y % TH) == 0, lft = (x % TW) == 0, rgt = (x % TW) == TW - 1;
  int p = y * W + x;
  if (k.vert && top) uf_merge(L, p, p - W);
  if (k.left && lft) uf_merge(L, p, p - 1);
  if (k.dul && (top || lft)) uf_merge(L, p, p - W - 1);
  if (k.dur && (top || rgt)) uf_merge(L, p, p - W + 1);
}

__device__ __forceinline__ bool bit_at(const uint32_t* __restrict__ frame, int WW, int x, int y) {
  return (frame[(size_t)y * WW + (x >> 5)] >> (x & 31)) & 1u;
}

// Flatten (4 pixels per thread); roots learn their enclosing node; strip pixels are counted.
// grid (ceil(W / 4 / 128), H, B)
template <bool VEC>
__global__ void k_cn_flatten(const uint32_t* __restrict__ bits, int* __restrict__ Lall, Node* __restrict__ nall, int H,
                             int W, int WW, int y_top) {
  int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x0 >= W) return;
  int y = blockIdx.y;
  int* L = Lall + (size_t)blockIdx.z * H * W;
  Node* nodes = nall + (size_t)blockIdx.z * H * W;
  const uint32_t* frame = bits + (size_t)blockIdx.z * H * WW;
  int p0 = y * W + x0;
  int n = min(4, W - x0);
  int l[4] = {-1, -1, -1, -1};
  if (VEC) {
    int4 v = *reinterpret_cast<const int4*>(L + p0);
    l[0] = v.x; l[1] = v.y; l[2] = v.z; l[3] = v.w;
  } else {
    for (int j = 0; j < n; ++j) l[j] = L[p0 + j];
  }
  uint32_t word = frame[(size_t)y * WW + (x0 >> 5)] >> (x0 & 31);     // 4 pixels never straddle a word (x0 % 4 == 0)
  int r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int p = p0 + j;
    if (j >= n || l[j] < 0) r[j] = -1;
    else if (l[j] == p) r[j] = p;
    else if (j > 0 && l[j] == l[j - 1]) r[j] = r[j - 1];          // same tile-level label as the pixel before
    else r[j] = uf_find(L, l[j]);
    if (j < n && r[j] == p) {
      // component: the pixel left of the raster-first pixel lies in the enclosing region;
      // hole: the pixel above the raster-first pixel belongs to the enclosing component
      bool fg = (word >> j) & 1u;
      nodes[p].up = fg ? (x0 + j == 0 ? -1 : uf_find(L, p - 1)) : uf_find(L, p - W);
    }
  }
  if (VEC) {
    *reinterpret_cast<int4*>(L + p0) = make_int4(r[0], r[1], r[2], r[3]);
  } else {
    for (int j = 0; j < n; ++j) L[p0 + j] = r[j];
  }
  if (y < y_top) return;
  // strip counts: one atomic per run of equal roots
  int run_r = -1, run_n = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < n && r[j] >= 0 && r[j] == run_r) { ++run_n; continue; }
    if (run_n) atomicAdd(&nodes[run_r].own, run_n);
    run_r = j < n ? r[j] : -1;
    run_n = run_r >= 0 ? 1 : 0;
  }
  if (run_n) atomicAdd(&nodes[run_r].own, run_n);
  // ring pixels: component pixels 4-adjacent to one of the component's own holes
  const int dx[4] = {0, 0, -1, 1}, dy[4] = {-1, 1, 0, 0};
  for (int j = 0; j < n; ++j) {
    if (!((word >> j) & 1u)) continue;
    int x = x0 + j;
    int seen[4];
    int ns = 0, upC = 0;
    bool have_up = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int xx = x + dx[q], yy = y + dy[q];
      if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
      if (bit_at(frame, WW, xx, yy)) continue;
      int h = uf_find(L, yy * W + xx);
      if (h < 0) continue;
      if (!have_up) { upC = r[j] % W == 0 ? -1 : uf_find(L, r[j] - 1); have_up = true; }   // region enclosing the component
      if (h == upC) continue;
      bool dup = false;
      for (int s = 0; s < ns; ++s) dup |= seen[s] == h;
      if (dup) continue;
      seen[ns++] = h;
      atomicAdd(&nodes[h].ring, 1);
    }
  }
}

// every node with strip pixels adds them to itself and to everything that encloses it; the strip
// pixel that zeroes the node's own count first does it.  grid (ceil(W / 128), strip rows, B)
__global__ void k_cn_accumulate(const int* __restrict__ Lall, Node* __restrict__ nall, int H, int W, int y_top) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  int y = y_top + blockIdx.y;
  const int* L = Lall + (size_t)blockIdx.z * H * W;
  Node* nodes = nall + (size_t)blockIdx.z * H * W;
  int r = L[y * W + x];
  if (r < 0) return;
  if (x > 0 && L[y * W + x - 1] == r) return;       // one candidate per run is enough
  int own = atomicExch(&nodes[r].own, 0);
  if (own == 0) return;
  for (int a = r; a >= 0; a = nodes[a].up) atomicAdd(&nodes[a].enc, own);
}

// Even-odd parity of the kept contours that enclose the pixels of node `lbl`, its own contour
// included: walk component -> hole -> component ... up to the exterior.
//
// This parity is the whole answer.  fillPoly also draws the contour chains themselves, which could
// only matter for a pixel that lies on a kept chain while its parity is 0.  That cannot happen: a
// contour's fill contains the fill of every contour nested inside it (a hole's fill, ring included,
// is a subset of its component's), so "kept" is monotone up the tree; a chain pixel belongs to a
// kept component (a kept hole implies its component is kept), all of whose ancestors are then kept
// too, and a component has an even number of ancestors: parity 1.
__device__ __forceinline__ int enclosing_parity(const Node* __restrict__ nodes, int lbl, bool fg, int thr) {
  int par = 0, a = lbl;
  if (fg) {
    Node c = nodes[lbl];
    par = c.enc >= thr;
    a = c.up;
  }
  while (a >= 0) {
    Node hn = nodes[a];
    par ^= (hn.enc + hn.ring >= thr);
    Node cn = nodes[hn.up];
    par ^= (cn.enc >= thr);
    a = cn.up;
  }
  return par;
}

// grid (ceil(W / 4 / 128), H, B); 4 pixels per thread
template <bool VEC>
__global__ void k_cn_fill(const uint32_t* __restrict__ bits, const int* __restrict__ Lall, const Node* __restrict__ nall,
                          uint8_t* __restrict__ out, int H, int W, int WW, int thr) {
  int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x0 >= W) return;
  int y = blockIdx.y;
  const int* L = Lall + (size_t)blockIdx.z * H * W;
  const Node* nodes = nall + (size_t)blockIdx.z * H * W;
  int p0 = y * W + x0;
  int n = min(4, W - x0);
  int l[4] = {-1, -1, -1, -1};
  if (VEC) {
    int4 v = *reinterpret_cast<const int4*>(L + p0);
    l[0] = v.x; l[1] = v.y; l[2] = v.z; l[3] = v.w;
  } else {
    for (int j = 0; j < n; ++j) l[j] = L[p0 + j];
  }
  uint32_t cur4 = (bits[((size_t)blockIdx.z * H + y) * WW + (x0 >> 5)] >> (x0 & 31)) & 0xfu;
  uint32_t res = 0;
  int memo_lbl = -2, memo = 0;               // the four pixels mostly share one label
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j >= n || l[j] < 0) continue;        // exterior background stays 0
    if (l[j] != memo_lbl) { memo = enclosing_parity(nodes, l[j], (cur4 >> j) & 1u, thr); memo_lbl = l[j]; }
    res |= (uint32_t)memo << (8 * j);
  }
  uint8_t* o = out + (size_t)blockIdx.z * H * W + p0;
  if (VEC) {
    *reinterpret_cast<uint32_t*>(o) = res;
  } else {
    for (int j = 0; j < n; ++j) o[j] = (uint8_t)(res >> (8 * j));
  }
}

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

size_t contour_scratch_bytes(int B, int H, int W) {
  size_t WW = (size_t)(W + 31) / 32;
  size_t px = (size_t)B * H * W;
  return 2 * up256((size_t)B * H * WW * 4) + up256(px * 4) + px * sizeof(Node);
}

int contour_launch_count() { return 8; }

void launch_contour_noise_removal(const uint8_t* seg, int H, int W, int B, uint8_t* out, void* scratch, cudaStream_t s) {
  int WW = (W + 31) / 32;
  size_t px = (size_t)B * H * W;
  size_t bits_bytes = up256((size_t)B * H * WW * 4);
  uint8_t* base = (uint8_t*)scratch;
  uint32_t* bitsA = (uint32_t*)base;
  uint32_t* bitsB = (uint32_t*)(base + bits_bytes);
  int* L = (int*)(base + 2 * bits_bytes);
  Node* nodes = (Node*)(base + 2 * bits_bytes + up256(px * 4));
  int k = (int)(std::min(H, W) / 50);                                  // image_processing_utils.py:7-8
  int y_top = (int)(H * (1 - 0.1));                                    // :19,22 (fp64, as Python evaluates it)
  double T = (double)((long long)W * (H - y_top)) * 0.4;               // :27,31,38
  int thr = (int)floor(T) + 1;                                         // area > T  <=>  area >= thr
  auto cdiv = [](long long a, long long b) { return (unsigned)((a + b - 1) / b); };

  unsigned halves = (unsigned)((size_t)B * H * WW * 2);
  bool vec16 = (W % 16 == 0) && ((uintptr_t)seg % 16 == 0);
  if (vec16) k_cn_pack<true><<<cdiv(halves, 256), 256, 0, s>>>(seg, bitsA, W, WW, halves);
  else k_cn_pack<false><<<cdiv(halves, 256), 256, 0, s>>>(seg, bitsA, W, WW, halves);
  dim3 gm(cdiv((long long)H * WW, 256), 1, B);
  k_cn_morph<false><<<gm, 256, 0, s>>>(bitsA, bitsB, H, W, WW, k);
  k_cn_morph<true><<<gm, 256, 0, s>>>(bitsB, bitsA, H, W, WW, k);
  k_cn_tile<<<dim3(cdiv(W, TW), cdiv(H, TH), B), TW, 0, s>>>(bitsA, L, nodes, H, W, WW);
  int nby = (H - 1) / TH, nbx = (W - 1) / TW;      // interior tile borders
  long long nborder = (long long)nby * W + (long long)nbx * 2 * H;
  if (nborder > 0) k_cn_border<<<dim3(cdiv(nborder, 128), B), 128, 0, s>>>(bitsA, L, H, W, WW, nby, nbx);
  else k_cn_border<<<dim3(1, 1), 32, 0, s>>>(bitsA, L, H, W, WW, 0, 0);     // keeps the launch count fixed
  dim3 g4(cdiv(cdiv(W, 4), 128), H, B);
  bool vec4 = (W % 4 == 0) && ((uintptr_t)out % 4 == 0);
  if (vec4) k_cn_flatten<true><<<g4, 128, 0, s>>>(bitsA, L, nodes, H, W, WW, y_top);
  else k_cn_flatten<false><<<g4, 128, 0, s>>>(bitsA, L, nodes, H, W, WW, y_top);
  k_cn_accumulate<<<dim3(cdiv(W, 128), H - y_top, B), 128, 0, s>>>(L, nodes, H, W, y_top);
  if (vec4) k_cn_fill<true><<<g4, 128, 0, s>>>(bitsA, L, nodes, out, H, W, WW, thr);
  else k_cn_fill<false><<<g4, 128, 0, s>>>(bitsA, L, nodes, out, H, W, WW, thr);
}

}  // namespace bc
