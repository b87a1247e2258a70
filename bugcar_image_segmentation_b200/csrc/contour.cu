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
// and one 16-byte node record at the same index, touched only for run heads.
#include "internal.h"
#include <algorithm>
#include <cmath>

namespace bc {
namespace {

struct Node { int up; int own; int enc; int ring; };   // enclosing node, strip pixels of the node itself,
                                                       // strip pixels of everything it encloses, strip pixels
                                                       // of the component ring around a hole

// ------------------------------------------------------------------------ bit rows
__global__ void k_cn_pack(const uint8_t* __restrict__ seg, uint32_t* __restrict__ bits, int H, int W, int WW,
                          long long rows) {
  // one warp per 32-pixel word; 4 words per warp iteration would not matter: the mask is 33 MB at bs 256
  long long word = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (word >= rows * WW) return;
  long long row = word / WW;
  int wx = (int)(word - row * WW);
  int x = wx * 32 + lane;
  bool on = x < W && seg[row * W + x] != 0;
  uint32_t m = __ballot_sync(0xffffffffu, on);
  if (lane == 0) bits[word] = m;
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

template <bool ERODE>
__global__ void k_cn_morph(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int H, int W, int WW, int k,
                           long long words) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= words) return;
  int wx = (int)(i % WW);
  long long row = i / WW;
  int y = (int)(row % H);
  const uint32_t* frame = in + (row - y) * WW;
  int a = k / 2;
  uint32_t acc = ERODE ? 0xffffffffu : 0u;
  for (int dy = -a; dy <= k - 1 - a; ++dy) {
    int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    uint32_t v = hwindow<ERODE>(frame + (long long)yy * WW, wx, WW, W, k, a);
    acc = ERODE ? (acc & v) : (acc | v);
  }
  int tail = W & 31;
  if (tail && wx == WW - 1) acc &= (1u << tail) - 1u;
  out[i] = acc;
}

// ------------------------------------------------------------------------ union-find
__device__ __forceinline__ int uf_find(const int* __restrict__ L, int x) {
  while (true) {
    int l = L[x];
    if (l < 0) return -1;
    if (l == x) return x;
    x = l;
  }
}

__device__ void uf_merge(int* L, int a, int b) {
  while (true) {
    a = a < 0 ? -1 : uf_find(L, a);
    b = b < 0 ? -1 : uf_find(L, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }      // a > b >= -1: hang a under b
    int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;                                      // somebody re-parented a meanwhile: retry from there
  }
}

// 34-bit window of a bit row around word wx: bit (lane + 1) = pixel x, bit lane = x - 1, bit (lane + 2) = x + 1
__device__ __forceinline__ uint64_t window(const uint32_t* __restrict__ row, int wx, int WW) {
  uint32_t lo = wx > 0 ? row[wx - 1] : 0u;
  uint32_t hi = wx + 1 < WW ? row[wx + 1] : 0u;
  return ((uint64_t)(hi & 1u) << 33) | ((uint64_t)row[wx] << 1) | (uint64_t)(lo >> 31);
}

// label[p] = head of p's run inside its 32-pixel word (same class to the left); run heads reset their node
__global__ void k_cn_init(const uint32_t* __restrict__ bits, int* __restrict__ L, Node* __restrict__ nodes, int H, int W,
                          int WW, long long threads) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= threads) return;
  int lane = (int)(t & 31);
  long long word = t >> 5;
  int wx = (int)(word % WW);
  long long row = word / WW;
  int x = wx * 32 + lane;
  if (x >= W) return;
  uint32_t m = bits[word];
  uint32_t cls = ((m >> lane) & 1u) ? m : ~m;
  uint32_t other_below = ~cls & ((1u << lane) - 1u);
  int head = other_below ? 32 - __clz(other_below) : 0;
  int y = (int)(row % H);
  int p = y * W + x;
  long long base = (row - y) * W;
  L[base + p] = p - lane + head;
  if (head == lane) nodes[base + p] = Node{-1, 0, 0, 0};
}

__global__ void k_cn_merge(const uint32_t* __restrict__ bits, int* __restrict__ Lall, int H, int W, int WW,
                           long long threads) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= threads) return;
  int lane = (int)(t & 31);
  long long word = t >> 5;
  int wx = (int)(word % WW);
  long long row = word / WW;
  int x = wx * 32 + lane;
  if (x >= W) return;
  int y = (int)(row % H);
  const uint32_t* frame = bits + (row - y) * WW;
  int* L = Lall + (row - y) * W;
  int p = y * W + x;
  uint64_t cur = window(frame + (long long)y * WW, wx, WW);
  uint64_t up = y > 0 ? window(frame + (long long)(y - 1) * WW, wx, WW) : 0ull;
  bool fg = (cur >> (lane + 1)) & 1ull;
  // validity window (pixels inside the row), same bit convention
  uint64_t valid = 0;
  {
    int lo = wx * 32 - 1;                       // pixel index of window bit 0
    int n_lo = lo < 0 ? 1 : 0;                  // bit 0 invalid for the first word
    int hi_px = W - lo;                         // bits [n_lo, hi_px) are pixels
    valid = hi_px >= 34 ? 0x3ffffffffull : ((1ull << hi_px) - 1ull);
    if (n_lo) valid &= ~1ull;
  }
  uint64_t curc = (fg ? cur : ~cur) & valid;    // same class as p, current row
  uint64_t upc = y > 0 ? ((fg ? up : ~up) & valid) : 0ull;
  uint64_t v = curc & upc;
  int b = lane + 1;
  // vertical link, once per overlap run
  if (((v >> b) & 1ull) && !(lane > 0 && ((v >> (b - 1)) & 1ull))) uf_merge(L, p, p - W);
  // runs continue across the word boundary
  if (lane == 0 && ((curc >> (b - 1)) & 1ull)) uf_merge(L, p, p - 1);
  if (fg) {
    // diagonal links only where no 4-connected detour exists
    bool up_bg = !((up >> b) & 1ull);
    if (up_bg && ((up >> (b - 1)) & 1ull) && !((cur >> (b - 1)) & 1ull)) uf_merge(L, p, p - W - 1);
    if (up_bg && ((up >> (b + 1)) & 1ull) && !((cur >> (b + 1)) & 1ull)) uf_merge(L, p, p - W + 1);
  } else if (x == 0 || y == 0 || x == W - 1 || y == H - 1) {
    uf_merge(L, p, -1);                          // background touching the frame = exterior
  }
}

// flatten; roots learn their enclosing node; strip pixels are counted into their node
__global__ void k_cn_flatten(const uint32_t* __restrict__ bits, int* __restrict__ Lall, Node* __restrict__ nall, int H,
                             int W, int WW, int y_top, long long threads) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= threads) return;
  int lane = (int)(t & 31);
  long long word = t >> 5;
  int wx = (int)(word % WW);
  long long row = word / WW;
  int x = wx * 32 + lane;
  int y = (int)(row % H);
  int* L = Lall + (row - y) * W;
  Node* nodes = nall + (row - y) * W;
  const uint32_t* frame = bits + (row - y) * WW;
  bool live = x < W;
  int p = y * W + x;
  int r = -1;
  bool fg = false;
  if (live) {
    r = uf_find(L, p);
    L[p] = r;
    fg = (frame[(long long)y * WW + wx] >> lane) & 1u;
    if (r == p) {
      // component: the pixel left of the raster-first pixel lies in the enclosing region;
      // hole: the pixel above the raster-first pixel belongs to the enclosing component
      nodes[p].up = fg ? (x == 0 ? -1 : uf_find(L, p - 1)) : uf_find(L, p - W);
    }
  }
  if (y < y_top) return;                          // whole warp: rows are warp-uniform
  // strip counts, one atomic per run of equal roots inside the warp
  unsigned act = __ballot_sync(0xffffffffu, live && r >= 0);
  int prev = __shfl_up_sync(0xffffffffu, r, 1);
  bool head = live && r >= 0 && (lane == 0 || prev != r);
  unsigned heads = __ballot_sync(0xffffffffu, head);
  if (head) {
    unsigned later = heads & ~((2u << lane) - 1u);       // heads strictly after this lane
    unsigned same = act & ~((1u << lane) - 1u);          // active lanes from here on
    int end = later ? __ffs(later) - 1 : 32;
    // lanes [lane, end) share r unless an inactive lane (exterior / outside) interrupts: count actives with equal root
    int n = 0;
    for (int l = lane; l < end; ++l) n += (same >> l) & 1u;
    // an inactive lane inside [lane, end) would have made the next active lane a head, so this is exact
    atomicAdd(&nodes[r].own, n);
  }
  if (live && fg) {
    // ring pixels: component pixels 4-adjacent to one of the component's own holes
    int upC = r % W == 0 ? -1 : uf_find(L, r - 1);      // region enclosing this pixel's component
    int seen[4];
    int ns = 0;
    const int dx[4] = {0, 0, -1, 1}, dy[4] = {-1, 1, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int xx = x + dx[q], yy = y + dy[q];
      if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
      if ((frame[(long long)yy * WW + (xx >> 5)] >> (xx & 31)) & 1u) continue;
      int h = uf_find(L, yy * W + xx);
      if (h < 0 || h == upC) continue;
      bool dup = false;
      for (int s = 0; s < ns; ++s) dup |= seen[s] == h;
      if (dup) continue;
      seen[ns++] = h;
      atomicAdd(&nodes[h].ring, 1);
    }
  }
}

// every node with strip pixels adds them to itself and to everything that encloses it
__global__ void k_cn_accumulate(const int* __restrict__ Lall, Node* __restrict__ nall, int HW, long long pixels) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  int p = (int)(i % HW);
  if (Lall[i] != p) return;
  Node* nodes = nall + (i - p);
  int own = nodes[p].own;
  if (own == 0) return;
  for (int a = p; a >= 0; a = nodes[a].up) atomicAdd(&nodes[a].enc, own);
}

__global__ void k_cn_fill(const uint32_t* __restrict__ bits, const int* __restrict__ Lall, const Node* __restrict__ nall,
                          uint8_t* __restrict__ out, int H, int W, int WW, int thr, long long threads) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= threads) return;
  int lane = (int)(t & 31);
  long long word = t >> 5;
  int wx = (int)(word % WW);
  long long row = word / WW;
  int x = wx * 32 + lane;
  if (x >= W) return;
  int y = (int)(row % H);
  const int* L = Lall + (row - y) * W;
  const Node* nodes = nall + (row - y) * W;
  const uint32_t* frame = bits + (row - y) * WW;
  int p = y * W + x;
  bool fg = (frame[(long long)y * WW + wx] >> lane) & 1u;
  int node = L[p];
  int par = 0;
  bool drawn = false;
  if (fg) {
    Node c = nodes[node];
    bool keep = c.enc >= thr;
    bool on_outer = false;
    const int dx[4] = {0, 0, -1, 1}, dy[4] = {-1, 1, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int xx = x + dx[q], yy = y + dy[q];
      if (xx < 0 || yy < 0 || xx >= W || yy >= H) { on_outer = true; continue; }
      if ((frame[(long long)yy * WW + (xx >> 5)] >> (xx & 31)) & 1u) continue;
      int h = L[yy * W + xx];
      if (h < 0 || h == c.up) { on_outer = true; continue; }
      Node hn = nodes[h];
      drawn |= hn.enc + hn.ring >= thr;           // on the chain of a kept hole contour
    }
    drawn |= keep && on_outer;                    // on the chain of the kept outer contour
    par = keep;
    node = c.up;
  }
  // node is now a background region: walk hole -> component -> hole ... up to the exterior
  while (!drawn && node >= 0) {
    Node hn = nodes[node];
    par ^= (hn.enc + hn.ring >= thr);
    Node cn = nodes[hn.up];
    par ^= (cn.enc >= thr);
    node = cn.up;
  }
  out[(row - y) * W + p] = drawn ? 1 : (uint8_t)par;
}

}  // namespace

size_t contour_scratch_bytes(int B, int H, int W) {
  size_t WW = (size_t)(W + 31) / 32;
  size_t px = (size_t)B * H * W;
  size_t bits = (size_t)B * H * WW * 4;
  return 2 * ((bits + 255) & ~(size_t)255) + ((px * 4 + 255) & ~(size_t)255) + px * sizeof(Node);
}

int contour_launch_count() { return 8; }

void launch_contour_noise_removal(const uint8_t* seg, int H, int W, int B, uint8_t* out, void* scratch, cudaStream_t s) {
  int WW = (W + 31) / 32;
  long long rows = (long long)B * H;
  long long words = rows * WW;
  long long threads = words * 32;
  long long px = rows * W;
  size_t bits_bytes = ((size_t)words * 4 + 255) & ~(size_t)255;
  uint8_t* base = (uint8_t*)scratch;
  uint32_t* bitsA = (uint32_t*)base;
  uint32_t* bitsB = (uint32_t*)(base + bits_bytes);
  int* L = (int*)(base + 2 * bits_bytes);
  Node* nodes = (Node*)(base + 2 * bits_bytes + (((size_t)px * 4 + 255) & ~(size_t)255));
  int k = (int)(std::min(H, W) / 50);                                  // image_processing_utils.py:7-8
  int y_top = (int)(H * (1 - 0.1));                                    // :19,22 (fp64, as Python evaluates it)
  double T = (double)((long long)W * (H - y_top)) * 0.4;               // :27,31,38
  int thr = (int)floor(T) + 1;                                         // area > T  <=>  area >= thr
  const int TB = 256;
  auto grid = [&](long long n) { return (unsigned)((n + TB - 1) / TB); };
  k_cn_pack<<<grid(threads), TB, 0, s>>>(seg, bitsA, H, W, WW, rows);
  k_cn_morph<false><<<grid(words), TB, 0, s>>>(bitsA, bitsB, H, W, WW, k, words);
  k_cn_morph<true><<<grid(words), TB, 0, s>>>(bitsB, bitsA, H, W, WW, k, words);
  k_cn_init<<<grid(threads), TB, 0, s>>>(bitsA, L, nodes, H, W, WW, threads);
  k_cn_merge<<<grid(threads), TB, 0, s>>>(bitsA, L, H, W, WW, threads);
  k_cn_flatten<<<grid(threads), TB, 0, s>>>(bitsA, L, nodes, H, W, WW, y_top, threads);
  k_cn_accumulate<<<grid(px), TB, 0, s>>>(L, nodes, H * W, px);
  k_cn_fill<<<grid(threads), TB, 0, s>>>(bitsA, L, nodes, out, H, W, WW, thr, threads);
}

}  // namespace bc
